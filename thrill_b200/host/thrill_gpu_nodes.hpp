/*******************************************************************************
 * thrill_b200/host/thrill_gpu_nodes.hpp — the C++ host side of the drop-in.
 *
 * Two DOpNode classes with exactly the StageBuilder protocol of the reference's SortNode
 * (thrill/api/sort.hpp:64-271) and ReduceNode (thrill/api/reduce_by_key.hpp:64-211), and the front doors
 *     thrill_gpu::Sort(dia [, std::less<T>/std::greater<T>])     <->  DIA<T>::Sort      (api/sort.hpp:800)
 *     thrill_gpu::ReducePair(dia, std::plus<double> ...)          <->  DIA<T>::ReducePair(api/reduce_by_key.hpp:410)
 * Everything else of the pipeline (sources, LOps, other DOps, actions, the net/data layers) is the
 * UNMODIFIED reference library: this header only includes it.  The heavy lifting happens behind the C ABI
 * of include/thrill_gpu.h (libthrill_gpu.so): the nodes hand the Blocks of their input data::File to
 * tg_sort_file / tg_reduce_file and wrap the result bytes in fresh ByteBlocks of the worker's BlockPool with
 * the geometry BlockWriter would have produced (tg_file_geometry), so children see an ordinary data::File
 * (PushFile, thrill/api/dia_node.hpp:156-180).
 *
 * One Thrill worker thread = one GPU = one tg_ctx (device = Context::local_worker_id()); the NCCL id is
 * created by worker 0 and broadcast over the reference's own control plane (ctx.net.Broadcast,
 * net/flow_control_channel.hpp:424).  Only a closed set of (type, functor) pairs maps to a descriptor;
 * anything else is a compile-time error (static_assert) — there is no CPU fallback inside these nodes:
 * a user who wants the CPU path calls the stock dia.Sort() / dia.ReducePair().
 * Requires Release builds (common::g_self_verify == false, common/config.hpp:32), which is asserted.
 ******************************************************************************/
#pragma once
#ifndef THRILL_GPU_NODES_HEADER
#define THRILL_GPU_NODES_HEADER

#include <thrill/api/dia.hpp>
#include <thrill/api/dop_node.hpp>
#include <thrill/common/config.hpp>
#include <thrill/data/file.hpp>

#include <array>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/thrill_gpu.h"

namespace thrill_gpu {

using thrill::api::Context;
using thrill::api::DIA;
using thrill::api::DIAMemUse;

//! non-zero status of the C ABI -> die() (tlx::DieException), the reference's error path
inline void Check(tg_ctx* c, int status, const char* what) {
    if (status != TG_OK)
        die("thrill_gpu: " << what << " failed: " << tg_strerror(status) << ": " << tg_last_error(c));
}

//! One tg_ctx per worker THREAD, created on first use (collective: every worker must reach its first GPU node) and shut
//! down when the thread ends.  api::Run starts every worker on a thread of its own (api/context.cpp:RunLoopbackThreads /
//! RunBackendLoopback), so the ctx — stream, workspaces, exchange window, NCCL communicator — lives exactly as long as the
//! job's Context does: a second api::Run in the same process gets fresh threads and a fresh ctx, never a stale one.  On a
//! hit the cached ctx is checked against the Context (a re-used thread, api::RunLocalSameThread) and rebuilt if it differs.
struct WorkerCtxHolder {
    tg_ctx* c = nullptr;
    const Context* owner = nullptr;
    size_t rank = 0, nranks = 0;
    void Reset() { if (c) tg_shutdown(c); c = nullptr; owner = nullptr; }
    ~WorkerCtxHolder() { Reset(); }
};

inline tg_ctx * WorkerCtx(Context& ctx) {
    static thread_local WorkerCtxHolder holder;
    if (holder.c && holder.owner == &ctx && holder.rank == ctx.my_rank() && holder.nranks == ctx.num_workers())
        return holder.c;
    holder.Reset();
    die_unless(!thrill::common::g_self_verify);     // Debug builds prefix every item with a typecode
    using Id = std::array<char, 128>;
    Id id;
    id.fill(0);
    if (ctx.num_workers() > 1) {
        if (ctx.my_rank() == 0) Check(nullptr, tg_get_unique_id(id.data()), "tg_get_unique_id");
        id = ctx.net.Broadcast(id, 0);
    }
    // worker -> GPU: the worker's index on its host (api/context.hpp: local_worker_id).  The in-process test networks
    // (THRILL_NET=mock/local with several "hosts") put every host's workers into this one process: there the global rank
    // picks the GPU, and two workers are never given the same device (NCCL and the peer mapping both need distinct GPUs).
    int ndev = tg_device_count();
    if (ndev <= 0) die("thrill_gpu: no sm_100 GPU visible (there is no CPU fallback in the GPU nodes)");
    size_t device = ctx.local_worker_id();
    const char* net = getenv("THRILL_NET");
    if (ctx.num_hosts() > 1 && net && (!strcmp(net, "mock") || !strcmp(net, "local"))) device = ctx.my_rank();
    if (device >= static_cast<size_t>(ndev))
        die("thrill_gpu: worker " << ctx.my_rank() << " needs GPU " << device << " but only " << ndev << " are visible "
            "(run with THRILL_WORKERS_PER_HOST <= number of GPUs)");
    tg_ctx* c = nullptr;
    int st = tg_init(static_cast<int>(device), static_cast<int>(ctx.my_rank()),
                     static_cast<int>(ctx.num_workers()), id.data(), &c);
    Check(c, st, "tg_init");
    holder.c = c;
    holder.owner = &ctx;
    holder.rank = ctx.my_rank();
    holder.nranks = ctx.num_workers();
    return c;
}

/******************************************************************************/
// descriptors of the recognised (type, functor) pairs

template <typename ValueType, typename Compare, typename Enable = void>
struct SortDesc {
    static constexpr bool supported = false;
};
//! Items whose order is the lexicographic order of a run of key bytes (TeraSort's Record{uint8 key[10]; uint8 value[90]} with
//! operator< = std::lexicographical_compare of the keys, examples/terasort/terasort.cpp:31-42): specialise for the item type
//!   template <> struct thrill_gpu::ByteKeyTraits<Record> { static constexpr bool is_byte_key = true;
//!                                                          static constexpr uint32_t key_offset = 0, key_bytes = 10; };
//! and thrill_gpu::Sort(dia) / Sort(dia, std::less<Record>()) take the GPU path (key_bytes <= 12, sizeof(T) % 4 == 0).
template <typename ValueType>
struct ByteKeyTraits {
    static constexpr bool is_byte_key = false;
};
template <typename ValueType>
struct SortDesc<ValueType, std::less<ValueType>, typename std::enable_if<ByteKeyTraits<ValueType>::is_byte_key>::type>{
    static_assert(sizeof(ValueType) % 4 == 0 && ByteKeyTraits<ValueType>::key_bytes <= 12 &&
                  ByteKeyTraits<ValueType>::key_offset + ByteKeyTraits<ValueType>::key_bytes <= sizeof(ValueType),
                  "byte-key records: sizeof % 4 == 0 and a key of at most 12 bytes inside the item");
    static constexpr bool supported = true;
    static tg_key_desc make() {
        return tg_key_desc { static_cast<uint32_t>(sizeof(ValueType)), ByteKeyTraits<ValueType>::key_offset,
                             ByteKeyTraits<ValueType>::key_bytes, TG_KEY_BYTES_BE, 0, 0 };
    }
};
template <typename Compare>
struct SortDesc<uint64_t, Compare, typename std::enable_if<
                    std::is_same<Compare, std::less<uint64_t> >::value ||
                    std::is_same<Compare, std::greater<uint64_t> >::value>::type>{
    static constexpr bool supported = true;
    static tg_key_desc make() {
        return tg_key_desc { 8, 0, 8, TG_KEY_UINT_LE,
                             std::is_same<Compare, std::greater<uint64_t> >::value ? 1u : 0u, 0 };
    }
};
//! pair<uint64_t, 8-byte POD> ordered by .first (serialized member-wise as 16 bytes, data/serialization.hpp:67-84)
struct LessFirst {
    template <typename P>
    bool operator () (const P& a, const P& b) const { return a.first < b.first; }
};
template <typename V>
struct SortDesc<std::pair<uint64_t, V>, LessFirst,
                typename std::enable_if<sizeof(V) == 8 && std::is_pod<V>::value>::type>{
    static constexpr bool supported = true;
    static tg_key_desc make() { return tg_key_desc { 16, 0, 8, TG_KEY_UINT_LE, 0, 1 }; }
};

template <typename Value, typename ReduceFunction>
struct ReduceDesc {
    static constexpr bool supported = false;
};
template <>
struct ReduceDesc<double, std::plus<double> >{
    static constexpr bool supported = true;
    static constexpr uint32_t op = TG_OP_SUM_F64;
};
template <>
struct ReduceDesc<uint64_t, std::plus<uint64_t> >{
    static constexpr bool supported = true;
    static constexpr uint32_t op = TG_OP_SUM_U64;
};
struct MinU64 { uint64_t operator () (uint64_t a, uint64_t b) const { return b < a ? b : a; } };
struct MaxU64 { uint64_t operator () (uint64_t a, uint64_t b) const { return a < b ? b : a; } };
template <>
struct ReduceDesc<uint64_t, MinU64>{
    static constexpr bool supported = true;
    static constexpr uint32_t op = TG_OP_MIN_U64;
};
template <>
struct ReduceDesc<uint64_t, MaxU64>{
    static constexpr bool supported = true;
    static constexpr uint32_t op = TG_OP_MAX_U64;
};

//! The functors thrill_gpu::ReduceByKey recognises (DIA::ReduceByKey takes a key extractor and a reduce function over whole
//! items, api/reduce_by_key.hpp:312-363): the key is pair.first, the reduce function folds pair.second and keeps the key.
struct KeyFirst {
    template <typename P>
    const typename P::first_type& operator () (const P& p) const { return p.first; }
};
template <typename ValueFunction>
struct OnSecond {
    ValueFunction fn;
    template <typename P>
    P operator () (const P& a, const P& b) const { return P(a.first, fn(a.second, b.second)); }
};

/******************************************************************************/
// shared File <-> C ABI plumbing

//! pin every Block of a File and describe it for the C ABI (replaces File::GetReader + per-item Next)
class PinnedFileView
{
public:
    PinnedFileView(const thrill::data::File& file, size_t local_worker_id) {
        pins_.reserve(file.num_blocks());
        blocks_.reserve(file.num_blocks());
        for (const thrill::data::Block& b : file.blocks()) {
            pins_.emplace_back(b.PinWait(local_worker_id));
            const thrill::data::PinnedBlock& pb = pins_.back();
            blocks_.push_back(tg_block { pb.data_begin(), pb.size() });
        }
    }
    const tg_block * data() const { return blocks_.data(); }
    size_t size() const { return blocks_.size(); }

private:
    std::vector<thrill::data::PinnedBlock> pins_;
    std::vector<tg_block> blocks_;
};

/******************************************************************************/
// GPU node -> GPU node hand-off without the PCIe round trip

//! A node result that lives in HBM (tg_dev_file); shared by the parent and the children it was handed to.
class DeviceFile
{
public:
    DeviceFile(tg_ctx* c, const tg_dev_file& f) : ctx_(c), f_(f) { }
    DeviceFile(const DeviceFile&) = delete;
    DeviceFile& operator = (const DeviceFile&) = delete;
    ~DeviceFile() { tg_dev_file_free(ctx_, &f_); }
    const tg_dev_file * get() const { return &f_; }
    size_t items() const { return f_.items; }

private:
    tg_ctx* ctx_;
    tg_dev_file f_;
};
using DeviceFilePtr = std::shared_ptr<DeviceFile>;

//! Implemented by the GPU nodes: a parent GPU node offers its device-resident result instead of a data::File.  The offer is
//! made inside the parent's PushData, i.e. between the child's StartPreOp and StopPreOp, exactly where OnPreOpFile would be
//! called (api/dia_node.hpp:156-180); a child whose function stack towards this parent is not empty declines.
class GpuNodeBase
{
public:
    virtual ~GpuNodeBase() { }
    virtual bool OnPreOpDeviceFile(const DeviceFilePtr& file, size_t item_bytes) = 0;
};

//! materialise a device File as a host data::File (the lazy D2H, only for children that are not GPU nodes)
inline void FetchDeviceFileIntoFile(tg_ctx* c, Context& ctx, const DeviceFile& df, uint32_t item_bytes,
                                    thrill::data::File& out) {
    size_t num_items = df.items();
    size_t nblocks = tg_file_geometry(num_items, item_bytes, thrill::data::start_block_size,
                                      thrill::data::default_block_size, nullptr, 0);
    std::vector<tg_block_geom> geom(nblocks);
    tg_file_geometry(num_items, item_bytes, thrill::data::start_block_size,
                     thrill::data::default_block_size, geom.data(), geom.size());
    std::vector<thrill::data::PinnedByteBlockPtr> bytes;
    std::vector<tg_block_mut> targets;
    bytes.reserve(nblocks);
    for (const tg_block_geom& g : geom) {
        size_t cap = thrill::data::start_block_size;
        while (cap < g.bytes) cap *= 2;
        bytes.emplace_back(ctx.block_pool().AllocateByteBlock(cap, ctx.local_worker_id()));
        targets.push_back(tg_block_mut { bytes.back()->data(), static_cast<size_t>(g.bytes) });
    }
    Check(c, tg_dev_file_fetch(c, df.get(), targets.data(), targets.size()), "tg_dev_file_fetch");
    for (size_t i = 0; i < nblocks; ++i) {
        thrill::data::PinnedBlock pb(std::move(bytes[i]), 0, geom[i].bytes, geom[i].first_item,
                                     geom[i].num_items, /* typecode_verify */ false);
        out.AppendBlock(std::move(pb).MoveToBlock());
    }
}

//! true if every child of `node` is a GPU node (so no host File is needed)
template <typename Node>
bool AllChildrenAreGpuNodes(const Node& node) {
    std::vector<thrill::api::DIABase*> ch = node.children();
    if (ch.empty()) return false;
    for (thrill::api::DIABase* c : ch)
        if (dynamic_cast<GpuNodeBase*>(c) == nullptr) return false;
    return true;
}

/******************************************************************************/

template <typename ValueType>
class GpuSortNode final : public thrill::api::DOpNode<ValueType>, public GpuNodeBase
{
    using Super = thrill::api::DOpNode<ValueType>;
    using Super::context_;

public:
    template <typename ParentDIA>
    GpuSortNode(const ParentDIA& parent, const tg_key_desc& desc)
        : Super(parent.ctx(), "GpuSort", { parent.id() }, { parent.node() }),
          desc_(desc), parent_stack_empty_(ParentDIA::stack_empty) {
        // hook the per-item PreOp exactly as SortNode does (api/sort.hpp:131-136)
        auto pre_op_fn = [this](const ValueType& input) { unsorted_writer_.Put(input); };
        auto lop_chain = parent.stack().push(pre_op_fn).fold();
        parent.node()->AddChild(this, lop_chain);
    }

    void StartPreOp(size_t /* parent_index */) final { unsorted_writer_ = unsorted_file_.GetWriter(); }

    //! whole-File hand-off: the bulk path (api/sort.hpp:151-175)
    bool OnPreOpFile(const thrill::data::File& file, size_t /* parent_index */) final {
        if (!parent_stack_empty_) return false;
        unsorted_file_ = file.Copy();
        return true;
    }

    //! a parent GPU node hands its result over in HBM
    bool OnPreOpDeviceFile(const DeviceFilePtr& file, size_t item_bytes) final {
        if (!parent_stack_empty_ || item_bytes != sizeof(ValueType)) return false;
        device_input_ = file;
        return true;
    }

    void StopPreOp(size_t /* parent_index */) final { unsorted_writer_.Close(); }

    DIAMemUse ExecuteMemUse() final { return DIAMemUse::Max(); }

    //! MainOp (api/sort.hpp:537-663) + the local sort, all behind tg_sort_file / tg_sort_dev.  Collective.  The result stays
    //! in HBM; PushData hands it to GPU children as it is and writes a host File only if another kind of child needs one.
    void Execute() final {
        tg_ctx* c = WorkerCtx(context_);
        size_t out_items = 0;
        if (device_input_) {
            Check(c, tg_sort_dev(c, &desc_, device_input_->get(), context_.rng_(), &out_items), "tg_sort_dev");
            device_input_.reset();
        }
        else {
            PinnedFileView view(unsorted_file_, context_.local_worker_id());
            Check(c, tg_sort_file(c, &desc_, view.data(), view.size(), context_.rng_(), &out_items), "tg_sort_file");
        }
        unsorted_file_.Clear();
        tg_dev_file f;
        Check(c, tg_output_detach(c, &f), "tg_output_detach");
        device_result_ = std::make_shared<DeviceFile>(c, f);
        have_host_file_ = false;
    }

    DIAMemUse PushDataMemUse() final { return 0; }

    //! one sorted run per worker: always the files_.size() == 1 branch of SortNode::PushData (:224-227)
    void PushData(bool consume) final {
        if (device_result_ && AllChildrenAreGpuNodes(*this)) {
            bool all = true;
            for (thrill::api::DIABase* ch : this->children())
                all = dynamic_cast<GpuNodeBase*>(ch)->OnPreOpDeviceFile(device_result_, sizeof(ValueType)) && all;
            if (all) return;
        }
        if (!have_host_file_) {
            FetchDeviceFileIntoFile(WorkerCtx(context_), context_, *device_result_, sizeof(ValueType), sorted_file_);
            have_host_file_ = true;
        }
        this->PushFile(sorted_file_, consume);
    }

    void Dispose() final { sorted_file_.Clear(); device_result_.reset(); have_host_file_ = false; }

private:
    tg_key_desc desc_;
    const bool parent_stack_empty_;
    thrill::data::File unsorted_file_ { context_.GetFile(this) };
    thrill::data::File::Writer unsorted_writer_;
    thrill::data::File sorted_file_ { context_.GetFile(this) };
    DeviceFilePtr device_input_, device_result_;
    bool have_host_file_ = false;
};

template <typename ValueType>
class GpuReduceNode final : public thrill::api::DOpNode<ValueType>, public GpuNodeBase
{
    using Super = thrill::api::DOpNode<ValueType>;
    using Super::context_;

public:
    template <typename ParentDIA>
    GpuReduceNode(const ParentDIA& parent, const tg_kv_desc& desc)
        : GpuReduceNode(parent, desc, false, 0, ValueType()) { }

    //! to_index: ReduceToIndexNode (api/reduce_to_index.hpp:60-237) — dense result of result_size items, neutral_element
    //! where no item has that index; worker r holds the index range Range(0, size).Partition(r, p)
    template <typename ParentDIA>
    GpuReduceNode(const ParentDIA& parent, const tg_kv_desc& desc, bool to_index, size_t result_size,
                  const ValueType& neutral_element)
        : Super(parent.ctx(), to_index ? "GpuReduceToIndex" : "GpuReducePair", { parent.id() }, { parent.node() }),
          desc_(desc), parent_stack_empty_(ParentDIA::stack_empty),
          to_index_(to_index), result_size_(result_size), neutral_(neutral_element) {
        // ReduceNode inserts each item into the pre-phase table (api/reduce_by_key.hpp:126-133); the GPU pre
        // phase wants the whole shard, so items are collected in a File first
        auto pre_op_fn = [this](const ValueType& input) { input_writer_.Put(input); };
        auto lop_chain = parent.stack().push(pre_op_fn).fold();
        parent.node()->AddChild(this, lop_chain);
    }

    DIAMemUse PreOpMemUse() final { return DIAMemUse::Max(); }

    void StartPreOp(size_t /* parent_index */) final { input_writer_ = input_file_.GetWriter(); }

    bool OnPreOpFile(const thrill::data::File& file, size_t /* parent_index */) final {
        if (!parent_stack_empty_) return false;
        input_file_ = file.Copy();
        return true;
    }

    bool OnPreOpDeviceFile(const DeviceFilePtr& file, size_t item_bytes) final {
        if (!parent_stack_empty_ || item_bytes != sizeof(ValueType)) return false;
        device_input_ = file;
        return true;
    }

    //! pre phase flush + exchange + post phase (api/reduce_by_key.hpp:157-211) behind tg_reduce_file / tg_reduce_dev.
    //! Collective.  The result stays in HBM until PushData knows who wants it.
    void StopPreOp(size_t /* parent_index */) final {
        input_writer_.Close();
        tg_ctx* c = WorkerCtx(context_);
        size_t out_items = 0;
        static_assert(sizeof(ValueType) == 16, "16-byte (key, value) items");
        uint64_t begin = 0;
        if (device_input_) {
            if (to_index_)
                Check(c, tg_reduce_to_index_dev(c, &desc_, device_input_->get(), result_size_, &neutral_, &out_items, &begin),
                      "tg_reduce_to_index_dev");
            else
                Check(c, tg_reduce_dev(c, &desc_, device_input_->get(), &out_items), "tg_reduce_dev");
            device_input_.reset();
        }
        else {
            PinnedFileView view(input_file_, context_.local_worker_id());
            if (to_index_)
                Check(c, tg_reduce_to_index_file(c, &desc_, view.data(), view.size(), result_size_, &neutral_, &out_items, &begin),
                      "tg_reduce_to_index_file");
            else
                Check(c, tg_reduce_file(c, &desc_, view.data(), view.size(), &out_items), "tg_reduce_file");
        }
        input_file_.Clear();
        tg_dev_file f;
        Check(c, tg_output_detach(c, &f), "tg_output_detach");
        device_result_ = std::make_shared<DeviceFile>(c, f);
        have_host_file_ = false;
    }

    void Execute() final { }

    DIAMemUse PushDataMemUse() final { return 0; }

    void PushData(bool consume) final {
        if (device_result_ && AllChildrenAreGpuNodes(*this)) {
            bool all = true;
            for (thrill::api::DIABase* ch : this->children())
                all = dynamic_cast<GpuNodeBase*>(ch)->OnPreOpDeviceFile(device_result_, sizeof(ValueType)) && all;
            if (all) return;
        }
        if (!have_host_file_) {
            FetchDeviceFileIntoFile(WorkerCtx(context_), context_, *device_result_, sizeof(ValueType), reduced_file_);
            have_host_file_ = true;
        }
        this->PushFile(reduced_file_, consume);
    }

    void Dispose() final { reduced_file_.Clear(); device_result_.reset(); have_host_file_ = false; }

private:
    tg_kv_desc desc_;
    const bool parent_stack_empty_;
    const bool to_index_ = false;
    const size_t result_size_ = 0;
    const ValueType neutral_ = ValueType();
    thrill::data::File input_file_ { context_.GetFile(this) };
    thrill::data::File::Writer input_writer_;
    thrill::data::File reduced_file_ { context_.GetFile(this) };
    DeviceFilePtr device_input_, device_result_;
    bool have_host_file_ = false;
};

/******************************************************************************/
// front doors (same argument meaning as DIA<T>::Sort / DIA<T>::ReducePair)

template <typename ValueType, typename Stack, typename CompareFunction = std::less<ValueType> >
auto Sort(const DIA<ValueType, Stack>& dia, const CompareFunction& /* compare_function */ = CompareFunction()) {
    static_assert(SortDesc<ValueType, CompareFunction>::supported,
                  "thrill_gpu::Sort: this (ValueType, CompareFunction) pair has no GPU descriptor; "
                  "use the stock dia.Sort(cmp)");
    assert(dia.IsValid());
    auto node = tlx::make_counting<GpuSortNode<ValueType> >(
        dia, SortDesc<ValueType, CompareFunction>::make());
    return DIA<ValueType>(node);
}

//! DIA<T>::SortStable(cmp) (api/sort.hpp:873-937): equal keys keep their global input order.  The GPU sort is stable by
//! construction (stable partition passes, exchange in worker order), so this is the same node with the flag set.
template <typename ValueType, typename Stack, typename CompareFunction = std::less<ValueType> >
auto SortStable(const DIA<ValueType, Stack>& dia, const CompareFunction& /* compare_function */ = CompareFunction()) {
    static_assert(SortDesc<ValueType, CompareFunction>::supported,
                  "thrill_gpu::SortStable: this (ValueType, CompareFunction) pair has no GPU descriptor; "
                  "use the stock dia.SortStable(cmp)");
    assert(dia.IsValid());
    tg_key_desc d = SortDesc<ValueType, CompareFunction>::make();
    d.stable = 1;
    auto node = tlx::make_counting<GpuSortNode<ValueType> >(dia, d);
    return DIA<ValueType>(node);
}

template <typename Key, typename Value, typename Stack, typename ReduceFunction>
auto ReducePair(const DIA<std::pair<Key, Value>, Stack>& dia, const ReduceFunction& /* reduce_function */) {
    static_assert(std::is_same<Key, uint64_t>::value && sizeof(Value) == 8 &&
                  ReduceDesc<Value, ReduceFunction>::supported,
                  "thrill_gpu::ReducePair: this (Key, Value, ReduceFunction) has no GPU descriptor; "
                  "use the stock dia.ReducePair(fn)");
    assert(dia.IsValid());
    using ValueType = std::pair<Key, Value>;
    auto node = tlx::make_counting<GpuReduceNode<ValueType> >(
        dia, tg_kv_desc { 16, ReduceDesc<Value, ReduceFunction>::op });
    return DIA<ValueType>(node);
}

//! DIA<T>::ReduceByKey(key_extractor, reduce_function) (api/reduce_by_key.hpp:312-363) for the recognised functor pair:
//! key_extractor = thrill_gpu::KeyFirst, reduce_function = thrill_gpu::OnSecond<F> with F one of the ReducePair functions
//! (std::plus<double>, std::plus<uint64_t>, thrill_gpu::MinU64, thrill_gpu::MaxU64).  ReducePair is exactly this pair of
//! functors in the reference too (:444-449).
template <typename Key, typename Value, typename Stack, typename ValueFunction>
auto ReduceByKey(const DIA<std::pair<Key, Value>, Stack>& dia, const KeyFirst& /* key_extractor */,
                 const OnSecond<ValueFunction>& reduce_function) {
    return ReducePair(dia, reduce_function.fn);
}

//! DIA<pair<uint64_t index, V>>::ReduceToIndex(key = .first, reduce function on .second, size, neutral_element)
//! (api/reduce_to_index.hpp:239-393 front doors; examples/page_rank/page_rank.hpp:125-135)
template <typename Key, typename Value, typename Stack, typename ReduceFunction>
auto ReduceToIndex(const DIA<std::pair<Key, Value>, Stack>& dia, const ReduceFunction& /* reduce_function */,
                   size_t size, const std::pair<Key, Value>& neutral_element = std::pair<Key, Value>()) {
    static_assert(std::is_same<Key, uint64_t>::value && sizeof(Value) == 8 &&
                  ReduceDesc<Value, ReduceFunction>::supported,
                  "thrill_gpu::ReduceToIndex: this (Key, Value, ReduceFunction) has no GPU descriptor; "
                  "use the stock dia.ReduceToIndex(key_extractor, fn, size)");
    assert(dia.IsValid());
    using ValueType = std::pair<Key, Value>;
    auto node = tlx::make_counting<GpuReduceNode<ValueType> >(
        dia, tg_kv_desc { 16, ReduceDesc<Value, ReduceFunction>::op }, true, size, neutral_element);
    return DIA<ValueType>(node);
}

} // namespace thrill_gpu

#endif // !THRILL_GPU_NODES_HEADER
