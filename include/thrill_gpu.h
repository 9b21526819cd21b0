/*******************************************************************************
 * include/thrill_gpu.h — C ABI of libthrill_gpu.so
 *
 * The drop-in boundary for Thrill's Sort / ReduceByKey hot path on B200 (sm_100a).
 * Thrill (thrill/thrill @ 12c5b59b) has no FFI of its own — operators are header templates — so every
 * entry point below cites the reference interface it REPLACES (paths relative to the reference root);
 * INTEGRATION.md shows the host-side DOpNode classes (thrill_b200/host/) that bind them.
 *
 * Conventions: plain pointers and sizes, no C++ types, no exceptions.  Every function returns TG_OK (0)
 * or a negative tg_status; tg_last_error(ctx) gives a message.  The host shim turns non-zero into
 * die() → tlx::DieException, the reference's only error path (api/dia_base.cpp:143-150).
 * One tg_ctx per worker thread / GPU / CUDA stream / NCCL rank (api/context.hpp:243-245: one worker =
 * one DIA shard).  Functions are asynchronous on the ctx stream unless stated; they are not re-entrant
 * per ctx.  Collective entry points (tg_sort, tg_reduce_by_key with nranks > 1) must be entered by all
 * ranks in the same order — the rule Thrill has for GetNewMixStream (api/context.hpp:308-316).
 * Device buffers passed in must be 16-byte aligned.  There is NO CPU fallback anywhere behind this ABI.
 ******************************************************************************/
#ifndef THRILL_GPU_H
#define THRILL_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tg_ctx tg_ctx;

typedef enum {
    TG_OK = 0,
    TG_ERR_CUDA = -1,          /* a CUDA runtime call failed */
    TG_ERR_NCCL = -2,          /* an NCCL call failed */
    TG_ERR_ARG = -3,           /* bad argument / unsupported descriptor */
    TG_ERR_TOO_LARGE = -4,     /* n exceeds the per-call limit (2^30 - 1 items per GPU) */
    TG_ERR_NO_DEVICE = -5,     /* no sm_100 device: the product path fails loudly, never falls back */
    TG_ERR_OOM = -6
} tg_status;

/* ---- descriptors: the closed set of (type, functor) pairs the GPU path accepts ----------------------
 * Reference operators take arbitrary C++ lambdas (api/dia.hpp:1753-1816 Sort(cmp), :929-1170
 * ReduceByKey(key_ex, red_fn)).  The host shim recognises the supported functor types and fills these. */
enum { TG_KEY_UINT_LE = 0,     /* unsigned little-endian integer key of 1..8 bytes inside 8- or 16-byte items: std::less<T> */
       TG_KEY_BYTES_BE = 1 };  /* byte-string key compared lexicographically (TeraSort Record,
                                  examples/terasort/terasort.cpp:35-37): <= 16 bytes in 16-byte items, <= 12 bytes in
                                  records (item_bytes % 4 == 0, e.g. 100) */
typedef struct {
    uint32_t item_bytes;       /* serialized item size (data/serialization.hpp:34-49): 8, 16 or 100 */
    uint32_t key_offset;
    uint32_t key_bytes;
    uint32_t key_kind;         /* TG_KEY_* */
    uint32_t descending;       /* std::greater<T> */
    uint32_t stable;           /* SortStable (api/sort.hpp:873-937); the GPU path is always stable */
} tg_key_desc;

enum { TG_OP_SUM_F64 = 0, TG_OP_SUM_U64 = 1, TG_OP_MIN_U64 = 2, TG_OP_MAX_U64 = 3,
       TG_OP_MIN_F64 = 4, TG_OP_MAX_F64 = 5, TG_OP_FIRST = 6 };
typedef struct {
    uint32_t item_bytes;       /* 16: pair<uint64_t key, 8-byte value> serialized member-wise
                                  (data/serialization.hpp:67-84), the ReducePair TableItem
                                  (api/reduce_by_key.hpp:410-449, core/reduce_functional.hpp:156-209) */
    uint32_t op;               /* TG_OP_*: the recognised ReduceFunction (std::plus<double>, ...) */
} tg_kv_desc;

/* one data::Block of a data::File as the host sees it after PinWait (data/block.hpp:52-145):
 * `data` = ByteBlock::data() + begin, `bytes` = end - begin */
typedef struct {
    const void* data;
    size_t bytes;
} tg_block;
typedef struct {
    void* data;
    size_t bytes;
} tg_block_mut;

/* ---- context ---------------------------------------------------------------------------------------- */
int tg_version(void);
int tg_device_count(void);                          /* sm_100 GPUs visible to this process (0: none, no CPU fallback) */
const char* tg_strerror(int status);
const char* tg_last_error(const tg_ctx* ctx);

/* 128-byte NCCL unique id for the job (rank 0 creates it, the host control plane — Thrill's
 * net::FlowControlChannel broadcast, net/flow_control_channel.hpp:404 — distributes it). */
int tg_get_unique_id(void* out128);
/* Replaces per-worker setup in api/context.cpp:1179 (Context ctor) for the GPU operators: binds `device`
 * (= Context::local_worker_id()), creates the stream and, if nranks > 1, the NCCL communicator that stands
 * in for data::MixStream (data/mix_stream.cpp:52-113).  unique_id may be NULL when nranks == 1. */
int tg_init(int device, int rank, int nranks, const void* unique_id128, tg_ctx** out_ctx);
int tg_shutdown(tg_ctx* ctx);
int tg_rank(const tg_ctx* ctx);
int tg_nranks(const tg_ctx* ctx);
void* tg_stream(const tg_ctx* ctx);                 /* cudaStream_t of the ctx */
int tg_sync(tg_ctx* ctx);                           /* cudaStreamSynchronize */
int tg_barrier(tg_ctx* ctx);                        /* ncclAllReduce of one int + sync: ctx.net.Barrier() */

/* device memory owned by the ctx (BlockPool analogue for HBM, data/block_pool.hpp:40) */
int tg_alloc(tg_ctx* ctx, size_t bytes, void** out_dptr);
int tg_free(tg_ctx* ctx, void* dptr);
/* device timing on the ctx stream (CUDA events) */
int tg_timer_start(tg_ctx* ctx);
int tg_timer_stop(tg_ctx* ctx, float* out_ms);      /* synchronises */
/* kernels launched by this ctx since tg_init (bench.py's gpu_launches) */
uint64_t tg_launch_count(const tg_ctx* ctx);
/* how many local sorts on this ctx abandoned the prefix sort (top digits + finishing pass) for the plain LSD passes
 * because a run of equal key prefixes was too long (heavy duplicates); see tg_radix_sort_local */
uint64_t tg_prefix_sort_fallbacks(const tg_ctx* ctx);
/* records of popular keys that the aggregations on this ctx folded in their counting read, cumulative (such records are read once
 * and never moved: bench.py needs the count for the algorithmic bytes of the hash passes) */
uint64_t tg_hot_records(const tg_ctx* ctx);
/* per-kernel-class device timing (CUDA events around every launch of the class while enabled):
 * bench.py's live roofline measurement.  tg_profile_get synchronises and returns the summed duration
 * and the number of launches of `kernel_class` since tg_profile_enable(ctx, 1). */
enum { TG_K_RADIX_HIST = 0, TG_K_PARTITION = 1, TG_K_MERGE = 2, TG_K_PREAGG = 3, TG_K_AGGREGATE = 4,
       TG_K_COMPACT = 5, TG_K_OTHER = 6, TG_K_FIXUP = 7, TG_K_SEGCOUNT = 8,
       TG_K_EXCHANGE = 9 /* the NCCL Alltoallv (not a kernel of ours: timed like one) */, TG_K_NUM = 10 };
int tg_profile_enable(tg_ctx* ctx, int on);
int tg_profile_get(tg_ctx* ctx, int kernel_class, float* out_total_ms, uint64_t* out_launches);
/* the individual launch durations of `kernel_class` in launch order (up to `capacity`); *out_n = how many there are */
int tg_profile_list(tg_ctx* ctx, int kernel_class, float* out_ms, size_t capacity, size_t* out_n);
/* page-locked host memory (what the BlockPool arenas should be for full PCIe bandwidth; the host shim
 * can equally cudaHostRegister its existing ByteBlocks) */
int tg_host_alloc(tg_ctx* ctx, size_t bytes, void** out_hptr);
int tg_host_free(tg_ctx* ctx, void* hptr);

/* ---- File <-> flat device buffer codec (data/file.hpp:56-283) -------------------------
 * A File of fixed-size POD items is the concatenation of its Blocks' [begin,end) with zero framing
 * (data/serialization.hpp:34-49; items may straddle Blocks).  Upload gathers Blocks into consecutive
 * device bytes (replaces File::GetReader + BlockReader::Next per item, data/block_reader.hpp:87-139);
 * download scatters consecutive device bytes into caller-allocated ByteBlocks (replaces
 * BlockWriter::Put per item, data/block_writer.hpp:208-335).  Host memory is the caller's. */
int tg_upload(tg_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int tg_download(tg_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int tg_upload_blocks(tg_ctx* ctx, void* dst_dev, const tg_block* blocks, size_t nblocks, size_t* out_bytes);
int tg_download_blocks(tg_ctx* ctx, const void* src_dev, const tg_block_mut* blocks, size_t nblocks);
/* The Block geometry BlockWriter would produce for `num_items` fixed-size items (block sizes start at
 * start_block_size, double while 2*bs < max_block_size: data/block_writer.hpp:61-67,405-420); fills
 * {bytes, first_item, num_items} per block so the host shim can build data::Block(begin=0,end,first_item,
 * num_items).  Pure host arithmetic.  Returns the number of blocks (or needed count if > capacity). */
typedef struct { uint64_t bytes, first_item, num_items; } tg_block_geom;
size_t tg_file_geometry(uint64_t num_items, uint32_t item_bytes, uint64_t start_block_size,
                        uint64_t max_block_size, tg_block_geom* out, size_t capacity);

/* ---- kernel-level entry points (parity tests, ncu captures) ------------------------------------------ */

/* Stable local sort of n items in place (result in d_items); d_tmp >= n*item_bytes scratch.  Replaces
 * SortNode::SortAndWriteToFile's sort_algorithm_(begin,end,cmp) = std::sort (api/sort.hpp:696-742,
 * :789-796) and the common::RadixSort functor hook (common/radix_sort.hpp:147-162).  Radix sort on the
 * non-constant key bytes: partition passes on the K most significant ones + one finishing pass on the runs
 * of equal prefixes, or plain LSD passes where that does not apply. */
int tg_radix_sort_local(tg_ctx* ctx, const tg_key_desc* desc, void* d_items, void* d_tmp, size_t n);

/* Sample count and host-side splitter selection: common/reservoir_sampling.hpp:270-275 (eps = 0.1,
 * api/sort.hpp:298) and FindAndSendSplitters (api/sort.hpp:337-378, LessSampleIndex :419-422).
 * samples: nsamples x (item_bytes + 8) packed (item, u64 global index) = SampleIndexPair on the wire;
 * sorted in place; writes p-1 splitters in the same packing.  Pure host arithmetic. */
uint64_t tg_sample_size(uint64_t local_items);
int tg_select_splitters(const tg_key_desc* desc, void* samples, uint64_t nsamples, uint32_t p,
                        void* out_splitters);

/* Draw the local sample on the device: min(n, tg_sample_size(n)) items at indices rng % n (the
 * OnPreOpFile path, api/sort.hpp:151-175), packed (item, global_index_base + index) into host memory. */
int tg_draw_samples(tg_ctx* ctx, const tg_key_desc* desc, const void* d_items, size_t n,
                    uint64_t global_index_base, uint64_t rng_seed, void* out_samples_host, uint64_t* out_nsamples);

/* Splitter classify + histogram + scatter.  Replaces SortNode::TransmitItems (api/sort.hpp:434-535):
 * tree descent over the p-1 splitters, ties to a splitter broken by global index
 * (EqualSampleGreaterIndex :424-426, :487-502); item i of d_in has global index global_index_base + i.
 * d_out receives the items grouped by destination worker in stable order; out_counts[p] (host) the
 * per-destination counts (what BlockWriter::Put into data_writers[b] accumulated, :507-508). */
int tg_classify_scatter(tg_ctx* ctx, const tg_key_desc* desc, const void* d_in, size_t n,
                        uint64_t global_index_base, const void* splitters_host, uint32_t p,
                        void* d_out, uint64_t* out_counts);

/* k-way merge of sorted runs laid back to back in d_runs (run r has run_items[r] items).  Replaces
 * core::MultiwayMergeTree::Next over tlx::LoserTree (core/multiway_merge.hpp:30-116,
 * extlib/tlx/tlx/container/loser_tree.hpp:54-292) as driven by SortNode::PushData (api/sort.hpp:216-271).
 * Ties are resolved by run index (the stable variant, :265-292), which is a valid unstable outcome too. */
int tg_kway_merge(tg_ctx* ctx, const tg_key_desc* desc, const void* d_runs, const uint64_t* run_items,
                  uint32_t k, void* d_out, void* d_tmp);

/* Open-addressing hash aggregate of n (key,value) items into distinct keys.  Replaces
 * ReducePrePhase::Insert / ReduceByHashPostPhase::Insert -> ReduceProbingHashTable::Insert
 * (core/reduce_probing_hash_table.hpp:190-268) + FlushAll (:484-488).  The key 0 (== Key(), the
 * reference's empty-slot sentinel, :195-218) is supported through a side accumulator.
 * d_out must hold min(n, capacity_hint) items; *out_distinct (host) = number of distinct keys.
 * Output order is table order (unspecified, as in the reference). */
int tg_hash_aggregate(tg_ctx* ctx, const tg_kv_desc* desc, const void* d_in, size_t n,
                      void* d_out, uint64_t* out_distinct);

/* Hash partition: destination worker of key = Hash128to64(0, key) % p (common/hash.hpp:64-73,
 * core/reduce_functional.hpp:60-72 with std::hash<uint64_t> = identity), items grouped by destination in
 * d_out, counts in out_counts[p] (host).  Replaces ReducePrePhaseEmitter::Emit into
 * writer_[partition_id] (core/reduce_pre_phase.hpp:57-61). */
int tg_hash_partition(tg_ctx* ctx, const tg_kv_desc* desc, const void* d_in, size_t n, uint32_t p,
                      void* d_out, uint64_t* out_counts);

/* The arithmetic of one exchange, pure host code (what every rank derives from the all-gathered p x p count matrix; exported so
 * that the N > 1 host logic is testable without GPUs): counts[src * p + dst] = items rank src holds for rank dst.  For rank `me`:
 * send_cnt[d], recv_cnt[s], recv_before[d] = items of the ranks below `me` in rank d's window (where this rank's share starts),
 * *n_recv = items this rank receives, *worst = the largest receive size of ANY rank (window growth and TG_ERR_TOO_LARGE are
 * decided on it, identically everywhere).  Replaces the per-(src,dst) block headers of the MixStream (data/multiplexer_header.hpp:36-72). */
int tg_exchange_plan(uint32_t p, uint32_t me, const uint32_t* counts, uint64_t* send_cnt, uint64_t* recv_cnt,
                     uint64_t* recv_before, uint64_t* n_recv, uint64_t* worst);

/* ---- operator-level entry points -------------------------------------------------------------------- */

/* Whole SortNode::MainOp + PushData (api/sort.hpp:537-663, :216-271) on device-resident items:
 * ExPrefixSumTotal (:541) -> samples -> splitters -> classify/scatter -> NCCL Alltoallv (replaces the
 * MixStream exchange :615-641: here the classification pass stores into the peers' exchange windows) -> local radix sort
 * of what was received.  d_in holds n_local items (it is clobbered); *out_dptr points to *out_n items inside a ctx-owned
 * workspace, the exchange window or d_in itself: valid until the next operator call on this ctx, never to be freed by the
 * caller (tg_free rejects it), to be copied (or detached with tg_output_detach after a *_file / *_dev call) before it is
 * fed to another operator.  Limits: n_local < 2^30, at most 16 ranks.  Collective: sizes are agreed on by all ranks,
 * TG_ERR_TOO_LARGE is returned by every rank or by none. */
int tg_sort(tg_ctx* ctx, const tg_key_desc* desc, void* d_in, size_t n_local, uint64_t rng_seed,
            void** out_dptr, size_t* out_n);

/* Whole ReduceNode (api/reduce_by_key.hpp:100-211): local pre-aggregation (pre phase), hash partition,
 * NCCL Alltoallv (replaces MixStream, :109-114), final aggregation (post phase).  Collective. */
int tg_reduce_by_key(tg_ctx* ctx, const tg_kv_desc* desc, const void* d_in, size_t n_local,
                     void** out_dptr, size_t* out_n);

/* DIA<pair<size_t, V>>::ReduceToIndex(key = .first, reduce function on .second, result_size, neutral_element)
 * (api/reduce_to_index.hpp:60-237; the PageRank step, examples/page_rank/page_rank.hpp:125-135): pre phase = local
 * aggregation, range partition k -> k * p / result_size (core/reduce_functional.hpp:83-149, common/math.hpp:83-120),
 * NCCL Alltoallv, dense post phase (core/reduce_by_index_post_phase.hpp:43-330).  Rank r receives the contiguous index
 * range [ceil(r*size/p), ceil((r+1)*size/p)) as a dense array of 16-byte items in *out_dptr: item i = (i, fold of the values
 * whose index is i), or *neutral_item16 where no item has that index.  *out_begin = first index of the range.  An index
 * >= result_size is TG_ERR_ARG (the reference asserts).  Collective. */
int tg_reduce_to_index(tg_ctx* ctx, const tg_kv_desc* desc, const void* d_in, size_t n_local, uint64_t result_size,
                       const void* neutral_item16, void** out_dptr, size_t* out_n, uint64_t* out_begin);
/* ... with a host File as input; the dense result is fetched with tg_fetch_output */
int tg_reduce_to_index_file(tg_ctx* ctx, const tg_kv_desc* desc, const tg_block* in_blocks, size_t n_in_blocks,
                            uint64_t result_size, const void* neutral_item16, size_t* out_items, uint64_t* out_begin);

/* The same two operators with HOST Files on both sides (the drop-in call: GpuSortNode::Execute /
 * GpuReduceNode::StopPreOp in thrill_b200/host/): gathers the input Blocks to the device, runs the
 * operator, reports the output size; tg_fetch_output then scatters the result into caller-allocated
 * ByteBlocks and releases it. */
int tg_sort_file(tg_ctx* ctx, const tg_key_desc* desc, const tg_block* in_blocks, size_t n_in_blocks,
                 uint64_t rng_seed, size_t* out_items);
int tg_reduce_file(tg_ctx* ctx, const tg_kv_desc* desc, const tg_block* in_blocks, size_t n_in_blocks,
                   size_t* out_items);
int tg_fetch_output(tg_ctx* ctx, const tg_block_mut* out_blocks, size_t n_out_blocks);

/* ---- device-resident Files: GPU node -> GPU node without the PCIe round trip ----------------------
 * The reference hands a node's result to its children as a data::File (DIANode::PushFile -> child->OnPreOpFile,
 * api/dia_node.hpp:156-180; api/sort.hpp:151-175 takes it whole).  When the child is another GPU node the File need not
 * exist on the host at all: the parent keeps its result as a device File: flat items in HBM, the same layout as the
 * concatenated Blocks, and the child's operator reads it there.  A host File is materialised only when a child that is
 * not a GPU node asks for it (tg_dev_file_fetch: the lazy D2H of a PinnedBlock, data/block.hpp:116). */
typedef struct {
    void* dptr;                /* HBM buffer owned by the handle (tg_dev_file_free) */
    uint64_t items;
    uint32_t item_bytes;
    uint32_t reserved;
} tg_dev_file;
/* take the result of the last *_file / *_dev operator as a device File instead of fetching it (tg_fetch_output) */
int tg_output_detach(tg_ctx* ctx, tg_dev_file* out);
int tg_dev_file_fetch(tg_ctx* ctx, const tg_dev_file* f, const tg_block_mut* out_blocks, size_t n_out_blocks);
int tg_dev_file_free(tg_ctx* ctx, tg_dev_file* f);
/* the operators with a device File as input (the handle is left intact: a DIA may have several children) */
int tg_sort_dev(tg_ctx* ctx, const tg_key_desc* desc, const tg_dev_file* in, uint64_t rng_seed, size_t* out_items);
int tg_reduce_dev(tg_ctx* ctx, const tg_kv_desc* desc, const tg_dev_file* in, size_t* out_items);
int tg_reduce_to_index_dev(tg_ctx* ctx, const tg_kv_desc* desc, const tg_dev_file* in, uint64_t result_size,
                           const void* neutral_item16, size_t* out_items, uint64_t* out_begin);
/* bytes this ctx has moved over PCIe through the File codec since tg_init (tests: a GPU -> GPU chain moves none in between) */
int tg_transfer_bytes(const tg_ctx* ctx, uint64_t* out_h2d, uint64_t* out_d2h);

/* ---- synthetic inputs of SURVEY.md §8(d), generated on the device (bench / tests support) ------------ */
int tg_gen_sort_uniform(tg_ctx* ctx, void* d_out, uint64_t begin, uint64_t n, uint64_t seed);
int tg_gen_reduce_uniform(tg_ctx* ctx, void* d_out, uint64_t begin, uint64_t n, uint64_t seed,
                          uint64_t universe, int exact);
/* d_cdf: universe doubles (cumulative Zipf table built on the host exactly as
 * common/zipf_distribution.hpp:119-140 and uploaded once) */
int tg_gen_sort_zipf(tg_ctx* ctx, void* d_out, uint64_t begin, uint64_t n, uint64_t seed,
                     const void* d_cdf, uint64_t universe);
int tg_gen_reduce_zipf(tg_ctx* ctx, void* d_out, uint64_t begin, uint64_t n, uint64_t seed,
                       const void* d_cdf, uint64_t universe, int exact);
int tg_gen_records(tg_ctx* ctx, void* d_out, uint64_t begin, uint64_t n, uint64_t seed);
/* order-independent 64-bit checksum of n items (sum and xor of a per-item hash) and sortedness check;
 * used by the full-size parity properties (sortedness + multiset preservation) */
int tg_checksum(tg_ctx* ctx, const void* d_items, size_t n, uint32_t item_bytes, uint64_t out_sum_xor[2]);
int tg_is_sorted(tg_ctx* ctx, const tg_key_desc* desc, const void* d_items, size_t n, uint64_t* out_violations);

#ifdef __cplusplus
}
#endif
#endif /* THRILL_GPU_H */
