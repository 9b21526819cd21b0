"""Host-side mirror of the reference's DIA<T> operator interface for the Sort / ReduceByKey path.

Mirrors (same names, argument meaning and error behaviour) the slice of thrill/api that drives the hot path:
    api::Run / Context                     thrill/api/context.hpp:243-377  (my_rank, num_workers)
    api::Generate(ctx, n, fn)              thrill/api/generate.hpp         (even range split over workers)
    DIA<T>::Sort(cmp) / SortStable(cmp)    thrill/api/sort.hpp:800-937
    DIA<T>::ReducePair(reduce_fn)          thrill/api/reduce_by_key.hpp:410-449
    DIA<T>::ReduceByKey(key_ex, reduce_fn) thrill/api/reduce_by_key.hpp:312-363
    DIA<T>::Size / AllGather / Gather      thrill/api/size.hpp, all_gather.hpp, gather.hpp
A DIA here holds its local shard as a host numpy array — the stand-in for a data::File whose Blocks are
1 MiB ByteBlocks (data/byte_block.cpp:23-24, data/block_writer.hpp:405-420).  Operators hand the Blocks to the
C ABI (tg_sort_file / tg_reduce_file / tg_fetch_output): exactly what thrill_b200/host/gpu_sort_node.hpp does
from C++.  Like the C++ host shim, only a closed set of functors is recognised (SURVEY.md §7 "UDFs");
anything else raises — there is no CPU fallback.

torch.distributed is used only as the control plane between the one-process-per-GPU workers (what Thrill's
net::FlowControlChannel does): broadcasting the NCCL id and gathering results for AllGather().
"""
import ctypes as C
import os

import numpy as np

from . import capi

KV = np.dtype([("key", "<u8"), ("val", "<u8")])          # std::pair<uint64_t, 8-byte value>, member-wise
BLOCK_BYTES = 1 << 20                                      # largest ByteBlock BlockWriter produces by default


class _Functor(object):
    def __init__(self, name, code=None):
        self.name, self.code = name, code

    def __repr__(self):
        return "<thrill_b200 functor %s>" % self.name


# recognised comparators (std::less<T> / std::greater<T>) and reduce functions
Less = _Functor("std::less")
Greater = _Functor("std::greater")
PlusDouble = _Functor("std::plus<double>", capi.OP_SUM_F64)
PlusU64 = _Functor("std::plus<uint64_t>", capi.OP_SUM_U64)
MinU64 = _Functor("min<uint64_t>", capi.OP_MIN_U64)
MaxU64 = _Functor("max<uint64_t>", capi.OP_MAX_U64)
MinDouble = _Functor("min<double>", capi.OP_MIN_F64)
MaxDouble = _Functor("max<double>", capi.OP_MAX_F64)
First = _Functor("first", capi.OP_FIRST)
KeyIsFirst = _Functor("pair.first")                       # the key extractor ReducePair builds (:444-449)


def bind_to_gpu_numa_node(device):
    """Run this worker process on the CPUs next to its GPU (nvidia-smi topo -m "CPU Affinity"): the pinned staging
    buffers it allocates afterwards, and the copies it issues, stay on the GPU's NUMA node.  What a Thrill launcher does
    with numactl per worker; without it eight workers share one node's memory controllers for their PCIe traffic."""
    try:
        import pynvml
        pynvml.nvmlInit()
        # NVML counts physical GPUs; CUDA ordinals go through CUDA_VISIBLE_DEVICES
        vis = [v.strip() for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
        if vis and device < len(vis):
            h = pynvml.nvmlDeviceGetHandleByUUID(vis[device]) if vis[device].startswith("GPU-") else \
                pynvml.nvmlDeviceGetHandleByIndex(int(vis[device]))
        else:
            h = pynvml.nvmlDeviceGetHandleByIndex(device)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:          # noqa: BLE001  (no NVML, no permission: run unbound)
        pass
    return None


class Context(object):
    """One worker = one GPU (api/context.hpp:243-245)."""

    def __init__(self, rank=0, nranks=1, device=None, unique_id=None, rng_seed=None):
        self._rank, self._n = rank, nranks
        self.tg = capi.Ctx(device=rank if device is None else device, rank=rank, nranks=nranks, unique_id=unique_id)
        # the reference seeds each worker's rng from std::random_device (api/context.cpp:1190)
        self.rng_seed = int.from_bytes(os.urandom(8), "little") if rng_seed is None else rng_seed
        self._op_counter = 0

    @classmethod
    def from_env(cls, rng_seed=None):
        """one process per GPU under torchrun: RANK / LOCAL_RANK / WORLD_SIZE, id broadcast over torch.distributed"""
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", str(rank)))
        uid = None
        if world > 1 and os.environ.get("TG_NUMA_BIND", "1") != "0":
            bind_to_gpu_numa_node(local)
        if world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                dist.init_process_group("gloo")
            buf = [None]
            if rank == 0:
                raw = C.create_string_buffer(128)
                capi.check(capi.lib().tg_get_unique_id(raw))
                buf[0] = raw.raw
            dist.broadcast_object_list(buf, src=0)
            uid = buf[0]
        return cls(rank, world, device=local, unique_id=uid, rng_seed=rng_seed)

    def my_rank(self):
        return self._rank

    def num_workers(self):
        return self._n

    def close(self):
        self.tg.close()

    def _next_seed(self):
        self._op_counter += 1
        return (self.rng_seed + 0x9E3779B97F4A7C15 * self._op_counter) % (1 << 64)


def _local_range(n, p, r):
    """common::CalculateLocalRange (thrill/common/math.hpp): [r*n/p, (r+1)*n/p) in double arithmetic"""
    per = float(n) / float(p)
    lo = int(np.ceil(r * per))
    hi = min(n, int(np.ceil((r + 1) * per)))
    return lo, hi


def Generate(ctx, size, generator_function, dtype=np.uint64):
    """api::Generate: item i = generator_function(i) for the worker's index range; the generator here is
    vectorised: it receives a numpy array of global indices and returns the items."""
    lo, hi = _local_range(size, ctx.num_workers(), ctx.my_rank())
    idx = np.arange(lo, hi, dtype=np.uint64)
    items = np.ascontiguousarray(generator_function(idx))
    return DIA(ctx, items if dtype is None else items.astype(dtype, copy=False))


def _item_bytes(items):
    return items.dtype.itemsize if items.ndim == 1 else items.shape[1] * items.dtype.itemsize


class DIA(object):
    def __init__(self, ctx, items):
        self.ctx = ctx
        self.items = np.ascontiguousarray(items)

    # ---- Block views of the local File ----------------------------------------------------------------
    def _blocks(self, arr, mutable=False):
        raw = arr.view(np.uint8).reshape(-1)
        n = len(raw)
        nb = (n + BLOCK_BYTES - 1) // BLOCK_BYTES
        blocks = (capi.Block * max(nb, 1))()
        for i in range(nb):
            lo = i * BLOCK_BYTES
            blocks[i].data = raw.ctypes.data + lo
            blocks[i].bytes = min(n, lo + BLOCK_BYTES) - lo
        return blocks, nb

    def _fetch(self, n_items, dtype, item_bytes, pinned_out=None):
        nbytes = n_items * item_bytes
        out = pinned_out if pinned_out is not None else np.empty(nbytes, dtype=np.uint8)
        out = out[:nbytes]
        blocks, nb = self._blocks(out)
        self.ctx.tg.ck(self.ctx.tg.L.tg_fetch_output(self.ctx.tg.h, blocks, nb))
        if dtype is None:
            return out.reshape(n_items, item_bytes)
        return out.view(dtype)

    # ---- DIA<T>::Sort ----------------------------------------------------------------------------------
    def _key_desc(self, compare_function):
        if compare_function not in (None, Less, Greater):
            raise capi.ThrillGpuError("Sort: comparator %r is not one the GPU path recognises "
                                      "(std::less / std::greater on the key)" % (compare_function,))
        desc = 1 if compare_function is Greater else 0
        it = self.items
        if it.ndim == 1 and it.dtype == np.uint64:
            return capi.KeyDesc(8, 0, 8, capi.KEY_UINT_LE, desc, 0), np.uint64
        if it.ndim == 1 and it.dtype == KV:
            return capi.KeyDesc(16, 0, 8, capi.KEY_UINT_LE, desc, 0), KV
        if it.ndim == 2 and it.dtype == np.uint8 and it.shape[1] == 100 and not desc:
            # TeraSort Record{uint8 key[10]; uint8 value[90]}, operator< = lexicographic on the key
            # (examples/terasort/terasort.cpp:31-42)
            return capi.KeyDesc(100, 0, 10, capi.KEY_BYTES_BE, 0, 0), None
        raise capi.ThrillGpuError("Sort: item type %r/%r is not supported by the GPU path" % (it.dtype, it.shape))

    def Sort(self, compare_function=None, _pinned_out=None):
        desc, dtype = self._key_desc(compare_function)
        blocks, nb = self._blocks(self.items)
        n_out = C.c_size_t()
        tg = self.ctx.tg
        tg.ck(tg.L.tg_sort_file(tg.h, C.byref(desc), blocks, nb, self.ctx._next_seed(), C.byref(n_out)))
        return DIA(self.ctx, self._fetch(n_out.value, dtype, desc.item_bytes, _pinned_out))

    def SortStable(self, compare_function=None):
        return self.Sort(compare_function)         # the GPU path is stable by construction

    # ---- DIA<T>::ReducePair / ReduceByKey --------------------------------------------------------------
    def ReducePair(self, reduce_function, _pinned_out=None):
        if not isinstance(reduce_function, _Functor) or reduce_function.code is None:
            raise capi.ThrillGpuError("ReducePair: reduce function %r is not one the GPU path recognises" % (reduce_function,))
        if not (self.items.ndim == 1 and self.items.dtype == KV):
            raise capi.ThrillGpuError("ReducePair: items must be pair<uint64_t, 8-byte value>")
        desc = capi.KVDesc(16, reduce_function.code)
        blocks, nb = self._blocks(self.items)
        n_out = C.c_size_t()
        tg = self.ctx.tg
        tg.ck(tg.L.tg_reduce_file(tg.h, C.byref(desc), blocks, nb, C.byref(n_out)))
        return DIA(self.ctx, self._fetch(n_out.value, KV, 16, _pinned_out))

    # ---- DIA<T>::ReduceToIndex (api/reduce_to_index.hpp:60-237) -------------------------------------------------------
    def ReduceToIndex(self, key_extractor, reduce_function, size, neutral_element=(0, 0), _pinned_out=None):
        """items: pair<uint64_t index, 8-byte value>; result: this worker's contiguous slice of the dense array of `size`
        16-byte items ((i, reduced value), or the neutral element where no item has index i); .index_begin = first index"""
        if key_extractor is not KeyIsFirst:
            raise capi.ThrillGpuError("ReduceToIndex: only the pair.first key extractor is recognised by the GPU path")
        if not isinstance(reduce_function, _Functor) or reduce_function.code is None:
            raise capi.ThrillGpuError("ReduceToIndex: reduce function %r is not one the GPU path recognises" % (reduce_function,))
        if not (self.items.ndim == 1 and self.items.dtype == KV):
            raise capi.ThrillGpuError("ReduceToIndex: items must be pair<uint64_t, 8-byte value>")
        desc = capi.KVDesc(16, reduce_function.code)
        blocks, nb = self._blocks(self.items)
        neutral = np.zeros(1, dtype=KV)
        neutral["key"], neutral["val"] = int(neutral_element[0]), int(neutral_element[1])
        n_out, begin = C.c_size_t(), C.c_uint64()
        tg = self.ctx.tg
        tg.ck(tg.L.tg_reduce_to_index_file(tg.h, C.byref(desc), blocks, nb, int(size), neutral.ctypes.data,
                                           C.byref(n_out), C.byref(begin)))
        out = DIA(self.ctx, self._fetch(n_out.value, KV, 16, _pinned_out))
        out.index_begin = int(begin.value)
        return out

    def ReduceByKey(self, key_extractor, reduce_function):
        if key_extractor is not KeyIsFirst:
            raise capi.ThrillGpuError("ReduceByKey: only the pair.first key extractor is recognised by the GPU path")
        return self.ReducePair(reduce_function)

    # ---- actions ---------------------------------------------------------------------------------------
    def Size(self):
        n = len(self.items)
        if self.ctx.num_workers() > 1:
            import torch
            import torch.distributed as dist
            t = torch.tensor([n], dtype=torch.int64)
            dist.all_reduce(t)
            n = int(t.item())
        return n

    def AllGather(self):
        if self.ctx.num_workers() == 1:
            return self.items
        import torch.distributed as dist
        parts = [None] * self.ctx.num_workers()
        dist.all_gather_object(parts, self.items)
        return np.concatenate(parts)

    def Gather(self, root=0):
        allv = self.AllGather()
        return allv if self.ctx.my_rank() == root else allv[:0]
