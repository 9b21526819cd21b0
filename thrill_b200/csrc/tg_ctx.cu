// tg_ctx.cu — context, device memory, File<->device codec, timers (C ABI: include/thrill_gpu.h)
#include <stdarg.h>

#include "tg_common.cuh"

namespace tgp { void xwin_release(tg_ctx* ctx); }

int tg_set_error(tg_ctx* ctx, int status, const char* fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return status;
}

int tg_pinned_list(tg_ctx* ctx, int which, size_t bytes, void** out) {
    if (ctx->pinned_list_bytes[which] < bytes) {
        if (ctx->pinned_list[which]) {
            TG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            TG_CUDA(ctx, cudaFreeHost(ctx->pinned_list[which]));
            ctx->pinned_list[which] = nullptr;
            ctx->pinned_list_bytes[which] = 0;
        }
        size_t want = bytes + (bytes >> 2) + 4096;
        TG_CUDA(ctx, cudaMallocHost(&ctx->pinned_list[which], want));
        ctx->pinned_list_bytes[which] = want;
    }
    *out = ctx->pinned_list[which];
    return TG_OK;
}

int tg_ws_get(tg_ctx* ctx, int slot, size_t bytes, void** out) {
    if (bytes == 0) bytes = 256;
    if (ctx->ws_bytes[slot] < bytes) {
        if (ctx->ws[slot]) {
            TG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            TG_CUDA(ctx, cudaFree(ctx->ws[slot]));
            ctx->ws[slot] = nullptr;
            ctx->ws_bytes[slot] = 0;
        }
        size_t want = bytes + (bytes >> 3) + 4096;      // head-room so slowly growing sizes don't realloc
        cudaError_t e = cudaMalloc(&ctx->ws[slot], want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&ctx->ws[slot], want);
        }
        if (e != cudaSuccess)
            return tg_set_error(ctx, TG_ERR_OOM, "workspace slot %d: cudaMalloc(%zu) -> %s", slot, want, cudaGetErrorString(e));
        ctx->ws_bytes[slot] = want;
    }
    *out = ctx->ws[slot];
    return TG_OK;
}

int tg_prof_begin(tg_ctx* ctx, int cls) {
    tg_ctx::ProfEv e;
    e.cls = cls;
    for (cudaEvent_t* ev : { &e.a, &e.b }) {
        if (!ctx->prof_pool.empty()) { *ev = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
        else if (cudaEventCreate(ev) != cudaSuccess) return -1;
    }
    cudaEventRecord(e.a, ctx->stream);
    ctx->prof_events.push_back(e);
    return (int)ctx->prof_events.size() - 1;
}
void tg_prof_end(tg_ctx* ctx, int slot) { cudaEventRecord(ctx->prof_events[slot].b, ctx->stream); }

extern "C" {

int tg_profile_enable(tg_ctx* ctx, int on) {
    TG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& e : ctx->prof_events) { ctx->prof_pool.push_back(e.a); ctx->prof_pool.push_back(e.b); }
    ctx->prof_events.clear();
    ctx->profile = on != 0;
    return TG_OK;
}

int tg_profile_get(tg_ctx* ctx, int kernel_class, float* out_total_ms, uint64_t* out_launches) {
    TG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float tot = 0.f;
    uint64_t cnt = 0;
    for (auto& e : ctx->prof_events)
        if (e.cls == kernel_class) {
            float ms = 0.f;
            TG_CUDA(ctx, cudaEventElapsedTime(&ms, e.a, e.b));
            tot += ms;
            ++cnt;
        }
    if (out_total_ms) *out_total_ms = tot;
    if (out_launches) *out_launches = cnt;
    return TG_OK;
}

int tg_profile_list(tg_ctx* ctx, int kernel_class, float* out_ms, size_t capacity, size_t* out_n) {
    TG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    size_t cnt = 0;
    for (auto& e : ctx->prof_events)
        if (e.cls == kernel_class) {
            float ms = 0.f;
            TG_CUDA(ctx, cudaEventElapsedTime(&ms, e.a, e.b));
            if (out_ms && cnt < capacity) out_ms[cnt] = ms;
            ++cnt;
        }
    if (out_n) *out_n = cnt;
    return TG_OK;
}

int tg_host_alloc(tg_ctx* ctx, size_t bytes, void** out_hptr) {
    if (!out_hptr) return TG_ERR_ARG;
    cudaError_t e = cudaMallocHost(out_hptr, bytes ? bytes : 16);
    if (e != cudaSuccess) { cudaGetLastError(); return tg_set_error(ctx, TG_ERR_OOM, "cudaMallocHost(%zu) -> %s", bytes, cudaGetErrorString(e)); }
    return TG_OK;
}
int tg_host_free(tg_ctx* ctx, void* hptr) {
    if (hptr) TG_CUDA(ctx, cudaFreeHost(hptr));
    return TG_OK;
}

int tg_version(void) { return 200; }

int tg_device_count(void) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < ndev; ++d) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) ++ok;
    }
    return ok == ndev ? ndev : 0;
}

const char* tg_strerror(int status) {
    switch (status) {
    case TG_OK: return "ok";
    case TG_ERR_CUDA: return "CUDA error";
    case TG_ERR_NCCL: return "NCCL error";
    case TG_ERR_ARG: return "bad argument / unsupported descriptor";
    case TG_ERR_TOO_LARGE: return "too many items for one call";
    case TG_ERR_NO_DEVICE: return "no sm_100 CUDA device (there is no CPU fallback)";
    case TG_ERR_OOM: return "out of device memory";
    default: return "unknown status";
    }
}

const char* tg_last_error(const tg_ctx* ctx) { return ctx ? ctx->err : "no context"; }

int tg_get_unique_id(void* out128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return TG_ERR_NCCL;
    memcpy(out128, &id, sizeof(id));
    return TG_OK;
}

int tg_init(int device, int rank, int nranks, const void* unique_id128, tg_ctx** out_ctx) {
    if (!out_ctx || nranks < 1 || rank < 0 || rank >= nranks) return TG_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return TG_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) return TG_ERR_ARG;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return TG_ERR_CUDA;
    if (prop.major != 10) return TG_ERR_NO_DEVICE;      // kernels are built for sm_100a only
    tg_ctx* ctx = new tg_ctx();
    ctx->device = device; ctx->rank = rank; ctx->nranks = nranks;
    ctx->sm_count = prop.multiProcessorCount;
    *out_ctx = ctx;
    TG_CUDA(ctx, cudaSetDevice(device));
    TG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    TG_CUDA(ctx, cudaEventCreate(&ctx->ev_start));
    TG_CUDA(ctx, cudaEventCreate(&ctx->ev_stop));
    ctx->pinned_bytes = 1 << 20;
    TG_CUDA(ctx, cudaMallocHost(&ctx->pinned, ctx->pinned_bytes));
    if (nranks > 1) {
        if (!unique_id128) return tg_set_error(ctx, TG_ERR_ARG, "nranks > 1 needs a unique id");
        ncclUniqueId id;
        memcpy(&id, unique_id128, sizeof(id));
        TG_NCCL(ctx, ncclCommInitRank(&ctx->comm, nranks, id, rank));
    }
    return TG_OK;
}

int tg_shutdown(tg_ctx* ctx) {
    if (!ctx) return TG_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    tgp::xwin_release(ctx);
    if (ctx->comm) ncclCommDestroy(ctx->comm);
    for (auto& kv : ctx->allocs) cudaFree(kv.first);
    for (int i = 0; i < TG_NUM_WS; ++i)
        if (ctx->ws[i]) cudaFree(ctx->ws[i]);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (int i = 0; i < 2; ++i) if (ctx->pinned_list[i]) cudaFreeHost(ctx->pinned_list[i]);
    for (auto& e : ctx->prof_events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto& e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return TG_OK;
}

int tg_rank(const tg_ctx* ctx) { return ctx->rank; }
int tg_nranks(const tg_ctx* ctx) { return ctx->nranks; }
void* tg_stream(const tg_ctx* ctx) { return (void*)ctx->stream; }
uint64_t tg_launch_count(const tg_ctx* ctx) { return ctx->launches; }
uint64_t tg_prefix_sort_fallbacks(const tg_ctx* ctx) { return ctx->prefix_sort_fallbacks; }
uint64_t tg_hot_records(const tg_ctx* ctx) { return ctx->hot_records; }

int tg_sync(tg_ctx* ctx) {
    TG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TG_OK;
}

int tg_barrier(tg_ctx* ctx) {
    if (ctx->nranks > 1) {
        void* d;
        TG_TRY(tg_ws_get(ctx, WS_MISC, 4096, &d));
        TG_CUDA(ctx, cudaMemsetAsync(d, 0, 8, ctx->stream));
        TG_NCCL(ctx, ncclAllReduce(d, d, 1, ncclInt32, ncclSum, ctx->comm, ctx->stream));
    }
    return tg_sync(ctx);
}

int tg_alloc(tg_ctx* ctx, size_t bytes, void** out_dptr) {
    if (!out_dptr) return TG_ERR_ARG;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 256);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return tg_set_error(ctx, TG_ERR_OOM, "cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
    }
    ctx->allocs[p] = bytes;
    *out_dptr = p;
    return TG_OK;
}

int tg_free(tg_ctx* ctx, void* dptr) {
    if (!dptr) return TG_OK;
    auto it = ctx->allocs.find(dptr);
    if (it == ctx->allocs.end()) return tg_set_error(ctx, TG_ERR_ARG, "tg_free of a pointer this ctx does not own");
    TG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    TG_CUDA(ctx, cudaFree(dptr));
    ctx->allocs.erase(it);
    return TG_OK;
}

int tg_timer_start(tg_ctx* ctx) {
    TG_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    return TG_OK;
}

int tg_timer_stop(tg_ctx* ctx, float* out_ms) {
    TG_CUDA(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
    TG_CUDA(ctx, cudaEventSynchronize(ctx->ev_stop));
    TG_CUDA(ctx, cudaEventElapsedTime(out_ms, ctx->ev_start, ctx->ev_stop));
    return TG_OK;
}

int tg_upload(tg_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    if (bytes) TG_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->bytes_h2d += bytes;
    return TG_OK;
}

int tg_download(tg_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    if (bytes) TG_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->bytes_d2h += bytes;
    return TG_OK;
}

// Blocks that happen to be adjacent in host memory (one arena, a memory-mapped file, ReadBinary's ranges) go as one copy:
// fewer, longer DMA transfers
int tg_upload_blocks(tg_ctx* ctx, void* dst_dev, const tg_block* blocks, size_t nblocks, size_t* out_bytes) {
    size_t off = 0, i = 0;
    while (i < nblocks) {
        const char* base = (const char*)blocks[i].data;
        size_t len = blocks[i].bytes, k = i + 1;
        while (k < nblocks && (blocks[k].bytes == 0 || (const char*)blocks[k].data == base + len)) len += blocks[k++].bytes;
        if (len) TG_CUDA(ctx, cudaMemcpyAsync((char*)dst_dev + off, base, len, cudaMemcpyHostToDevice, ctx->stream));
        off += len;
        i = k;
    }
    ctx->bytes_h2d += off;
    if (out_bytes) *out_bytes = off;
    return TG_OK;
}

int tg_download_blocks(tg_ctx* ctx, const void* src_dev, const tg_block_mut* blocks, size_t nblocks) {
    size_t off = 0, i = 0;
    while (i < nblocks) {
        char* base = (char*)blocks[i].data;
        size_t len = blocks[i].bytes, k = i + 1;
        while (k < nblocks && (blocks[k].bytes == 0 || (char*)blocks[k].data == base + len)) len += blocks[k++].bytes;
        if (len) TG_CUDA(ctx, cudaMemcpyAsync(base, (const char*)src_dev + off, len, cudaMemcpyDeviceToHost, ctx->stream));
        off += len;
        i = k;
    }
    ctx->bytes_d2h += off;
    return TG_OK;
}

int tg_transfer_bytes(const tg_ctx* ctx, uint64_t* out_h2d, uint64_t* out_d2h) {
    if (!ctx) return TG_ERR_ARG;
    if (out_h2d) *out_h2d = ctx->bytes_h2d;
    if (out_d2h) *out_d2h = ctx->bytes_d2h;
    return TG_OK;
}

// ---- device-resident Files (include/thrill_gpu.h) ----------------------------------------------------------------------
int tg_output_detach(tg_ctx* ctx, tg_dev_file* out) {
    if (!ctx || !out) return TG_ERR_ARG;
    TG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = ctx->out_items * (size_t)ctx->out_item_bytes;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 256);
    if (e != cudaSuccess) { cudaGetLastError(); return tg_set_error(ctx, TG_ERR_OOM, "device File: cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e)); }
    // (the result lies in a workspace slot, the exchange window or the operator's input buffer, all of which the next operator
    // re-uses: one device-to-device copy, ~0.1 ms per GB, makes the File independent of them)
    if (bytes) TG_CUDA(ctx, cudaMemcpyAsync(p, ctx->out_ptr, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    TG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out->dptr = p; out->items = ctx->out_items; out->item_bytes = ctx->out_item_bytes; out->reserved = 0;
    ctx->out_ptr = nullptr; ctx->out_items = 0;
    return TG_OK;
}

int tg_dev_file_fetch(tg_ctx* ctx, const tg_dev_file* f, const tg_block_mut* out_blocks, size_t n_out_blocks) {
    if (!ctx || !f) return TG_ERR_ARG;
    TG_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t bytes = 0;
    for (size_t i = 0; i < n_out_blocks; ++i) bytes += out_blocks[i].bytes;
    if (bytes != f->items * (size_t)f->item_bytes)
        return tg_set_error(ctx, TG_ERR_ARG, "dev_file_fetch: blocks hold %zu bytes, the File has %zu", bytes, (size_t)(f->items * f->item_bytes));
    if (bytes) TG_TRY(tg_download_blocks(ctx, f->dptr, out_blocks, n_out_blocks));
    TG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TG_OK;
}

int tg_dev_file_free(tg_ctx* ctx, tg_dev_file* f) {
    if (!f || !f->dptr) return TG_OK;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    cudaFree(f->dptr);
    f->dptr = nullptr; f->items = 0;
    return TG_OK;
}

// the operator input of a device File: copied into the ctx's input workspace (operators clobber their input)
static int stage_dev_file(tg_ctx* ctx, const tg_dev_file* in, uint32_t item_bytes, void** d_in) {
    if (!in || in->item_bytes != item_bytes) return tg_set_error(ctx, TG_ERR_ARG, "device File: item size %u, operator expects %u", in ? in->item_bytes : 0, item_bytes);
    const size_t bytes = in->items * (size_t)in->item_bytes;
    TG_TRY(tg_ws_get(ctx, WS_IN, bytes + 16, d_in));
    if (bytes) TG_CUDA(ctx, cudaMemcpyAsync(*d_in, in->dptr, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return TG_OK;
}

int tg_sort_dev(tg_ctx* ctx, const tg_key_desc* desc, const tg_dev_file* in, uint64_t rng_seed, size_t* out_items) {
    if (!ctx || !desc || !out_items) return TG_ERR_ARG;
    TG_CUDA(ctx, cudaSetDevice(ctx->device));
    void* d_in;
    TG_TRY(stage_dev_file(ctx, in, desc->item_bytes, &d_in));
    void* out = nullptr;
    size_t n_out = 0;
    TG_TRY(tg_sort(ctx, desc, d_in, in->items, rng_seed, &out, &n_out));
    ctx->out_ptr = out; ctx->out_items = n_out; ctx->out_item_bytes = desc->item_bytes;
    *out_items = n_out;
    return TG_OK;
}

int tg_reduce_dev(tg_ctx* ctx, const tg_kv_desc* desc, const tg_dev_file* in, size_t* out_items) {
    if (!ctx || !desc || !out_items) return TG_ERR_ARG;
    TG_CUDA(ctx, cudaSetDevice(ctx->device));
    void* d_in;
    TG_TRY(stage_dev_file(ctx, in, 16, &d_in));
    void* out = nullptr;
    size_t n_out = 0;
    TG_TRY(tg_reduce_by_key(ctx, desc, d_in, in->items, &out, &n_out));
    ctx->out_ptr = out; ctx->out_items = n_out; ctx->out_item_bytes = 16;
    *out_items = n_out;
    return TG_OK;
}

int tg_reduce_to_index_dev(tg_ctx* ctx, const tg_kv_desc* desc, const tg_dev_file* in, uint64_t result_size,
                           const void* neutral_item16, size_t* out_items, uint64_t* out_begin) {
    if (!ctx || !desc || !out_items || !out_begin) return TG_ERR_ARG;
    TG_CUDA(ctx, cudaSetDevice(ctx->device));
    void* d_in;
    TG_TRY(stage_dev_file(ctx, in, 16, &d_in));
    void* out = nullptr;
    size_t n_out = 0;
    TG_TRY(tg_reduce_to_index(ctx, desc, d_in, in->items, result_size, neutral_item16, &out, &n_out, out_begin));
    ctx->out_ptr = out; ctx->out_items = n_out; ctx->out_item_bytes = 16;
    *out_items = n_out;
    return TG_OK;
}

// data/block_writer.hpp:61-67 (initial block size), :405-420 (doubling), :183-203 / :339-372 (item starts,
// straddling Append) for fixed-size items
size_t tg_file_geometry(uint64_t num_items, uint32_t item_bytes, uint64_t start_block_size,
                        uint64_t max_block_size, tg_block_geom* out, size_t capacity) {
    uint64_t total = num_items * (uint64_t)item_bytes;
    uint64_t bs = start_block_size < max_block_size ? start_block_size : max_block_size;
    uint64_t pos = 0;
    size_t nb = 0;
    while (pos < total) {
        uint64_t len = total - pos < bs ? total - pos : bs;
        // first item starting at/after pos, items starting inside [pos, pos+len)
        uint64_t first_idx = (pos + item_bytes - 1) / item_bytes;
        uint64_t end_idx = (pos + len + item_bytes - 1) / item_bytes;     // first item starting at/after block end
        if (end_idx > num_items) end_idx = num_items;
        if (nb < capacity) {
            out[nb].bytes = len;
            out[nb].num_items = end_idx > first_idx ? end_idx - first_idx : 0;
            out[nb].first_item = out[nb].num_items ? first_idx * item_bytes - pos : 0;
        }
        ++nb;
        pos += len;
        if (2 * bs < max_block_size) bs *= 2;
    }
    return nb;
}

}  // extern "C"
