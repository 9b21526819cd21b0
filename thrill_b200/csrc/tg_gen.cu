// tg_gen.cu — synthetic inputs of SURVEY.md §8(d) generated on the device, plus the order-independent
// checksum and sortedness probes used by the full-size parity properties (bench / tests support).
#include "tg_common.cuh"

namespace {

__device__ __forceinline__ double u01_dev(u64 r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

__device__ __forceinline__ u64 gen_val_dev(u64 i, u64 seed, int exact) {
    u64 r = splitmix64_dev(i + seed + (1ull << 40));
    if (exact == 2) return r % 1024;
    double v = exact ? (double)(r % 1024) : u01_dev(r);
    return (u64)__double_as_longlong(v);
}

// smallest k with cdf[k-1] > u (std::upper_bound), clamped to universe
__device__ __forceinline__ u64 zipf_rank_dev(const double* __restrict__ cdf, u64 universe, double u) {
    u64 lo = 0, hi = universe;
    while (lo < hi) {
        u64 mid = lo + (hi - lo) / 2;
        if (!(u < cdf[mid])) lo = mid + 1; else hi = mid;
    }
    if (lo >= universe) lo = universe - 1;
    return lo + 1;
}

__global__ void gen_sort_uniform_kernel(u64* out, u64 begin, u64 n, u64 seed) {
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = splitmix64_dev(begin + i + seed);
}
__global__ void gen_reduce_uniform_kernel(ulonglong2* out, u64 begin, u64 n, u64 seed, u64 universe, int exact) {
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        u64 i = begin + j;
        out[j] = make_ulonglong2(1 + splitmix64_dev(i + seed) % universe, gen_val_dev(i, seed, exact));
    }
}
__global__ void gen_sort_zipf_kernel(u64* out, u64 begin, u64 n, u64 seed, const double* cdf, u64 universe) {
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        out[j] = zipf_rank_dev(cdf, universe, u01_dev(splitmix64_dev(begin + j + seed)));
}
__global__ void gen_reduce_zipf_kernel(ulonglong2* out, u64 begin, u64 n, u64 seed, const double* cdf, u64 universe, int exact) {
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        u64 i = begin + j;
        out[j] = make_ulonglong2(zipf_rank_dev(cdf, universe, u01_dev(splitmix64_dev(i + seed))), gen_val_dev(i, seed, exact));
    }
}
// Record{uint8 key[10]; uint8 value[90]} (examples/terasort/terasort.cpp:31-42); one thread per record
__global__ void gen_records_kernel(unsigned char* out, u64 begin, u64 n, u64 seed) {
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        u64 i = begin + j;
        unsigned char* r = out + 100 * j;
        u64 a = splitmix64_dev(2 * i + seed), b = splitmix64_dev(2 * i + 1 + seed);
        for (int k = 0; k < 8; ++k) r[k] = (unsigned char)(a >> (8 * k));
        r[8] = (unsigned char)b; r[9] = (unsigned char)(b >> 8);
        for (int w = 0; w < 12; ++w) {
            u64 v = splitmix64_dev(i * 12 + w + (seed << 32));
            int len = (w == 11) ? 2 : 8;
            for (int k = 0; k < len; ++k) r[10 + 8 * w + k] = (unsigned char)(v >> (8 * k));
        }
    }
}

// order-independent multiset checksum: sum and xor of a per-item hash
__global__ void checksum_kernel(const unsigned char* __restrict__ items, size_t n, u32 item_bytes, u64* out2) {
    u64 s = 0, x = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned char* p = items + i * item_bytes;
        u64 h = 0x243F6A8885A308D3ull;
        u32 off = 0;
        for (; off + 8 <= item_bytes; off += 8) {
            u64 w;
            if ((item_bytes & 7) == 0) w = *(const u64*)(p + off);
            else { w = 0; for (int k = 0; k < 8; ++k) w |= (u64)p[off + k] << (8 * k); }
            h = splitmix64_dev(h ^ w);
        }
        if (off < item_bytes) {
            u64 w = 0;
            for (u32 k = 0; off + k < item_bytes; ++k) w |= (u64)p[off + k] << (8 * k);
            h = splitmix64_dev(h ^ w);
        }
        s += h; x ^= h;
    }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); x ^= __shfl_xor_sync(0xffffffffu, x, o); }
    if (lane_id() == 0) { atomicAdd(&out2[0], s); atomicXor(&out2[1], x); }
}

__device__ __forceinline__ int key_less_dev(const unsigned char* a, const unsigned char* b, u32 off, u32 kb, u32 kind) {
    if (kind == TG_KEY_UINT_LE) {
        for (int k = (int)kb - 1; k >= 0; --k) {
            unsigned char x = a[off + k], y = b[off + k];
            if (x != y) return x < y;
        }
        return 0;
    }
    for (u32 k = 0; k < kb; ++k) {
        unsigned char x = a[off + k], y = b[off + k];
        if (x != y) return x < y;
    }
    return 0;
}
__global__ void is_sorted_kernel(const unsigned char* __restrict__ items, size_t n, tg_key_desc d, u64* violations) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    u64 bad = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += stride) {
        const unsigned char* a = items + i * d.item_bytes;
        const unsigned char* b = a + d.item_bytes;
        int wrong = d.descending ? key_less_dev(a, b, d.key_offset, d.key_bytes, d.key_kind)
                                 : key_less_dev(b, a, d.key_offset, d.key_bytes, d.key_kind);
        bad += wrong;
    }
    if (bad) atomicAdd(violations, bad);
}

}  // namespace

extern "C" {

int tg_gen_sort_uniform(tg_ctx* ctx, void* d_out, uint64_t begin, uint64_t n, uint64_t seed) {
    if (n) TG_LAUNCH(ctx, gen_sort_uniform_kernel, ctx->sm_count * 8, 256, 0, (u64*)d_out, begin, n, seed);
    return TG_OK;
}
int tg_gen_reduce_uniform(tg_ctx* ctx, void* d_out, uint64_t begin, uint64_t n, uint64_t seed, uint64_t universe, int exact) {
    if (n) TG_LAUNCH(ctx, gen_reduce_uniform_kernel, ctx->sm_count * 8, 256, 0, (ulonglong2*)d_out, begin, n, seed, universe, exact);
    return TG_OK;
}
int tg_gen_sort_zipf(tg_ctx* ctx, void* d_out, uint64_t begin, uint64_t n, uint64_t seed, const void* d_cdf, uint64_t universe) {
    if (n) TG_LAUNCH(ctx, gen_sort_zipf_kernel, ctx->sm_count * 8, 256, 0, (u64*)d_out, begin, n, seed, (const double*)d_cdf, universe);
    return TG_OK;
}
int tg_gen_reduce_zipf(tg_ctx* ctx, void* d_out, uint64_t begin, uint64_t n, uint64_t seed, const void* d_cdf, uint64_t universe, int exact) {
    if (n) TG_LAUNCH(ctx, gen_reduce_zipf_kernel, ctx->sm_count * 8, 256, 0, (ulonglong2*)d_out, begin, n, seed, (const double*)d_cdf, universe, exact);
    return TG_OK;
}
int tg_gen_records(tg_ctx* ctx, void* d_out, uint64_t begin, uint64_t n, uint64_t seed) {
    if (n) TG_LAUNCH(ctx, gen_records_kernel, ctx->sm_count * 8, 256, 0, (unsigned char*)d_out, begin, n, seed);
    return TG_OK;
}

int tg_checksum(tg_ctx* ctx, const void* d_items, size_t n, uint32_t item_bytes, uint64_t out_sum_xor[2]) {
    u64* d;
    TG_TRY(tg_ws_get(ctx, WS_MISC, 4096, (void**)&d));
    TG_CUDA(ctx, cudaMemsetAsync(d, 0, 16, ctx->stream));
    if (n) TG_LAUNCH(ctx, checksum_kernel, ctx->sm_count * 8, 256, 0, (const unsigned char*)d_items, n, item_bytes, d);
    TG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
    TG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out_sum_xor, ctx->pinned, 16);
    return TG_OK;
}

int tg_is_sorted(tg_ctx* ctx, const tg_key_desc* desc, const void* d_items, size_t n, uint64_t* out_violations) {
    u64* d;
    TG_TRY(tg_ws_get(ctx, WS_MISC, 4096, (void**)&d));
    TG_CUDA(ctx, cudaMemsetAsync(d, 0, 8, ctx->stream));
    if (n > 1) TG_LAUNCH(ctx, is_sorted_kernel, ctx->sm_count * 8, 256, 0, (const unsigned char*)d_items, n, *desc, d);
    TG_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    TG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out_violations, ctx->pinned, 8);
    return TG_OK;
}

}  // extern "C"
