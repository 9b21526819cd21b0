// tg_partition.cuh — the stable single-pass multi-way partition kernel ("onesweep") shared by
//   * the LSB radix sort            (digit = 8 key bits)                       tg_radix_sort.cu
//   * the splitter classify/scatter (digit = destination worker by splitters)  tg_sample_sort.cu
//   * the hash partition            (digit = Hash128to64(0,key) % p)           tg_reduce.cu
//
// One persistent CTA per SM walks tiles in static round-robin order.  A tile (64 KB) is staged into
// shared memory by the TMA unit (cp.async.bulk + mbarrier, double buffered so the next tile lands while
// the current one is processed), ranked stably with warp-synchronous match.any + warp-private counters,
// positioned globally by a chained scan with batched decoupled look-back, reordered by digit inside the
// (re-used) landing buffer and written out so that consecutive threads write consecutive addresses.
// HBM traffic: read n*s + write n*s + ~3 % look-back state.
#pragma once
#include <stdlib.h>

#include "tg_common.cuh"

namespace tgp {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;

constexpr u32 FLAG_PARTIAL = 1u << 30;
constexpr u32 FLAG_INCL = 2u << 30;
constexpr u32 VALUE_MASK = (1u << 30) - 1;

template <int WORDS> struct ItemT;
template <> struct ItemT<1> { typedef u64 type; };
template <> struct ItemT<2> { typedef ulonglong2 type; };

__device__ __forceinline__ u64 item_word(const u64& v, int) { return v; }
__device__ __forceinline__ u64 item_word(const ulonglong2& v, int w) { return w ? v.y : v.x; }

// digit functors: u32 operator()(item, position of the item in the input) -> [0, RADIX)
struct RadixDigit {
    int word, shift;
    u32 flip;
    static constexpr bool kStoreDigit = false;      // cheap to recompute in the write-out
    template <class Item>
    __device__ __forceinline__ u32 operator()(const Item& v, u32) const {
        return ((u32)(item_word(v, word) >> shift) & (RADIX - 1)) ^ flip;
    }
};

// warp-wide "which lanes hold the same 8-bit digit": 8 ballots + bit ops.  (match.any.sync costs ~64 issue
// cycles of the ADU pipe per warp instruction on sm_100 — measured 66 % ADU-bound in profiles/ r1a — the
// ballot form runs on the ordinary ALU/vote path.)
template <bool CHECK_VALID>
__device__ __forceinline__ u32 match_digit(u32 d, bool valid) {
    u32 peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < RADIX_BITS; ++b) {
        bool bit = (d >> b) & 1u;
        u32 m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    if (CHECK_VALID) {
        u32 m = __ballot_sync(0xffffffffu, valid);
        peers &= valid ? m : ~m;
    }
    return peers;
}

template <int WORDS, int THREADS>
struct SweepCfg {
    static constexpr int ITEM_BYTES = 8 * WORDS;
    static constexpr int ITEMS = 16 / WORDS;                 // per thread: 128 bytes of items
    static constexpr int TILE = THREADS * ITEMS;             // items per tile
    static constexpr int TILE_BYTES = TILE * ITEM_BYTES;     // 64 KB at 512 threads
    static constexpr int NWARPS = THREADS / 32;
    static constexpr int SMEM = 2 * TILE_BYTES + 2 * NWARPS * RADIX * 4 + 2 * RADIX * 4 + 64 + 16 + TILE + 256;
};

// per-bucket counts of one digit function (the pre-pass of a stand-alone partition)
template <int WORDS, class DigitFn>
__global__ void __launch_bounds__(512) bucket_count_kernel(const typename ItemT<WORDS>::type* __restrict__ in, u32 n,
                                                           DigitFn fn, u32* __restrict__ gcount) {
    __shared__ u32 sh[RADIX];
    for (int i = threadIdx.x; i < RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const u32 stride = gridDim.x * blockDim.x;
    for (u32 base = blockIdx.x * blockDim.x; base < n; base += stride) {
        u32 i = base + threadIdx.x;
        bool valid = i < n;
        u32 act = __ballot_sync(0xffffffffu, valid);
        if (!valid) continue;
        u32 d = fn(in[i], i);
        u32 d0 = __shfl_sync(act, d, __ffs(act) - 1);
        if (__all_sync(act, d == d0)) {
            if (lane_id() == (u32)(__ffs(act) - 1)) atomicAdd(&sh[d], __popc(act));
        }
        else atomicAdd(&sh[d], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RADIX; i += blockDim.x)
        if (sh[i]) atomicAdd(&gcount[i], sh[i]);
}

// exclusive scan of npass digit histograms -> global bases; skip[p] = 1 if one bin holds everything
static __global__ void scan_hist_kernel(const u32* __restrict__ ghist, u32* __restrict__ gbase, u32* __restrict__ skip,
                                 int npass, u32 n) {
    __shared__ u32 warp_tot[RADIX / 32];
    const int d = threadIdx.x;      // RADIX threads
    for (int p = 0; p < npass; ++p) {
        u32 c = ghist[p * RADIX + d];
        u32 incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((d & 31) >= o) incl += t;
        }
        if ((d & 31) == 31) warp_tot[d >> 5] = incl;
        __syncthreads();
        u32 add = 0;
        for (int w = 0; w < (d >> 5); ++w) add += warp_tot[w];
        gbase[p * RADIX + d] = add + incl - c;
        u32 full = __syncthreads_or(c == n);
        if (d == 0 && skip) skip[p] = full ? 1u : 0u;
        __syncthreads();
    }
}

template <int WORDS, int THREADS, int MINB, int RANK, class DigitFn>
__global__ void __launch_bounds__(THREADS, MINB)
partition_kernel(const typename ItemT<WORDS>::type* __restrict__ in, typename ItemT<WORDS>::type* __restrict__ out,
                 u32 n, const DigitFn fn_param, const u32* __restrict__ gbase, u32* __restrict__ status, int dbg) {
    typedef typename ItemT<WORDS>::type Item;
    typedef SweepCfg<WORDS, THREADS> C;
    constexpr int ITEMS = C::ITEMS, TILE = C::TILE, NWARPS = C::NWARPS;
    constexpr int LB = 8;       // look-back batch: predecessors fetched concurrently

    // plain pointer arithmetic on the shared array keeps the shared address space (LDS/STS, 32-bit addresses)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Item* const buf0 = reinterpret_cast<Item*>(smem_raw);
    Item* const buf1 = reinterpret_cast<Item*>(smem_raw + C::TILE_BYTES);
    u32* const whist = reinterpret_cast<u32*>(smem_raw + 2 * C::TILE_BYTES);     // [NWARPS][RADIX]
    u32* const goff = whist + NWARPS * RADIX;                                    // [RADIX]
    u32* const warp_tot = goff + RADIX;                                          // [16]
    u64* const mbar = reinterpret_cast<u64*>(warp_tot + 16);                     // [2]
    u32* const wmask = reinterpret_cast<u32*>(mbar + 2);                         // [NWARPS][RADIX] (RANK != 0)
    unsigned char* const dig = reinterpret_cast<unsigned char*>(wmask + NWARPS * RADIX);   // [TILE], only if kStoreDigit

    const DigitFn fn = fn_param;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 num_tiles = (n + TILE - 1) / TILE;
    const u32 lt = lanemask_lt();
    u32* const whist_w = whist + warp * RADIX;
    u32* const wmask_w = wmask + warp * RADIX;
    const u32 wbase = warp * 32 * ITEMS;
    if (RANK != 0) {
#pragma unroll
        for (int i = 0; i < RADIX / 32; ++i) wmask_w[i * 32 + lane] = 0;
    }

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    u32 t = blockIdx.x;
    if (tid == 0 && t < num_tiles && (size_t)(t + 1) * TILE <= n) {
        mbar_expect_tx(&mbar[0], C::TILE_BYTES);
        bulk_g2s(buf0, in + (size_t)t * TILE, C::TILE_BYTES, &mbar[0]);
    }

    for (u32 it = 0; t < num_tiles; t += gridDim.x, ++it) {
        const int cur = it & 1;
        Item* const buf = cur ? buf1 : buf0;
        Item* const nbuf = cur ? buf0 : buf1;
        const u32 tile_base = t * TILE;
        const bool full_tile = (size_t)tile_base + TILE <= n;
        const u32 tile_valid = full_tile ? (u32)TILE : n - tile_base;

        // prefetch the CTA's next tile (TMA unit, async proxy) into the other buffer
        {
            u32 tn = t + gridDim.x;
            if (tid == 0 && tn < num_tiles && (size_t)(tn + 1) * TILE <= n) {
                fence_proxy_async();
                mbar_expect_tx(&mbar[cur ^ 1], C::TILE_BYTES);
                bulk_g2s(nbuf, in + (size_t)tn * TILE, C::TILE_BYTES, &mbar[cur ^ 1]);
            }
        }
        // zero this warp's private digit counters
#pragma unroll
        for (int i = 0; i < RADIX / 32; ++i) whist_w[i * 32 + lane] = 0;

        // ---- items to registers: warp w owns tile positions [w*32*ITEMS, (w+1)*32*ITEMS), round-striped
        Item key[ITEMS];
        if (full_tile) {
            mbar_wait(&mbar[cur], (it >> 1) & 1);
            const Item* src = buf + wbase + lane;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) key[i] = src[i * 32];
        }
        else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                u32 p = wbase + i * 32 + lane;
                if (p < tile_valid) key[i] = in[(size_t)tile_base + p];
            }
        }
        __syncwarp();

        // ---- stable rank inside the warp: ballots group equal digits, the group leader bumps the counter
        unsigned short rank[ITEMS];
        unsigned char mydig[ITEMS];
        if (full_tile && (dbg & 2)) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                u32 d = fn(key[i], tile_base + wbase + i * 32 + lane);
                if (DigitFn::kStoreDigit) mydig[i] = (unsigned char)d;
                rank[i] = (unsigned short)0;
                if (lane == 0) whist_w[d] += 1;
            }
        }
        else if (full_tile) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                u32 d = fn(key[i], tile_base + wbase + i * 32 + lane);
                if (DigitFn::kStoreDigit) mydig[i] = (unsigned char)d;
                u32 peers;
                if (RANK == 0 || (RANK == 2 && (i & 1))) peers = match_digit<false>(d, true);
                else {
                    // warp-private mask table: one native 32-bit shared-memory atomic per key instead of 8 ballots
                    atomicOr(&wmask_w[d], 1u << lane);
                    __syncwarp();
                    peers = wmask_w[d];
                    __syncwarp();
                }
                u32 below = peers & lt;
                u32 old = 0;
                if (below == 0) {                       // lowest lane of its group
                    if (!(RANK == 0 || (RANK == 2 && (i & 1)))) wmask_w[d] = 0;
                    old = whist_w[d];
                    whist_w[d] = old + __popc(peers);
                }
                old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
                rank[i] = (unsigned short)(old + __popc(below));
                __syncwarp();
            }
        }
        else {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                u32 p = wbase + i * 32 + lane;
                bool valid = p < tile_valid;
                u32 d = valid ? fn(key[i], tile_base + p) : 0u;
                if (DigitFn::kStoreDigit) mydig[i] = (unsigned char)d;
                u32 peers = match_digit<true>(d, valid);
                u32 below = peers & lt;
                u32 old = 0;
                if (below == 0 && valid) {
                    old = whist_w[d];
                    whist_w[d] = old + __popc(peers);
                }
                old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
                rank[i] = (unsigned short)(old + __popc(below));
                __syncwarp();
            }
        }
        __syncthreads();      // all items are in registers (buf is free), all warp counters final

        // ---- per-digit tile count; publish PARTIAL as early as possible; tile-local digit starts
        u32 count = 0, my_start = 0;
        if (tid < RADIX) {
#pragma unroll
            for (int w = 0; w < NWARPS; ++w) count += whist[w * RADIX + tid];
            st_relaxed_u32(&status[(size_t)t * RADIX + tid], count | (t == 0 ? FLAG_INCL : FLAG_PARTIAL));
            u32 incl = count;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                u32 v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) warp_tot[warp] = incl;
            my_start = incl - count;             // completed below with the preceding warps' totals
        }
        __syncthreads();
        if (tid < RADIX) {
            for (int w = 0; w < warp; ++w) my_start += warp_tot[w];
            // warp counters -> tile-local write positions of each (warp, digit) group
            u32 off = my_start;
#pragma unroll
            for (int w = 0; w < NWARPS; ++w) {
                u32 c = whist[w * RADIX + tid];
                whist[w * RADIX + tid] = off;
                off += c;
            }
        }
        __syncthreads();

        // ---- decoupled look-back (threads 0..RADIX-1, one digit each, LB predecessors per round trip)
        if (tid < RADIX) {
            u32 excl = 0;
            if (t > 0 && !(dbg & 1)) {
                int look = (int)t - 1;
                bool done = false;
                while (!done) {
                    u32 v[LB];
#pragma unroll
                    for (int k = 0; k < LB; ++k)
                        v[k] = (look - k >= 0) ? ld_relaxed_u32(&status[(size_t)(look - k) * RADIX + tid]) : FLAG_INCL;
                    bool stalled = false;
#pragma unroll
                    for (int k = 0; k < LB; ++k) {
                        if (!done && !stalled) {
                            if (v[k] & FLAG_INCL) { excl += v[k] & VALUE_MASK; done = true; }
                            else if (v[k] & FLAG_PARTIAL) { excl += v[k] & VALUE_MASK; look--; }
                            else stalled = true;
                        }
                    }
                }
                st_relaxed_u32(&status[(size_t)t * RADIX + tid], (excl + count) | FLAG_INCL);
            }
            goff[tid] = gbase[tid] + excl - my_start;
        }

        // ---- scatter registers -> digit-ordered exchange buffer (reuses the landing buffer)
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            u32 p = wbase + i * 32 + lane;
            if (full_tile || p < tile_valid) {
                u32 d = DigitFn::kStoreDigit ? (u32)mydig[i] : fn(key[i], tile_base + p);
                u32 q = whist_w[d] + rank[i];
                buf[q] = key[i];
                if (DigitFn::kStoreDigit) dig[q] = (unsigned char)d;
            }
        }
        __syncthreads();

        // ---- coalesced write-out: consecutive threads write consecutive addresses inside a digit run
        {
            Item* const outp = out + tid;
            const Item* const bufp = buf + tid;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                if (full_tile || (u32)(i * THREADS + tid) < tile_valid) {
                    Item v = bufp[i * THREADS];
                    u32 d = DigitFn::kStoreDigit ? (u32)dig[i * THREADS + tid] : fn(v, 0);
                    outp[goff[d] + i * THREADS] = v;
                }
            }
        }
        __syncthreads();      // exchange buffer is re-armed as the TMA landing buffer two iterations later
    }
}

template <int WORDS>
__global__ void copy_items_kernel(const typename ItemT<WORDS>::type* __restrict__ in,
                                  typename ItemT<WORDS>::type* __restrict__ out, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}

// launch configurations (threads per CTA, CTAs per SM); selected once per process, TG_SWEEP_CFG overrides
struct SweepVariant { int threads, minb; };
constexpr SweepVariant kSweepVariants[] = { { 512, 1 }, { 256, 2 }, { 256, 3 }, { 384, 2 } };
inline int sweep_cfg() {
    static int cfg = -1;
    if (cfg < 0) {
        const char* e = getenv("TG_SWEEP_CFG");
        cfg = e ? atoi(e) : 0;
        if (cfg < 0 || cfg > 3) cfg = 0;
    }
    return cfg;
}
template <int WORDS>
inline u32 num_tiles_for(size_t n) {
    u32 tile = (u32)kSweepVariants[sweep_cfg()].threads * (16 / WORDS);
    return (u32)((n + tile - 1) / tile);
}

// timing experiments only (results are wrong when set): bit0 = skip the look-back wait, bit1 = skip ranking
inline int debug_flags() {
    static int f = -1;
    if (f < 0) { const char* e = getenv("TG_DEBUG_FLAGS"); f = e ? atoi(e) : 0; }
    return f;
}

inline int rank_mode() {
    static int m = -1;
    if (m < 0) {
        const char* e = getenv("TG_RANK_MODE");
        m = e ? atoi(e) : 2;
        if (m < 0 || m > 2) m = 2;
    }
    return m;
}

template <int WORDS, int THREADS, int MINB, int RANK, class DigitFn>
int launch_partition_v(tg_ctx* ctx, const void* in, void* out, u32 n, const DigitFn& fn, const u32* gbase, u32* status) {
    typedef typename ItemT<WORDS>::type Item;
    typedef SweepCfg<WORDS, THREADS> C;
    auto kern = partition_kernel<WORDS, THREADS, MINB, RANK, DigitFn>;
    int ctas_per_sm = 0;
    auto it = ctx->kernel_cfg.find((const void*)kern);
    if (it != ctx->kernel_cfg.end()) ctas_per_sm = it->second;
    else {
        TG_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        // the look-back needs every CTA of the grid resident: size the grid from the real occupancy
        TG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, THREADS, C::SMEM));
        if (ctas_per_sm < 1) return tg_set_error(ctx, TG_ERR_CUDA, "partition kernel does not fit on an SM");
        if (ctas_per_sm > MINB) ctas_per_sm = MINB;
        ctx->kernel_cfg[(const void*)kern] = ctas_per_sm;
    }
    u32 num_tiles = (n + C::TILE - 1) / C::TILE;
    int grid = ctx->sm_count * ctas_per_sm;
    if (grid > (int)num_tiles) grid = (int)num_tiles;
    TG_LAUNCH_T(ctx, TG_K_PARTITION, kern, grid, THREADS, C::SMEM, (const Item*)in, (Item*)out, n, fn, gbase, status, debug_flags());
    return TG_OK;
}

// launch one partition pass with precomputed global bases (status must be zeroed, num_tiles*RADIX words)
template <int WORDS, class DigitFn>
int launch_partition(tg_ctx* ctx, const void* in, void* out, u32 n, const DigitFn& fn, const u32* gbase, u32* status) {
    const int rm = rank_mode();
    switch (sweep_cfg()) {
    case 1: return launch_partition_v<WORDS, 256, 2, 0, DigitFn>(ctx, in, out, n, fn, gbase, status);
    case 2: return launch_partition_v<WORDS, 256, 3, 0, DigitFn>(ctx, in, out, n, fn, gbase, status);
    case 3: return launch_partition_v<WORDS, 384, 2, 0, DigitFn>(ctx, in, out, n, fn, gbase, status);
    default:
        if (rm == 1) return launch_partition_v<WORDS, 512, 1, 1, DigitFn>(ctx, in, out, n, fn, gbase, status);
        if (rm == 2) return launch_partition_v<WORDS, 512, 1, 2, DigitFn>(ctx, in, out, n, fn, gbase, status);
        return launch_partition_v<WORDS, 512, 1, 0, DigitFn>(ctx, in, out, n, fn, gbase, status);
    }
}

// stand-alone stable partition of n items into <= RADIX buckets: count pre-pass, scan, partition.
// d_counts_out (device, RADIX u32) receives the bucket counts.
template <int WORDS, class DigitFn>
int partition_items(tg_ctx* ctx, const void* in, void* out, u32 n, const DigitFn& fn, u32** d_counts_out) {
    typedef typename ItemT<WORDS>::type Item;
    u32 num_tiles = num_tiles_for<WORDS>(n);
    u32* hist;
    TG_TRY(tg_ws_get(ctx, WS_SORT_HIST, (size_t)2 * RADIX * 4, (void**)&hist));
    u32* status;
    TG_TRY(tg_ws_get(ctx, WS_SORT_STATUS, (size_t)num_tiles * RADIX * 4, (void**)&status));
    TG_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)2 * RADIX * 4, ctx->stream));
    TG_CUDA(ctx, cudaMemsetAsync(status, 0, (size_t)num_tiles * RADIX * 4, ctx->stream));
    if (n) {
        auto cnt = bucket_count_kernel<WORDS, DigitFn>;
        TG_LAUNCH(ctx, cnt, ctx->sm_count * 2, 512, 0, (const Item*)in, n, fn, hist);
        TG_LAUNCH(ctx, scan_hist_kernel, 1, RADIX, 0, hist, hist + RADIX, (u32*)nullptr, 1, n);
        TG_TRY((launch_partition<WORDS, DigitFn>(ctx, in, out, n, fn, hist + RADIX, status)));
    }
    if (d_counts_out) *d_counts_out = hist;
    return TG_OK;
}

}  // namespace tgp
