// tg_exchange.cuh — the all-to-all exchange of the collective operators, fused into the partition pass.
//
// Reference path replaced: data::MixStream / CatStream writers of SortNode::TransmitItems (api/sort.hpp:434-535, :615-641)
// and of ReducePrePhaseEmitter (core/reduce_pre_phase.hpp:57-61, api/reduce_by_key.hpp:109-114): every item is appended
// to the Block stream of its destination worker, the multiplexer ships the Blocks (data/mix_stream.cpp:52-236).
//
// Here every worker owns an EXCHANGE WINDOW in its HBM that all peers of the job have mapped (CUDA IPC between the
// one-process-per-GPU workers, plain peer access between the worker threads of one Thrill process).  One exchange is
//   1. chunk histograms of the destination digit (one counting read)                                 chunk_hist_kernel
//   2. ncclAllGather of the p per-destination counts of every worker -> the p x p count matrix (the only host round trip:
//      every rank derives every rank's receive size from it, so window growth / errors are decided identically everywhere)
//   3. the stable partition pass with PEER = true: bucket d is stored straight into worker d's window at this worker's
//      offset (after the items of the lower ranks), over NVLink — classification, scatter and Alltoallv are ONE kernel
//   4. a tiny collective as the "all stores have landed" barrier.
// The received items lie grouped by source worker in rank order, each group in the sender's input order — the order
// CatStream delivers (stable).  TG_EXCHANGE=nccl (or peers that cannot map each other) selects the two-step form
// instead: local partition, then grouped ncclSend/ncclRecv into the window.
#pragma once
#include "tg_segmented.cuh"

namespace tgp {

struct XchgResult {
    void* d_recv = nullptr;                 // the received items (this worker's window)
    u64 n_recv = 0;
    u64 recv_cnt[TG_MAX_RANKS] = { 0 };     // items received from each source rank (rank order = layout order)
    u64 send_cnt[TG_MAX_RANKS] = { 0 };
};

// tg_exchange.cu
int xwin_negotiate(tg_ctx* ctx);                                    // collective, first use: decides P2P vs NCCL mode
int xwin_ensure(tg_ctx* ctx, size_t bytes_all_ranks);               // collective: every rank's window >= bytes
int xwin_barrier(tg_ctx* ctx);                                      // stream-ordered cross-rank barrier
int xchg_counts(tg_ctx* ctx, const u32* d_totals, int item_bytes, XchgResult* res, u64* need_bytes_max);
int xchg_upload_dest(tg_ctx* ctx, int item_bytes, const XchgResult& res, void*** d_dbase_out);
void xchg_recv_offsets(tg_ctx* ctx, u64* before);                  // before[d] = items of the lower ranks in worker d's window

// Stable partition of n local items by fn (destination worker, < p) + Alltoallv.  Collective.
template <int WORDS, class DigitFn>
int exchange_scatter(tg_ctx* ctx, const void* d_in, size_t n, const DigitFn& fn, XchgResult* res) {
    typedef typename ItemT<WORDS>::type Item;
    const int p = ctx->nranks;
    const size_t s = sizeof(Item);
    if (n >= (1u << 30)) n = 0;              // (the caller has validated n on every rank before the first collective)
    TG_TRY(xwin_negotiate(ctx));
    // (1) destination histogram per chunk
    const ChunkGeom g = chunk_geometry<WORDS>(ctx, n);
    const size_t cw = (size_t)(g.nchunks > 0 ? g.nchunks : 1) * RADIX;
    u32* tab;
    TG_TRY(tg_ws_get(ctx, WS_SORT_HIST2, (2 * cw + 2 * RADIX + 16) * 4, (void**)&tab));
    u32* chunkcount = tab;
    u32* chunkbase = tab + cw;
    u32* totals = chunkbase + cw;
    if (n) {
        TG_LAUNCH_T(ctx, TG_K_RADIX_HIST, (chunk_hist_kernel<WORDS, DigitFn, false>), g.nchunks, 512, 0, (const Item*)d_in, (u32)n,
                    g.chunk_items, fn, chunkcount, (u64*)nullptr);
        TG_LAUNCH(ctx, chunk_scan_kernel, 1, 4 * RADIX, 0, chunkcount, g.nchunks, totals, totals + RADIX, chunkbase);
    }
    else TG_CUDA(ctx, cudaMemsetAsync(totals, 0, 2 * RADIX * 4, ctx->stream));
    // (2) count matrix; every rank learns every rank's receive size
    u64 need = 0;
    TG_TRY(xchg_counts(ctx, totals, (int)s, res, &need));
    TG_TRY(xwin_ensure(ctx, need));
    res->d_recv = ctx->xwin.base;
    if (ctx->xwin.mode == 1) {
        // (3) classification + scatter + Alltoallv in one pass: stores into the peers' windows
        if (n) {
            void** d_dbase;
            TG_TRY(xchg_upload_dest(ctx, (int)s, *res, &d_dbase));
            std::vector<u32> chunk_size(g.nchunks, g.chunk_items);
            chunk_size[g.nchunks - 1] = (u32)(n - (size_t)(g.nchunks - 1) * g.chunk_items);
            uint4* d_tiles;
            u32 total = 0;
            TG_TRY(build_tile_list(ctx, g.nchunks, chunk_size.data(), tile_items<WORDS>(), 0, WS_SEG_TILES2, &d_tiles, &total));
            u32* status;
            TG_TRY(tg_ws_get(ctx, WS_SORT_STATUS, (size_t)total * RADIX * 4, (void**)&status));
            TG_CUDA(ctx, cudaMemsetAsync(status, 0, (size_t)total * RADIX * 4, ctx->stream));
            SegList sl = { d_tiles, chunkbase, total, nullptr };
            const int xprof = ctx->profile ? tg_prof_begin(ctx, TG_K_EXCHANGE) : -1;
            TG_TRY((launch_partition_peer<WORDS, DigitFn>(ctx, d_in, (u32)n, fn, status, sl, (Item* const*)d_dbase)));
            if (xprof >= 0) tg_prof_end(ctx, xprof);
        }
        // (4) every peer's stores into this window are complete after the barrier
        TG_TRY(xwin_barrier(ctx));
        return TG_OK;
    }
    // two-step form: local stable partition, then grouped send/recv into the window
    void* d_part;
    TG_TRY(tg_ws_get(ctx, WS_XCHG_SEND, (n + 2) * s, &d_part));
    if (n) {
        std::vector<u32> chunk_size(g.nchunks, g.chunk_items);
        chunk_size[g.nchunks - 1] = (u32)(n - (size_t)(g.nchunks - 1) * g.chunk_items);
        uint4* d_tiles;
        u32 total = 0;
        TG_TRY(build_tile_list(ctx, g.nchunks, chunk_size.data(), tile_items<WORDS>(), 0, WS_SEG_TILES2, &d_tiles, &total));
        u32* status;
        TG_TRY(tg_ws_get(ctx, WS_SORT_STATUS, (size_t)total * RADIX * 4, (void**)&status));
        TG_CUDA(ctx, cudaMemsetAsync(status, 0, (size_t)total * RADIX * 4, ctx->stream));
        SegList sl = { d_tiles, chunkbase, total, nullptr };
        TG_TRY((launch_partition_seg<WORDS, DigitFn>(ctx, d_in, d_part, (u32)n, fn, status, sl)));
    }
    const int xprof = ctx->profile ? tg_prof_begin(ctx, TG_K_EXCHANGE) : -1;
    TG_NCCL(ctx, ncclGroupStart());
    u64 soff = 0, roff = 0;
    for (int r = 0; r < p; ++r) {
        if (res->send_cnt[r]) TG_NCCL(ctx, ncclSend((const char*)d_part + soff * s, res->send_cnt[r] * s, ncclUint8, r, ctx->comm, ctx->stream));
        if (res->recv_cnt[r]) TG_NCCL(ctx, ncclRecv((char*)res->d_recv + roff * s, res->recv_cnt[r] * s, ncclUint8, r, ctx->comm, ctx->stream));
        soff += res->send_cnt[r];
        roff += res->recv_cnt[r];
    }
    TG_NCCL(ctx, ncclGroupEnd());
    if (xprof >= 0) tg_prof_end(ctx, xprof);
    return TG_OK;
}

}  // namespace tgp
