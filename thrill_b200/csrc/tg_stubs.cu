// temporary: entry points not implemented yet (removed as the kernels land)
#include "tg_common.cuh"
#define NI(ctx) return tg_set_error(ctx, TG_ERR_ARG, "%s: not implemented yet", __func__)
extern "C" {
uint64_t tg_sample_size(uint64_t) { return 0; }
int tg_select_splitters(const tg_key_desc*, void*, uint64_t, uint32_t, void*) { return TG_ERR_ARG; }
int tg_draw_samples(tg_ctx* c, const tg_key_desc*, const void*, size_t, uint64_t, uint64_t, void*, uint64_t*) { NI(c); }
int tg_classify_scatter(tg_ctx* c, const tg_key_desc*, const void*, size_t, uint64_t, const void*, uint32_t, void*, uint64_t*) { NI(c); }
int tg_kway_merge(tg_ctx* c, const tg_key_desc*, const void*, const uint64_t*, uint32_t, void*, void*) { NI(c); }
int tg_hash_aggregate(tg_ctx* c, const tg_kv_desc*, const void*, size_t, void*, uint64_t*) { NI(c); }
int tg_hash_partition(tg_ctx* c, const tg_kv_desc*, const void*, size_t, uint32_t, void*, uint64_t*) { NI(c); }
int tg_sort(tg_ctx* c, const tg_key_desc*, void*, size_t, uint64_t, void**, size_t*) { NI(c); }
int tg_reduce_by_key(tg_ctx* c, const tg_kv_desc*, const void*, size_t, void**, size_t*) { NI(c); }
int tg_sort_file(tg_ctx* c, const tg_key_desc*, const tg_block*, size_t, uint64_t, size_t*) { NI(c); }
int tg_reduce_file(tg_ctx* c, const tg_kv_desc*, const tg_block*, size_t, size_t*) { NI(c); }
int tg_fetch_output(tg_ctx* c, const tg_block_mut*, size_t) { NI(c); }
}
