/*******************************************************************************
 * oracle/thrill_oracle.h — TEST INFRASTRUCTURE ONLY (parity oracle; see thrill_oracle.c header).
 *
 * CPU restatement of the reference's Sort / ReduceByKey hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  Nothing under thrill_b200/ may include, link or dlopen it.
 ******************************************************************************/
#ifndef THRILL_ORACLE_H
#define THRILL_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- item/key descriptor (mirrors tg_key_desc of include/thrill_gpu.h) ---- */
enum { TO_KEY_UINT_LE = 0, TO_KEY_BYTES_BE = 1 };
typedef struct {
    uint32_t item_bytes;   /* sizeof(ValueType) as serialized (data/serialization.hpp:34-49) */
    uint32_t key_offset;   /* byte offset of the key inside the item */
    uint32_t key_bytes;    /* 1..8 for UINT_LE, any for BYTES_BE */
    uint32_t key_kind;     /* TO_KEY_* */
} to_key_desc;

/* reduce operators on the 8-byte value of a (u64 key, 8-byte value) TableItem */
enum { TO_OP_SUM_F64 = 0, TO_OP_SUM_U64 = 1, TO_OP_MIN_U64 = 2, TO_OP_MAX_U64 = 3,
       TO_OP_MIN_F64 = 4, TO_OP_MAX_F64 = 5, TO_OP_FIRST = 6 };
typedef struct { uint64_t key; uint64_t val; } to_kv;   /* pair<u64, 8-byte value>, member-wise 16 B */

/* ---- generators ---- */
uint64_t to_splitmix64(uint64_t x);
void to_gen_sort_uniform(uint64_t* out, uint64_t begin, uint64_t n, uint64_t seed);
void to_gen_reduce_uniform(to_kv* out, uint64_t begin, uint64_t n, uint64_t seed, uint64_t universe, int exact);
void to_zipf_build_cdf(double* cdf, uint64_t universe, double s);
uint64_t to_zipf_rank(const double* cdf, uint64_t universe, double u);
void to_gen_sort_zipf(uint64_t* out, uint64_t begin, uint64_t n, uint64_t seed, const double* cdf, uint64_t universe);
void to_gen_reduce_zipf(to_kv* out, uint64_t begin, uint64_t n, uint64_t seed, const double* cdf, uint64_t universe, int exact);
void to_gen_records(uint8_t* out, uint64_t begin, uint64_t n, uint64_t seed);

/* ---- hashing / partitioning (common/hash.hpp:64-73, core/reduce_functional.hpp:60-72) ---- */
uint64_t to_hash128to64(uint64_t upper, uint64_t lower);
void to_reduce_by_hash(uint64_t key, uint64_t salt, uint64_t num_partitions,
                       uint64_t* partition_id, uint64_t* remaining_hash);
void to_hash_partition_ids(const uint64_t* keys, uint64_t n, uint64_t stride_bytes, uint64_t salt,
                           uint64_t num_partitions, uint32_t* out_partition);

/* ---- Sort pieces ---- */
uint64_t to_sample_size(uint64_t count, double imbalance);                     /* reservoir_sampling.hpp:270-275 */
int to_less(const to_key_desc* d, const void* a, const void* b);
void to_sort_items(void* items, uint64_t n, const to_key_desc* d);              /* api/sort.hpp:789-796 */
/* samples: nsamples x (item_bytes + 8) packed (item, u64 global index); out: p-1 packed the same way */
uint64_t to_select_splitters(void* samples, uint64_t nsamples, uint64_t p, const to_key_desc* d, void* out_splitters);
/* splitters: k-1 packed (item,u64 idx) incl. sentinel padding; tree: (k+1) items, 1-based (api/sort.hpp:380-417) */
void to_build_tree(const void* splitters, uint64_t k, const to_key_desc* d, void* tree);
void to_classify(const void* items, uint64_t n, uint64_t prefix_items, const void* tree, uint64_t k,
                 uint64_t log_k, const void* splitters, const to_key_desc* d, uint32_t* out_bucket);
/* k-way merge, tlx::LoserTreeCopy semantics (loser_tree.hpp:54-292), core/multiway_merge.hpp:30-116 */
void to_multiway_merge(const void* const* runs, const uint64_t* run_items, uint32_t k,
                       const to_key_desc* d, int stable, void* out);
/* whole operator on p simulated workers: in = concatenation of worker shards (local_counts[p]);
 * out = concatenation of per-worker outputs, out_counts[p] (api/sort.hpp:537-663).  rng_seed drives the
 * sample picks (the reference seeds from std::random_device, api/context.cpp:1190). */
void to_sort_operator(const void* in, const uint64_t* local_counts, uint32_t p, const to_key_desc* d,
                      int stable, uint64_t rng_seed, void* out, uint64_t* out_counts);

/* ---- ReduceByKey pieces ---- */
typedef struct to_table to_table;
typedef void (*to_emit_fn)(void* ctx, uint64_t partition_id, const to_kv* item);
to_table* to_table_new(uint64_t num_partitions, uint64_t limit_memory_bytes, int immediate_flush,
                       uint64_t salt, int op, to_emit_fn emit, void* emit_ctx);
int  to_table_insert(to_table* t, const to_kv* kv);           /* reduce_probing_hash_table.hpp:190-268 */
void to_table_flush_all(to_table* t);                          /* :484-488 */
uint64_t to_table_num_items(const to_table* t);
uint64_t to_table_partition_size(const to_table* t, uint64_t partition);
void to_table_free(to_table* t);
/* pre phase (core/reduce_pre_phase.hpp:103-201): out_part[i] = partition of the i-th emitted item */
uint64_t to_reduce_pre_phase(const to_kv* in, uint64_t n, uint64_t p, uint64_t limit_memory_bytes, int op,
                             to_kv* out_items, uint32_t* out_part, uint64_t out_capacity);
/* post phase incl. spill + salted re-reduce (core/reduce_by_hash_post_phase.hpp:44-281) */
uint64_t to_reduce_post_phase(const to_kv* in, uint64_t n, uint64_t limit_memory_bytes, int op,
                              to_kv* out_items, uint64_t out_capacity, uint64_t* out_reduce_iterations);
/* whole operator on p simulated workers (api/reduce_by_key.hpp:100-211) */
uint64_t to_reduce_operator(const to_kv* in, const uint64_t* local_counts, uint32_t p,
                            uint64_t mem_limit_bytes, int op, to_kv* out, uint64_t* out_counts);
/* straightforward aggregate used as the scalable CPU checker: sort by key, fold equal keys in input order */
uint64_t to_reduce_simple(const to_kv* in, uint64_t n, int op, to_kv* out);
/* api/reduce_to_index.hpp:60-237, core/reduce_by_index_post_phase.hpp:141-222: dense result of `size` items */
uint64_t to_reduce_to_index(const to_kv* in, uint64_t n, uint64_t size, to_kv neutral, int op, to_kv* out);

/* ---- data::File / data::Block layout (data/file.hpp:56-283, data/block.hpp:52-145) ---- */
typedef struct {
    uint64_t begin, end;        /* valid byte range inside the ByteBlock */
    uint64_t first_item;        /* absolute offset (in the ByteBlock) of the first item STARTING in it */
    uint64_t num_items;         /* items starting in this block */
} to_block_meta;
/* emulate BlockWriter for fixed-size items (data/block_writer.hpp:311-335,386-418): block sizes start at
 * start_block_size and double up to max_block_size; items straddle blocks.  Returns number of blocks. */
uint64_t to_file_layout(uint64_t num_items, uint32_t item_bytes, uint64_t start_block_size,
                        uint64_t max_block_size, to_block_meta* out, uint64_t out_capacity);

#ifdef __cplusplus
}
#endif
#endif /* THRILL_ORACLE_H */
