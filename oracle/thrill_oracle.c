/*******************************************************************************
 * oracle/thrill_oracle.c — TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * Plain-C CPU restatement of the reference's (thrill/thrill @ 12c5b59b) Sort / ReduceByKey hot path,
 * SURVEY.md §8(a) rows a1..a12.  Every function cites the reference file:line it follows (paths relative
 * to /root/reference).  Parity status: PINNED — tests/test_oracle_*.py check this file against
 *   (1) the known-answer tests of the reference's own suite (tests/core/multiway_merge_test.cpp:33-86,
 *       tests/core/reduce_hash_table_test.cpp:54-144, tests/core/reduce_pre_phase_test.cpp:44-127,
 *       tests/core/reduce_post_phase_test.cpp:36-112, tests/api/sort_node_test.cpp, tests/api/
 *       reduce_node_test.cpp:47-139, tests/data/file_test.cpp:30-122), restated with the same inputs;
 *   (2) outputs of the UNMODIFIED reference itself (oracle/_ref/thrill_ref_driver, built by
 *       oracle/ref/Makefile) committed as fixtures under tests/golden/ by tests/golden/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the
 * library built from this file; thrill_b200/ (the product) never does.
 ******************************************************************************/
#include "thrill_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ========================================================================== */
/* generators — SURVEY.md §8(d); identical arithmetic in oracle/ref/ref_driver.cpp */

uint64_t to_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t dbits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline double bitsd(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

void to_gen_sort_uniform(uint64_t* out, uint64_t begin, uint64_t n, uint64_t seed) {
    for (uint64_t i = 0; i < n; ++i) out[i] = to_splitmix64(begin + i + seed);
}

static inline uint64_t gen_val(uint64_t i, uint64_t seed, int exact) {
    uint64_t r = to_splitmix64(i + seed + (1ull << 40));
    if (exact == 2) return r % 1024;                 /* u64 payload (reduce_u64) */
    return dbits(exact ? (double)(r % 1024) : u01(r));
}

void to_gen_reduce_uniform(to_kv* out, uint64_t begin, uint64_t n, uint64_t seed, uint64_t universe, int exact) {
    for (uint64_t j = 0; j < n; ++j) {
        uint64_t i = begin + j;
        out[j].key = 1 + to_splitmix64(i + seed) % universe;
        out[j].val = gen_val(i, seed, exact);
    }
}

/* probabilities as common/zipf_distribution.hpp:119-140 (k^-s normalised), accumulated to a CDF */
void to_zipf_build_cdf(double* cdf, uint64_t universe, double s) {
    double p_sum = 0.0;
    for (uint64_t k = 1; k <= universe; ++k) p_sum += 1.0 / pow((double)k, s);
    double p_norm = 1.0 / p_sum, acc = 0.0;
    for (uint64_t k = 1; k <= universe; ++k) {
        acc += (1.0 / pow((double)k, s)) * p_norm;
        cdf[k - 1] = acc;
    }
}
/* smallest k with cdf[k-1] > u (std::upper_bound), clamped to universe */
uint64_t to_zipf_rank(const double* cdf, uint64_t universe, double u) {
    uint64_t lo = 0, hi = universe;
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (!(u < cdf[mid])) lo = mid + 1; else hi = mid;
    }
    if (lo >= universe) lo = universe - 1;
    return lo + 1;
}
void to_gen_sort_zipf(uint64_t* out, uint64_t begin, uint64_t n, uint64_t seed, const double* cdf, uint64_t universe) {
    for (uint64_t j = 0; j < n; ++j) out[j] = to_zipf_rank(cdf, universe, u01(to_splitmix64(begin + j + seed)));
}
void to_gen_reduce_zipf(to_kv* out, uint64_t begin, uint64_t n, uint64_t seed, const double* cdf, uint64_t universe, int exact) {
    for (uint64_t j = 0; j < n; ++j) {
        uint64_t i = begin + j;
        out[j].key = to_zipf_rank(cdf, universe, u01(to_splitmix64(i + seed)));
        out[j].val = gen_val(i, seed, exact);
    }
}
/* Record{uint8 key[10]; uint8 value[90]} — examples/terasort/terasort.cpp:31-42 */
void to_gen_records(uint8_t* out, uint64_t begin, uint64_t n, uint64_t seed) {
    for (uint64_t j = 0; j < n; ++j) {
        uint64_t i = begin + j;
        uint8_t* r = out + 100 * j;
        uint64_t a = to_splitmix64(2 * i + seed), b = to_splitmix64(2 * i + 1 + seed);
        memcpy(r, &a, 8);
        memcpy(r + 8, &b, 2);
        for (int w = 0; w < 12; ++w) {
            uint64_t v = to_splitmix64(i * 12 + w + (seed << 32));
            memcpy(r + 10 + 8 * w, &v, (w == 11) ? 2 : 8);
        }
    }
}

/* ========================================================================== */
/* hashing — common/hash.hpp:64-73 (Hash128to64), core/reduce_functional.hpp:60-72 (ReduceByHash) */

uint64_t to_hash128to64(uint64_t upper, uint64_t lower) {
    const uint64_t k = 0x9DDFEA08EB382D69ull;
    uint64_t a = (lower ^ upper) * k;
    a ^= (a >> 47);
    uint64_t b = (upper ^ a) * k;
    b ^= (b >> 47);
    b *= k;
    return b;
}

/* hash_function_ = std::hash<uint64_t> = identity in libstdc++ (api/reduce_by_key.hpp:252) */
void to_reduce_by_hash(uint64_t key, uint64_t salt, uint64_t num_partitions,
                       uint64_t* partition_id, uint64_t* remaining_hash) {
    uint64_t hash = to_hash128to64(salt, key);
    *partition_id = hash % num_partitions;
    *remaining_hash = hash / num_partitions;
}

void to_hash_partition_ids(const uint64_t* keys, uint64_t n, uint64_t stride_bytes, uint64_t salt,
                           uint64_t num_partitions, uint32_t* out_partition) {
    const uint8_t* p = (const uint8_t*)keys;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t k;
        memcpy(&k, p + i * stride_bytes, 8);
        out_partition[i] = (uint32_t)(to_hash128to64(salt, k) % num_partitions);
    }
}

/* ========================================================================== */
/* Sort */

/* common/reservoir_sampling.hpp:270-275 with desired_imbalance = 0.1 (api/sort.hpp:298) */
uint64_t to_sample_size(uint64_t count, double imbalance) {
    uint64_t s = (uint64_t)(log2((double)count) * (1.0 / (imbalance * imbalance)));
    return s > 1 ? s : 1;
}

static inline uint64_t load_key_le(const to_key_desc* d, const uint8_t* item) {
    uint64_t k = 0;
    memcpy(&k, item + d->key_offset, d->key_bytes);   /* little-endian host */
    return k;
}

/* the CompareFunction: std::less on an unsigned integer key, or the Record comparator
 * (lexicographical_compare over the key bytes, examples/terasort/terasort.cpp:35-37) */
int to_less(const to_key_desc* d, const void* a, const void* b) {
    if (d->key_kind == TO_KEY_UINT_LE)
        return load_key_le(d, (const uint8_t*)a) < load_key_le(d, (const uint8_t*)b);
    return memcmp((const uint8_t*)a + d->key_offset, (const uint8_t*)b + d->key_offset, d->key_bytes) < 0;
}

/* stable bottom-up merge sort over generic fixed-size items (stride = bytes per element, the key
 * descriptor addresses the item at the start of each element) with an optional u64 tie-break field. */
typedef struct { const to_key_desc* d; size_t stride; int tie_off; } sort_ctx;
static inline int elem_less(const sort_ctx* c, const uint8_t* a, const uint8_t* b) {
    if (to_less(c->d, a, b)) return 1;
    if (c->tie_off >= 0 && !to_less(c->d, b, a)) {
        uint64_t ia, ib;
        memcpy(&ia, a + c->tie_off, 8);
        memcpy(&ib, b + c->tie_off, 8);
        return ia < ib;
    }
    return 0;
}
static void merge_sort(uint8_t* base, uint64_t n, const sort_ctx* c) {
    if (n < 2) return;
    size_t s = c->stride;
    uint8_t* tmp = (uint8_t*)malloc((size_t)n * s);
    uint8_t *src = base, *dst = tmp;
    /* insertion sort runs of 16 */
    for (uint64_t lo = 0; lo < n; lo += 16) {
        uint64_t hi = lo + 16 < n ? lo + 16 : n;
        uint8_t cur[256];
        uint8_t* curp = s <= sizeof(cur) ? cur : (uint8_t*)malloc(s);
        for (uint64_t i = lo + 1; i < hi; ++i) {
            memcpy(curp, src + i * s, s);
            uint64_t j = i;
            while (j > lo && elem_less(c, curp, src + (j - 1) * s)) {
                memcpy(src + j * s, src + (j - 1) * s, s);
                --j;
            }
            memcpy(src + j * s, curp, s);
        }
        if (curp != cur) free(curp);
    }
    for (uint64_t w = 16; w < n; w *= 2) {
        for (uint64_t lo = 0; lo < n; lo += 2 * w) {
            uint64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            uint64_t i = lo, j = mid, o = lo;
            while (i < mid && j < hi) {
                /* take right only if strictly less -> stable */
                if (elem_less(c, src + j * s, src + i * s)) { memcpy(dst + o * s, src + j * s, s); ++j; }
                else { memcpy(dst + o * s, src + i * s, s); ++i; }
                ++o;
            }
            if (i < mid) memcpy(dst + o * s, src + i * s, (size_t)(mid - i) * s), o += mid - i;
            if (j < hi) memcpy(dst + o * s, src + j * s, (size_t)(hi - j) * s);
        }
        uint8_t* t = src; src = dst; dst = t;
    }
    if (src != base) memcpy(base, src, (size_t)n * s);
    free(tmp);
}

static void radix_sort_u64(uint64_t* a, uint64_t n) {
    uint64_t* b = (uint64_t*)malloc((size_t)n * 8);
    uint64_t *src = a, *dst = b;
    for (int pass = 0; pass < 8; ++pass) {
        uint64_t cnt[257] = { 0 };
        int sh = pass * 8;
        for (uint64_t i = 0; i < n; ++i) cnt[((src[i] >> sh) & 255) + 1]++;
        if (cnt[((src[0] >> sh) & 255) + 1] == n) continue;   /* all same digit */
        for (int i = 0; i < 256; ++i) cnt[i + 1] += cnt[i];
        for (uint64_t i = 0; i < n; ++i) dst[cnt[(src[i] >> sh) & 255]++] = src[i];
        uint64_t* t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, (size_t)n * 8);
    free(b);
}

/* DefaultSortAlgorithm = std::sort(begin,end,cmp) (api/sort.hpp:789-796) / std::stable_sort (:862-869).
 * For items that are entirely key the result is unique; otherwise equal keys keep input order here
 * (one valid std::sort outcome and THE std::stable_sort outcome). */
void to_sort_items(void* items, uint64_t n, const to_key_desc* d) {
    if (n < 2) return;
    if (d->key_kind == TO_KEY_UINT_LE && d->item_bytes == 8 && d->key_offset == 0 && d->key_bytes == 8) {
        radix_sort_u64((uint64_t*)items, n);
        return;
    }
    sort_ctx c = { d, d->item_bytes, -1 };
    merge_sort((uint8_t*)items, n, &c);
}

/* FindAndSendSplitters (api/sort.hpp:337-378): sort samples by LessSampleIndex (:419-422), then
 * splitters[i-1] = samples[(size_t)(i * double(S)/double(p))], i = 1..p-1 */
uint64_t to_select_splitters(void* samples, uint64_t nsamples, uint64_t p, const to_key_desc* d, void* out_splitters) {
    if (nsamples == 0) return 0;
    size_t s = d->item_bytes + 8;
    sort_ctx c = { d, s, (int)d->item_bytes };
    merge_sort((uint8_t*)samples, nsamples, &c);
    double splitting_size = (double)nsamples / (double)p;
    for (uint64_t i = 1; i < p; ++i) {
        uint64_t idx = (uint64_t)((double)i * splitting_size);
        memcpy((uint8_t*)out_splitters + (i - 1) * s, (uint8_t*)samples + idx * s, s);
    }
    return p - 1;
}

/* TreeBuilder (api/sort.hpp:380-417): implicit 1-based binary search tree over ssplitter = k-1 splitters */
static void tree_recurse(const to_key_desc* d, const uint8_t* spl, size_t s, uint8_t* tree,
                         int64_t lo, int64_t hi, uint64_t treeidx, uint64_t ssplitter) {
    int64_t mid = lo + (hi - lo) / 2;
    memcpy(tree + treeidx * d->item_bytes, spl + mid * s, d->item_bytes);
    if (2 * treeidx < ssplitter) {
        tree_recurse(d, spl, s, tree, lo, mid, 2 * treeidx + 0, ssplitter);
        tree_recurse(d, spl, s, tree, mid + 1, hi, 2 * treeidx + 1, ssplitter);
    }
}
void to_build_tree(const void* splitters, uint64_t k, const to_key_desc* d, void* tree) {
    uint64_t ssplitter = k - 1;
    if (ssplitter != 0)
        tree_recurse(d, (const uint8_t*)splitters, d->item_bytes + 8, (uint8_t*)tree, 0, (int64_t)ssplitter, 1, ssplitter);
}

/* TransmitItems (api/sort.hpp:434-535): tree descent (:478-482) + equal-to-splitter tie-break by global
 * index, EqualSampleGreaterIndex (:424-426, :487-502).  out_bucket is the bucket b in [0,k). */
void to_classify(const void* items, uint64_t n, uint64_t prefix_items, const void* tree, uint64_t k,
                 uint64_t log_k, const void* splitters, const to_key_desc* d, uint32_t* out_bucket) {
    const uint8_t* it = (const uint8_t*)items;
    const uint8_t* tr = (const uint8_t*)tree;
    const uint8_t* spl = (const uint8_t*)splitters;
    size_t s = d->item_bytes + 8;
    for (uint64_t i = 0; i < n; ++i) {
        const uint8_t* el = it + i * d->item_bytes;
        uint64_t j = 1;
        for (uint64_t l = 0; l < log_k; ++l)
            j = 2 * j + (to_less(d, el, tr + j * d->item_bytes) ? 0 : 1);
        uint64_t b = j - k;
        while (b) {
            const uint8_t* sp = spl + (b - 1) * s;
            uint64_t sidx;
            memcpy(&sidx, sp + d->item_bytes, 8);
            /* EqualSampleGreaterIndex(splitter, (el, i)) = !cmp(spl.key, el) && spl.idx >= i */
            if (!to_less(d, sp, el) && sidx >= prefix_items + i) b--;
            else break;
        }
        out_bucket[i] = (uint32_t)b;
    }
}

/* ---- tlx::LoserTreeCopy (extlib/tlx/tlx/container/loser_tree.hpp:54-292) driven as
 * core::MultiwayMergeTree does (core/multiway_merge.hpp:37-106) ---- */
typedef struct { int sup; uint32_t source; uint8_t* key; } loser;

void to_multiway_merge(const void* const* runs, const uint64_t* run_items, uint32_t num_inputs,
                       const to_key_desc* d, int stable, void* outv) {
    const size_t s = d->item_bytes;
    uint8_t* out = (uint8_t*)outv;
    if (num_inputs == 0) return;
    uint32_t ik = num_inputs, k = 1;
    while (k < ik) k <<= 1;                               /* round_up_to_power_of_two */
    loser* L = (loser*)calloc(2 * (size_t)k, sizeof(loser));
    uint8_t* keys = (uint8_t*)calloc(2 * (size_t)k + 1, s);
    for (uint32_t i = 0; i < 2 * k; ++i) L[i].key = keys + (size_t)i * s;
    uint8_t* tmpkey = keys + (size_t)2 * k * s;
    for (uint32_t i = ik - 1; i < k; ++i) { L[i + k].sup = 1; L[i + k].source = (uint32_t)-1; }   /* :93-96 */
    uint64_t* pos = (uint64_t*)calloc(num_inputs, 8);
    uint64_t remaining = num_inputs;
    int first_insert = 1;
    /* insert_start (:112-134) */
    for (uint32_t t = 0; t < num_inputs; ++t) {
        int has = run_items[t] > 0;
        const uint8_t* keyp = has ? (const uint8_t*)runs[t] : NULL;
        uint32_t p = k + t;
        L[p].sup = !has;
        L[p].source = t;
        if (first_insert) {
            for (uint32_t i = 0; i < 2 * k; ++i) {
                if (keyp) memcpy(L[i].key, keyp, s); else memset(L[i].key, 0, s);
            }
            first_insert = 0;
        }
        else {
            if (keyp) memcpy(L[p].key, keyp, s); else memset(L[p].key, 0, s);
        }
        if (has) pos[t] = 1; else --remaining;
    }
    /* init_winner (:142-161), iterative bottom-up equivalent of the recursion */
    {
        uint32_t* winner = (uint32_t*)malloc(2 * (size_t)k * sizeof(uint32_t));
        for (uint32_t i = k; i < 2 * k; ++i) winner[i] = i;
        for (uint32_t root = k - 1; root >= 1; --root) {
            uint32_t left = winner[2 * root], right = winner[2 * root + 1];
            if (L[right].sup || (!L[left].sup && !to_less(d, L[right].key, L[left].key))) {
                /* left one is less or equal */
                L[root].sup = L[right].sup; L[root].source = L[right].source; memcpy(L[root].key, L[right].key, s);
                winner[root] = left;
            }
            else {
                L[root].sup = L[left].sup; L[root].source = L[left].source; memcpy(L[root].key, L[left].key, s);
                winner[root] = right;
            }
        }
        uint32_t w = (k == 1) ? 1 : winner[1];
        /* init(): losers_[0] = losers_[init_winner(1)] ; with k == 1 init_winner(1) returns 1 */
        L[0].sup = L[w].sup; L[0].source = L[w].source; memcpy(L[0].key, L[w].key, s);
        free(winner);
    }
    uint64_t o = 0;
    while (remaining != 0) {
        uint32_t top = L[0].source;                        /* min_source */
        memcpy(out + o * s, (const uint8_t*)runs[top] + (pos[top] - 1) * s, s);
        ++o;
        int sup;
        if (pos[top] < run_items[top]) {
            memcpy(tmpkey, (const uint8_t*)runs[top] + pos[top] * s, s);
            pos[top]++;
            sup = 0;
        }
        else {
            memset(tmpkey, 0, s);
            sup = 1;
            --remaining;
        }
        /* delete_min_insert (:181-209 unstable, :246-272 stable) */
        uint32_t source = L[0].source;
        uint8_t* key = tmpkey;
        for (uint32_t p = (k + source) / 2; p > 0; p /= 2) {
            int do_swap;
            if (!stable) {
                if (sup) do_swap = 1;                                  /* the other candidate is smaller */
                else if (L[p].sup) do_swap = 0;
                else do_swap = to_less(d, L[p].key, key);
                if (do_swap && !sup && 0) { }
            }
            else {
                do_swap = (sup && (!L[p].sup || L[p].source < source)) ||
                          (!sup && !L[p].sup &&
                           (to_less(d, L[p].key, key) ||
                            (!to_less(d, key, L[p].key) && L[p].source < source)));
            }
            if (do_swap) {
                int ts = L[p].sup; L[p].sup = sup; sup = ts;
                uint32_t tsrc = L[p].source; L[p].source = source; source = tsrc;
                /* swap key contents through the spare buffer */
                uint8_t buf[256];
                uint8_t* bp = s <= sizeof(buf) ? buf : (uint8_t*)malloc(s);
                memcpy(bp, L[p].key, s); memcpy(L[p].key, key, s); memcpy(key, bp, s);
                if (bp != buf) free(bp);
            }
        }
        L[0].sup = sup; L[0].source = source; memcpy(L[0].key, key, s);
    }
    free(pos); free(keys); free(L);
}

/* SortNode::MainOp on p simulated workers (api/sort.hpp:537-663, :151-175 sampling, :665-742 local sort) */
void to_sort_operator(const void* inv, const uint64_t* local_counts, uint32_t p, const to_key_desc* d,
                      int stable, uint64_t rng_seed, void* outv, uint64_t* out_counts) {
    (void)stable;   /* the local sort below is stable and sources are drained in worker order: valid for both */
    const uint8_t* in = (const uint8_t*)inv;
    uint8_t* out = (uint8_t*)outv;
    const size_t s = d->item_bytes, ss = s + 8;
    uint64_t total = 0;
    uint64_t* prefix = (uint64_t*)calloc(p + 1, 8);
    for (uint32_t w = 0; w < p; ++w) { prefix[w] = total; total += local_counts[w]; }   /* ExPrefixSumTotal :541 */
    prefix[p] = total;
    for (uint32_t w = 0; w < p; ++w) out_counts[w] = 0;
    if (total == 0) { free(prefix); return; }
    /* OnPreOpFile sampling (:151-175): pick_items = min(local, wanted) indices rng() % local_items */
    uint64_t nsamples = 0;
    for (uint32_t w = 0; w < p; ++w)
        if (local_counts[w]) {
            uint64_t want = to_sample_size(local_counts[w], 0.1);
            nsamples += want < local_counts[w] ? want : local_counts[w];
        }
    uint8_t* samples = (uint8_t*)malloc((size_t)nsamples * ss);
    uint64_t si = 0, rng = rng_seed;
    for (uint32_t w = 0; w < p; ++w) {
        if (!local_counts[w]) continue;
        uint64_t want = to_sample_size(local_counts[w], 0.1);
        uint64_t pick = want < local_counts[w] ? want : local_counts[w];
        for (uint64_t i = 0; i < pick; ++i) {
            uint64_t index = to_splitmix64(rng++) % local_counts[w];
            uint64_t gidx = prefix[w] + index;                                 /* :566-569 adds the prefix */
            memcpy(samples + si * ss, in + gidx * s, s);
            memcpy(samples + si * ss + s, &gidx, 8);
            ++si;
        }
    }
    uint32_t ceil_log = 0;
    while ((1u << ceil_log) < p) ++ceil_log;                                    /* integer_log2_ceil :575 */
    uint64_t k = 1ull << ceil_log;
    uint8_t* splitters = (uint8_t*)calloc(k + 1, ss);
    uint64_t ns = to_select_splitters(samples, nsamples, p, d, splitters);
    for (uint64_t i = p; i < k; ++i) { memcpy(splitters + ns * ss, splitters + (ns - 1) * ss, ss); ++ns; }   /* :607-609 */
    uint8_t* tree = (uint8_t*)calloc(k + 1, s);
    to_build_tree(splitters, k, d, tree);
    uint32_t* bucket = (uint32_t*)malloc((size_t)total * 4);
    for (uint32_t w = 0; w < p; ++w)
        to_classify(in + prefix[w] * s, local_counts[w], prefix[w], tree, k, ceil_log, splitters, d, bucket + prefix[w]);
    /* std::swap(data_writers[actual_k-1], data_writers[k-1]) (:460): bucket k-1 goes to worker p-1 */
    for (uint64_t i = 0; i < total; ++i) {
        uint32_t b = bucket[i];
        if (b == k - 1) b = p - 1;
        bucket[i] = b;
        out_counts[b]++;
    }
    uint64_t* off = (uint64_t*)calloc(p + 1, 8);
    for (uint32_t w = 0; w < p; ++w) off[w + 1] = off[w] + out_counts[w];
    uint64_t* cur = (uint64_t*)malloc(p * 8);
    memcpy(cur, off, p * 8);
    for (uint64_t i = 0; i < total; ++i) { memcpy(out + cur[bucket[i]] * s, in + i * s, s); cur[bucket[i]]++; }
    for (uint32_t w = 0; w < p; ++w) to_sort_items(out + off[w] * s, out_counts[w], d);     /* :696-742 */
    free(cur); free(off); free(bucket); free(tree); free(splitters); free(samples); free(prefix);
}

/* ========================================================================== */
/* ReduceByKey: core::ReduceProbingHashTable (core/reduce_probing_hash_table.hpp) for
 * TableItem = pair<uint64_t, 8-byte value>, Key() == 0 is the empty-slot sentinel. */

typedef struct { to_kv* v; uint64_t n, cap; } kv_vec;
static void kv_push(kv_vec* f, const to_kv* it) {
    if (f->n == f->cap) { f->cap = f->cap ? 2 * f->cap : 64; f->v = (to_kv*)realloc(f->v, f->cap * sizeof(to_kv)); }
    f->v[f->n++] = *it;
}

struct to_table {
    uint64_t num_partitions, limit_memory_bytes;
    int immediate_flush, op;
    uint64_t salt;
    to_emit_fn emit; void* emit_ctx;
    uint64_t num_buckets_per_partition, num_buckets, num_items;
    to_kv* items;
    uint64_t *partition_size, *limit_items, *items_per_partition;
    uint64_t sentinel_partition;
    kv_vec* partition_files;     /* spill Files (core/reduce_table.hpp:131-135) */
    double fill_rate;
};
#define INVALID_PARTITION ((uint64_t)-1)

static inline uint64_t apply_op(int op, uint64_t a, uint64_t b) {
    switch (op) {
    case TO_OP_SUM_F64: return dbits(bitsd(a) + bitsd(b));
    case TO_OP_SUM_U64: return a + b;
    case TO_OP_MIN_U64: return a < b ? a : b;
    case TO_OP_MAX_U64: return a > b ? a : b;
    case TO_OP_MIN_F64: return dbits(bitsd(b) < bitsd(a) ? bitsd(b) : bitsd(a));   /* std::min(a,b) */
    case TO_OP_MAX_F64: return dbits(bitsd(a) < bitsd(b) ? bitsd(b) : bitsd(a));   /* std::max(a,b) */
    default: return a;   /* FIRST: keep the value already in the table (benchmarks/hashtable/reduce.cpp:51-54) */
    }
}

/* Initialize (:111-168) with DefaultReduceConfig (core/reduce_table.hpp:40-80): fill rate 0.5, 512 initial */
to_table* to_table_new(uint64_t num_partitions, uint64_t limit_memory_bytes, int immediate_flush,
                       uint64_t salt, int op, to_emit_fn emit, void* emit_ctx) {
    to_table* t = (to_table*)calloc(1, sizeof(to_table));
    t->num_partitions = num_partitions; t->limit_memory_bytes = limit_memory_bytes;
    t->immediate_flush = immediate_flush; t->op = op; t->salt = salt; t->emit = emit; t->emit_ctx = emit_ctx;
    t->fill_rate = 0.5;
    uint64_t nbpp = (uint64_t)((double)limit_memory_bytes / (double)sizeof(to_kv) / (double)num_partitions);
    if (nbpp < 1) nbpp = 1;
    t->num_buckets_per_partition = nbpp;
    t->num_buckets = nbpp * num_partitions;
    t->partition_size = (uint64_t*)malloc(num_partitions * 8);
    t->limit_items = (uint64_t*)malloc(num_partitions * 8);
    t->items_per_partition = (uint64_t*)calloc(num_partitions, 8);
    uint64_t init = nbpp < 512 ? nbpp : 512;
    for (uint64_t i = 0; i < num_partitions; ++i) {
        t->partition_size[i] = init;
        t->limit_items[i] = (uint64_t)((double)init * t->fill_rate);
    }
    t->items = (to_kv*)malloc((t->num_buckets + 1) * sizeof(to_kv));   /* + 1 sentinel slot */
    for (uint64_t id = 0; id < num_partitions; ++id)
        memset(t->items + id * nbpp, 0, init * sizeof(to_kv));
    t->sentinel_partition = INVALID_PARTITION;
    t->partition_files = immediate_flush ? NULL : (kv_vec*)calloc(num_partitions, sizeof(kv_vec));
    return t;
}
void to_table_free(to_table* t) {
    if (!t) return;
    if (t->partition_files) { for (uint64_t i = 0; i < t->num_partitions; ++i) free(t->partition_files[i].v); free(t->partition_files); }
    free(t->items); free(t->partition_size); free(t->limit_items); free(t->items_per_partition); free(t);
}
uint64_t to_table_num_items(const to_table* t) { return t->num_items; }
uint64_t to_table_partition_size(const to_table* t, uint64_t partition) { return t->partition_size[partition]; }

static void table_grow_and_rehash(to_table* t, uint64_t pid);

/* GrowPartition (:337-366) — mem::memory_exceeded is never set in the oracle */
static void table_grow_partition(to_table* t, uint64_t pid) {
    if (t->partition_size[pid] == t->num_buckets_per_partition) return;
    uint64_t new_size = 2 * t->partition_size[pid];
    if (new_size > t->num_buckets_per_partition) new_size = t->num_buckets_per_partition;
    to_kv* pbegin = t->items + pid * t->num_buckets_per_partition;
    memset(pbegin + t->partition_size[pid], 0, (new_size - t->partition_size[pid]) * sizeof(to_kv));
    t->partition_size[pid] = new_size;
    t->limit_items[pid] = (uint64_t)((double)new_size * t->fill_rate);
}

/* FlushPartitionEmit (:443-482) with the table's emitter, consume = true */
static void table_flush_partition(to_table* t, uint64_t pid, int consume, int grow) {
    if (t->sentinel_partition == pid) {
        t->emit(t->emit_ctx, pid, &t->items[t->num_buckets]);
        if (consume) t->sentinel_partition = INVALID_PARTITION;
    }
    to_kv* iter = t->items + pid * t->num_buckets_per_partition;
    to_kv* pend = iter + t->partition_size[pid];
    for ( ; iter != pend; ++iter) {
        if (iter->key != 0) {
            t->emit(t->emit_ctx, pid, iter);
            if (consume) { iter->key = 0; iter->val = 0; }
        }
    }
    if (consume) { t->num_items -= t->items_per_partition[pid]; t->items_per_partition[pid] = 0; }
    if (grow) table_grow_partition(t, pid);
}

/* SpillPartition (:372-409) */
static void table_spill_partition(to_table* t, uint64_t pid) {
    if (t->immediate_flush) { table_flush_partition(t, pid, 1, 1); return; }
    if (t->items_per_partition[pid] == 0) return;
    kv_vec* f = &t->partition_files[pid];
    if (t->sentinel_partition == pid) { kv_push(f, &t->items[t->num_buckets]); t->sentinel_partition = INVALID_PARTITION; }
    to_kv* iter = t->items + pid * t->num_buckets_per_partition;
    to_kv* pend = iter + t->partition_size[pid];
    for ( ; iter != pend; ++iter)
        if (iter->key != 0) { kv_push(f, iter); iter->key = 0; iter->val = 0; }
    t->num_items -= t->items_per_partition[pid];
    t->items_per_partition[pid] = 0;
}

/* Insert (:190-268) */
int to_table_insert(to_table* t, const to_kv* kvp) {
    to_kv kv = *kvp;
    uint64_t pid, rem;
    to_reduce_by_hash(kv.key, t->salt, t->num_partitions, &pid, &rem);
    if (kv.key == 0) {
        to_kv* sentinel = &t->items[t->num_buckets];
        if (t->sentinel_partition == INVALID_PARTITION) { *sentinel = kv; t->sentinel_partition = pid; }
        else { sentinel->val = apply_op(t->op, sentinel->val, kv.val); return 0; }
        ++t->items_per_partition[pid];
        ++t->num_items;
        while (t->items_per_partition[pid] > t->limit_items[pid]) table_grow_and_rehash(t, pid);
        return 1;
    }
    uint64_t local_index = rem % t->partition_size[pid];
    to_kv* pbegin = t->items + pid * t->num_buckets_per_partition;
    to_kv* pend = pbegin + t->partition_size[pid];
    to_kv* begin_iter = pbegin + local_index;
    to_kv* iter = begin_iter;
    while (iter->key != 0) {
        if (iter->key == kv.key) { iter->val = apply_op(t->op, iter->val, kv.val); return 0; }
        ++iter;
        if (iter == pend) iter = pbegin;
        if (iter == begin_iter) { table_grow_and_rehash(t, pid); return to_table_insert(t, &kv); }
    }
    *iter = kv;
    ++t->items_per_partition[pid];
    ++t->num_items;
    while (t->items_per_partition[pid] >= t->limit_items[pid]) {
        uint64_t before = t->items_per_partition[pid];
        table_grow_and_rehash(t, pid);
        /* degenerate limit 0 (1-slot partitions) spins forever in the reference; stop instead */
        if (t->items_per_partition[pid] == 0 && before == 0) break;
    }
    return 1;
}

/* GrowAndRehash (:293-333) */
static void table_grow_and_rehash(to_table* t, uint64_t pid) {
    uint64_t old_size = t->partition_size[pid];
    table_grow_partition(t, pid);
    if (t->partition_size[pid] == old_size) { table_spill_partition(t, pid); return; }
    if (t->partition_size[pid] % old_size != 0) { table_spill_partition(t, pid); return; }
    to_kv* pbegin = t->items + pid * t->num_buckets_per_partition;
    to_kv* iter = pbegin;
    to_kv* pend = pbegin + old_size;
    int passed_first_half = 0, found_hole = 0;
    while (!passed_first_half || !found_hole) {
        int is_empty = (iter->key == 0);
        if (!is_empty) {
            --t->items_per_partition[pid];
            --t->num_items;
            to_kv item = *iter;
            iter->key = 0; iter->val = 0;
            to_table_insert(t, &item);
        }
        iter++;
        found_hole = passed_first_half && is_empty;
        passed_first_half = passed_first_half || iter == pend;
    }
}

void to_table_flush_all(to_table* t) {            /* FlushAll (:484-488) */
    for (uint64_t i = 0; i < t->num_partitions; ++i) table_flush_partition(t, i, 1, 0);
}

typedef struct { to_kv* items; uint32_t* part; uint64_t n, cap; } collect_ctx;
static void collect_emit(void* c, uint64_t pid, const to_kv* it) {
    collect_ctx* cc = (collect_ctx*)c;
    if (cc->n < cc->cap) { cc->items[cc->n] = *it; if (cc->part) cc->part[cc->n] = (uint32_t)pid; }
    cc->n++;
}

/* ReducePrePhase (core/reduce_pre_phase.hpp:103-201): p partitions, immediate_flush = true (:144),
 * Insert all, FlushAll (:179-183).  Emitted items go to worker `partition` (:57-61). */
uint64_t to_reduce_pre_phase(const to_kv* in, uint64_t n, uint64_t p, uint64_t limit_memory_bytes, int op,
                             to_kv* out_items, uint32_t* out_part, uint64_t out_capacity) {
    collect_ctx cc = { out_items, out_part, 0, out_capacity };
    to_table* t = to_table_new(p, limit_memory_bytes, 1, 0, op, collect_emit, &cc);
    for (uint64_t i = 0; i < n; ++i) to_table_insert(t, &in[i]);
    to_table_flush_all(t);
    to_table_free(t);
    return cc.n;
}

/* ReduceByHashPostPhase (core/reduce_by_hash_post_phase.hpp:44-281): 32 partitions (:77), spill to Files,
 * Flush (:95-225) re-reduces spilled partitions with IndexFunction(iteration, ...) (:166) */
uint64_t to_reduce_post_phase(const to_kv* in, uint64_t n, uint64_t limit_memory_bytes, int op,
                              to_kv* out_items, uint64_t out_capacity, uint64_t* out_reduce_iterations) {
    collect_ctx cc = { out_items, NULL, 0, out_capacity };
    to_table* t = to_table_new(32, limit_memory_bytes, 0, 0, op, collect_emit, &cc);
    for (uint64_t i = 0; i < n; ++i) to_table_insert(t, &in[i]);
    kv_vec* remaining = NULL; uint64_t nrem = 0;
    for (uint64_t id = 0; id < 32; ++id) {
        if (t->partition_files[id].n > 0) {
            table_spill_partition(t, id);
            remaining = (kv_vec*)realloc(remaining, (nrem + 1) * sizeof(kv_vec));
            remaining[nrem++] = t->partition_files[id];
            memset(&t->partition_files[id], 0, sizeof(kv_vec));
        }
        else table_flush_partition(t, id, 1, 0);
    }
    to_table_free(t);
    uint64_t iteration = 1;
    while (nrem) {
        kv_vec* next = NULL; uint64_t nnext = 0;
        to_table* sub = to_table_new(32, limit_memory_bytes, 0, iteration, op, collect_emit, &cc);
        for (uint64_t f = 0; f < nrem; ++f) {
            for (uint64_t i = 0; i < remaining[f].n; ++i) to_table_insert(sub, &remaining[f].v[i]);
            free(remaining[f].v);
            for (uint64_t id = 0; id < 32; ++id) {
                if (sub->partition_files[id].n > 0) {
                    table_spill_partition(sub, id);
                    next = (kv_vec*)realloc(next, (nnext + 1) * sizeof(kv_vec));
                    next[nnext++] = sub->partition_files[id];
                    memset(&sub->partition_files[id], 0, sizeof(kv_vec));
                }
                else table_flush_partition(sub, id, 1, 0);
            }
        }
        to_table_free(sub);
        free(remaining);
        remaining = next; nrem = nnext;
        ++iteration;
    }
    if (out_reduce_iterations) *out_reduce_iterations = iteration - 1;
    return cc.n;
}

/* ReduceNode on p simulated workers (api/reduce_by_key.hpp:100-211): pre phase with mem/2 per worker,
 * exchange by partition id, post phase with mem/2 (:142-155).  Arrival order over a MixStream is arbitrary;
 * here sources are drained in worker order. */
uint64_t to_reduce_operator(const to_kv* in, const uint64_t* local_counts, uint32_t p,
                            uint64_t mem_limit_bytes, int op, to_kv* out, uint64_t* out_counts) {
    uint64_t total = 0;
    for (uint32_t w = 0; w < p; ++w) total += local_counts[w];
    to_kv* pre_items = (to_kv*)malloc((size_t)(total + 1) * sizeof(to_kv));
    uint32_t* pre_part = (uint32_t*)malloc((size_t)(total + 1) * 4);
    uint64_t npre = 0, base = 0;
    for (uint32_t w = 0; w < p; ++w) {
        npre += to_reduce_pre_phase(in + base, local_counts[w], p, mem_limit_bytes / 2, op,
                                    pre_items + npre, pre_part + npre, total - npre);
        base += local_counts[w];
    }
    uint64_t* cnt = (uint64_t*)calloc(p + 1, 8);
    for (uint64_t i = 0; i < npre; ++i) cnt[pre_part[i] + 1]++;
    for (uint32_t w = 0; w < p; ++w) cnt[w + 1] += cnt[w];
    to_kv* recv = (to_kv*)malloc((size_t)(npre + 1) * sizeof(to_kv));
    uint64_t* cur = (uint64_t*)malloc(p * 8);
    memcpy(cur, cnt, p * 8);
    for (uint64_t i = 0; i < npre; ++i) recv[cur[pre_part[i]]++] = pre_items[i];
    uint64_t nout = 0;
    for (uint32_t w = 0; w < p; ++w) {
        uint64_t got = to_reduce_post_phase(recv + cnt[w], cnt[w + 1] - cnt[w], mem_limit_bytes / 2, op,
                                            out + nout, total - nout, NULL);
        out_counts[w] = got;
        nout += got;
    }
    free(cur); free(recv); free(cnt); free(pre_part); free(pre_items);
    return nout;
}

/* scalable checker: group equal keys (stable sort by key) and fold in input order */
uint64_t to_reduce_simple(const to_kv* in, uint64_t n, int op, to_kv* out) {
    if (n == 0) return 0;
    to_kv* tmp = (to_kv*)malloc((size_t)n * sizeof(to_kv));
    to_kv* buf = (to_kv*)malloc((size_t)n * sizeof(to_kv));
    memcpy(tmp, in, (size_t)n * sizeof(to_kv));
    to_kv *src = tmp, *dst = buf;
    for (int pass = 0; pass < 8; ++pass) {                 /* LSD radix on the key, stable */
        uint64_t c[257] = { 0 };
        int sh = pass * 8;
        for (uint64_t i = 0; i < n; ++i) c[((src[i].key >> sh) & 255) + 1]++;
        if (c[((src[0].key >> sh) & 255) + 1] == n) continue;
        for (int i = 0; i < 256; ++i) c[i + 1] += c[i];
        for (uint64_t i = 0; i < n; ++i) dst[c[(src[i].key >> sh) & 255]++] = src[i];
        to_kv* t = src; src = dst; dst = t;
    }
    uint64_t o = 0;
    out[0] = src[0];
    for (uint64_t i = 1; i < n; ++i) {
        if (src[i].key == out[o].key) out[o].val = apply_op(op, out[o].val, src[i].val);
        else out[++o] = src[i];
    }
    free(tmp); free(buf);
    return o + 1;
}

/* ReduceToIndex (api/reduce_to_index.hpp:60-237): the result has `size` items; item i is the fold (in input order) of
 * all items whose index (key) is i, or the neutral element where there is none — ReduceByIndexPostPhase fills its
 * dense table with neutral_element_ and folds the arrivals into it (core/reduce_by_index_post_phase.hpp:141-222); worker w
 * holds a contiguous index range and the ranges are ordered, so the concatenation over workers is the dense array.
 * Returns the number of items with an index >= size (an error for the caller: the reference asserts). */
uint64_t to_reduce_to_index(const to_kv* in, uint64_t n, uint64_t size, to_kv neutral, int op, to_kv* out) {
    unsigned char* seen = (unsigned char*)calloc((size_t)(size ? size : 1), 1);
    uint64_t bad = 0;
    for (uint64_t i = 0; i < size; ++i) out[i] = neutral;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t k = in[i].key;
        if (k >= size) { ++bad; continue; }
        if (!seen[k]) { out[k] = in[i]; seen[k] = 1; }
        else out[k].val = apply_op(op, out[k].val, in[i].val);
    }
    free(seen);
    return bad;
}

/* ========================================================================== */
/* data::File layout of fixed-size items as written by BlockWriter (data/block_writer.hpp:61-67 block_size_
 * = min(start_block_size, max), :405-420 AllocateBlock doubles while 2*bs < max; :183-203 item starts are
 * counted in the block where their first byte lands; items straddle blocks via Append :339-372) */
uint64_t to_file_layout(uint64_t num_items, uint32_t item_bytes, uint64_t start_block_size,
                        uint64_t max_block_size, to_block_meta* out, uint64_t out_capacity) {
    uint64_t block_size = start_block_size < max_block_size ? start_block_size : max_block_size;
    uint64_t nblocks = 0;
    uint64_t cur = 0, end = 0, nitems = 0, first_off = 0;   /* state of the block being filled */
    int have = 0;
    for (uint64_t i = 0; i < num_items; ++i) {
        if (!have || cur == end) {                           /* Put: if (current_ == end_) Flush(), AllocateBlock() */
            if (have) {
                if (nblocks < out_capacity) { out[nblocks].begin = 0; out[nblocks].end = cur; out[nblocks].first_item = first_off; out[nblocks].num_items = nitems; }
                ++nblocks;
            }
            end = block_size; cur = 0; nitems = 0; first_off = 0; have = 1;
            if (2 * block_size < max_block_size) block_size *= 2;
        }
        if (nitems == 0) first_off = cur;
        ++nitems;
        uint64_t left = item_bytes;
        while (left) {                                       /* Append: fill, flush, allocate, continue */
            uint64_t room = end - cur;
            if (room == 0) {
                if (nblocks < out_capacity) { out[nblocks].begin = 0; out[nblocks].end = cur; out[nblocks].first_item = first_off; out[nblocks].num_items = nitems; }
                ++nblocks;
                end = block_size; cur = 0; nitems = 0; first_off = 0;
                if (2 * block_size < max_block_size) block_size *= 2;
                room = end;
            }
            uint64_t take = left < room ? left : room;
            cur += take; left -= take;
        }
    }
    if (have && cur > 0) {                                   /* Close(): flush the last partially filled block */
        if (nblocks < out_capacity) { out[nblocks].begin = 0; out[nblocks].end = cur; out[nblocks].first_item = first_off; out[nblocks].num_items = nitems; }
        ++nblocks;
    }
    return nblocks;
}
