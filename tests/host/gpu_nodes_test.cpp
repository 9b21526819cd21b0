/*******************************************************************************
 * tests/host/gpu_nodes_test.cpp — integration test of the drop-in INSIDE the unmodified reference.
 *
 * A real Thrill job (api::Run, mock network, THRILL_WORKERS_PER_HOST = number of GPUs): the same DIAs go
 * through the stock CPU operators (dia.Sort(), dia.ReducePair()) and through the GPU nodes of
 * thrill_b200/host/thrill_gpu_nodes.hpp (thrill_gpu::Sort, thrill_gpu::ReducePair); the gathered results must
 * be identical (bit-exact order for Sort; same key set and exact sums for ReducePair in the integer-valued
 * double mode).  Mirrors tests/api/sort_node_test.cpp and tests/api/reduce_node_test.cpp of the reference.
 * Links the reference library built by oracle/ref/Makefile (it IS the rest of the pipeline) and
 * libthrill_gpu.so.  Prints "PASS ..." lines and exits non-zero on any mismatch.
 ******************************************************************************/
#include <thrill/api/all_gather.hpp>
#include <thrill/api/cache.hpp>
#include <thrill/api/collapse.hpp>
#include <thrill/api/generate.hpp>
#include <thrill/api/reduce_by_key.hpp>
#include <thrill/api/reduce_to_index.hpp>
#include <thrill/api/size.hpp>
#include <thrill/api/sort.hpp>
#include <thrill/api/zip.hpp>
#include <thrill/common/stats_timer.hpp>

#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "../../thrill_b200/host/thrill_gpu_nodes.hpp"

using namespace thrill; // NOLINT

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static std::atomic<int> g_failures { 0 };

//! PageRank as in examples/page_rank/page_rank.hpp:70-139 (cfg5's pipeline: Zip + FlatMap on the reference's CPU operators,
//! the ReduceToIndex step either the stock one or the GPU node); returns the gathered ranks
template <bool UseGpu>
static std::vector<double> PageRankRun(api::Context& ctx, size_t num_pages, size_t iterations) {
    using PageId = uint64_t;
    using OutgoingLinks = std::vector<PageId>;
    using OutgoingLinksRank = std::pair<OutgoingLinks, double>;
    using Contrib = std::pair<PageId, double>;
    const double n_d = static_cast<double>(num_pages);
    auto links = api::Generate(ctx, num_pages, [num_pages](size_t i) {
                                   OutgoingLinks ol;
                                   size_t deg = (i % 17 == 0) ? 0 : 1 + splitmix64(i) % 8;       // some pages have no out-links
                                   for (size_t j = 0; j < deg; ++j) {
                                       uint64_t r = splitmix64(i * 16 + j + 1);
                                       ol.push_back((r % 4 == 0) ? r % 64 : r % num_pages);      // a few popular pages
                                   }
                                   return ol;
                               }).Cache();
    api::DIA<double> ranks = api::Generate(ctx, num_pages, [n_d](size_t) { return 1.0 / n_d; }).Collapse();
    for (size_t iter = 0; iter < iterations; ++iter) {
        auto outs_rank = links.Zip(ranks, [](const OutgoingLinks& ol, const double& r) { return OutgoingLinksRank(ol, r); });
        auto contribs = outs_rank.template FlatMap<Contrib>(
            [](const OutgoingLinksRank& p, auto emit) {
                if (p.first.size() == 0) return;
                double c = p.second / static_cast<double>(p.first.size());
                for (const PageId& tgt : p.first) emit(Contrib(tgt, c));
            });
        auto dampen = [n_d](const Contrib& p) { return 0.85 * p.second + (1 - 0.85) / n_d; };
        if (UseGpu)
            ranks = thrill_gpu::ReduceToIndex(contribs, std::plus<double>(), num_pages).Map(dampen).Collapse();
        else
            ranks = contribs.ReduceToIndex(
                [](const Contrib& p) { return static_cast<size_t>(p.first); },
                [](const Contrib& a, const Contrib& b) { return Contrib(a.first, a.second + b.second); }, num_pages)
                    .Map(dampen).Collapse();
    }
    return ranks.AllGather();
}

//! TeraSort's item (examples/terasort/terasort.cpp:31-42): 10 key bytes compared lexicographically, 90 payload bytes
struct Record {
    uint8_t key[10];
    uint8_t value[90];
    bool operator < (const Record& b) const { return std::lexicographical_compare(key, key + 10, b.key, b.key + 10); }
    bool operator == (const Record& b) const { return memcmp(this, &b, sizeof(Record)) == 0; }
} TLX_ATTRIBUTE_PACKED;
static_assert(sizeof(Record) == 100, "struct Record packing incorrect.");
namespace thrill_gpu {
template <>
struct ByteKeyTraits<Record>{
    static constexpr bool is_byte_key = true;
    static constexpr uint32_t key_offset = 0, key_bytes = 10;
};
} // namespace thrill_gpu

int main(int argc, char** argv) {
    size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 2000000;
    int rc = api::Run(
        [&](api::Context& ctx) {
            // ---- Sort: uniform keys, then heavy duplicates (tie-break path), then empty ----
            for (int variant = 0; variant < 3; ++variant) {
                size_t nn = variant == 2 ? 0 : n;
                auto input = api::Generate(ctx, nn, [variant](size_t i) -> uint64_t {
                                               uint64_t r = splitmix64(i + 42);
                                               return variant == 1 ? r % 7 : r;
                                           }).Cache().Keep(2);
                common::StatsTimerStart t_cpu;
                std::vector<uint64_t> cpu = input.Sort().AllGather();
                t_cpu.Stop();
                common::StatsTimerStart t_gpu;
                std::vector<uint64_t> gpu = thrill_gpu::Sort(input).AllGather();
                t_gpu.Stop();
                bool ok = cpu == gpu && cpu.size() == nn;
                if (ctx.my_rank() == 0)
                    printf("%s Sort variant=%d n=%zu workers=%zu cpu=%.3fs gpu=%.3fs\n", ok ? "PASS" : "FAIL",
                           variant, nn, ctx.num_workers(), t_cpu.SecondsDouble(), t_gpu.SecondsDouble());
                if (!ok) g_failures++;
            }
            // ---- ReducePair<uint64_t,double>(plus), integer-valued doubles => exact sums; includes key 0 ----
            {
                using Pair = std::pair<uint64_t, double>;
                auto input = api::Generate(ctx, n, [](size_t i) {
                                               uint64_t r = splitmix64(i + 7);
                                               return Pair(r % 50021, static_cast<double>(splitmix64(i) % 1024));
                                           }).Cache().Keep(2);
                auto by_key = [](const Pair& a, const Pair& b) { return a.first < b.first; };
                std::vector<Pair> cpu = input.ReducePair(std::plus<double>()).AllGather();
                std::vector<Pair> gpu = thrill_gpu::ReducePair(input, std::plus<double>()).AllGather();
                std::sort(cpu.begin(), cpu.end(), by_key);
                std::sort(gpu.begin(), gpu.end(), by_key);
                bool ok = cpu == gpu;
                if (ctx.my_rank() == 0)
                    printf("%s ReducePair n=%zu distinct=%zu workers=%zu\n", ok ? "PASS" : "FAIL", n, cpu.size(), ctx.num_workers());
                if (!ok) g_failures++;
            }
            // ---- ReduceToIndex (the PageRank step): dense result, missing indices keep the neutral element ----
            {
                using Pair = std::pair<uint64_t, double>;
                const size_t size = 30011;
                auto input = api::Generate(ctx, n / 4, [size](size_t i) {
                                               uint64_t r = splitmix64(i + 99);
                                               return Pair(r % size / 3 * 3, static_cast<double>(splitmix64(i + 5) % 1024));
                                           }).Cache().Keep(2);
                std::vector<Pair> cpu = input.ReduceToIndex(
                    [](const Pair& p) { return static_cast<size_t>(p.first); },
                    [](const Pair& a, const Pair& b) { return Pair(a.first, a.second + b.second); }, size).AllGather();
                std::vector<Pair> gpu = thrill_gpu::ReduceToIndex(input, std::plus<double>(), size).AllGather();
                bool ok = cpu == gpu && cpu.size() == size;
                if (ctx.my_rank() == 0) printf("%s ReduceToIndex n=%zu size=%zu workers=%zu\n", ok ? "PASS" : "FAIL", n / 4, size, ctx.num_workers());
                if (!ok) g_failures++;
            }
            // ---- Sort feeding ReducePair (GPU node -> GPU node through PushFile / OnPreOpFile) ----
            {
                using Pair = std::pair<uint64_t, uint64_t>;
                auto pairs = api::Generate(ctx, n, [](size_t i) { return Pair(splitmix64(i) % 1000, i % 13); }).Cache().Keep(2);
                std::vector<Pair> cpu = pairs.ReducePair(std::plus<uint64_t>()).AllGather();
                // PCIe traffic of this worker's ctx: the chain uploads the input once and downloads the (small) result once;
                // nothing crosses the bus between the two GPU nodes (the Sort result stays in HBM as a device File)
                uint64_t h0 = 0, d0 = 0, h1 = 0, d1 = 0;
                tg_transfer_bytes(thrill_gpu::WorkerCtx(ctx), &h0, &d0);
                std::vector<Pair> gpu = thrill_gpu::ReducePair(
                    thrill_gpu::Sort(pairs, thrill_gpu::LessFirst()), std::plus<uint64_t>()).AllGather();
                tg_transfer_bytes(thrill_gpu::WorkerCtx(ctx), &h1, &d1);
                std::sort(cpu.begin(), cpu.end());
                std::sort(gpu.begin(), gpu.end());
                const uint64_t my_in = (h1 - h0), my_out = (d1 - d0);
                // uploaded exactly this worker's share of the input (16 bytes per item), downloaded at most the reduced result
                bool lean = my_in <= 16 * (n / ctx.num_workers() + 1) && my_out <= 16 * 1000 + 4096;
                bool ok = cpu == gpu && lean;
                if (ctx.my_rank() == 0)
                    printf("%s SortStable->ReducePair chain n=%zu (worker 0: %llu bytes H2D, %llu bytes D2H: device-resident between the nodes)\n",
                           ok ? "PASS" : "FAIL", n, (unsigned long long)my_in, (unsigned long long)my_out);
                if (!ok) g_failures++;
            }
            // ---- SortStable of pairs by key: identical to the stock SortStable (equal keys in global input order) ----
            {
                using Pair = std::pair<uint64_t, uint64_t>;
                auto pairs = api::Generate(ctx, n, [](size_t i) { return Pair(splitmix64(i + 3) % 997, i); }).Cache().Keep(2);
                std::vector<Pair> cpu = pairs.SortStable([](const Pair& a, const Pair& b) { return a.first < b.first; }).AllGather();
                std::vector<Pair> gpu = thrill_gpu::SortStable(pairs, thrill_gpu::LessFirst()).AllGather();
                bool ok = cpu == gpu && cpu.size() == n;
                if (ctx.my_rank() == 0) printf("%s SortStable pairs n=%zu\n", ok ? "PASS" : "FAIL", n);
                if (!ok) g_failures++;
            }
            // ---- ReduceByKey(key_extractor, reduce_function) front door with the recognised functor pair ----
            {
                using Pair = std::pair<uint64_t, uint64_t>;
                auto pairs = api::Generate(ctx, n, [](size_t i) { return Pair(splitmix64(i + 11) % 4099, i % 17); }).Cache().Keep(2);
                std::vector<Pair> cpu = pairs.ReduceByKey(
                    [](const Pair& p) { return p.first; },
                    [](const Pair& a, const Pair& b) { return Pair(a.first, a.second + b.second); }).AllGather();
                std::vector<Pair> gpu = thrill_gpu::ReduceByKey(
                    pairs, thrill_gpu::KeyFirst(), thrill_gpu::OnSecond<std::plus<uint64_t> >()).AllGather();
                std::sort(cpu.begin(), cpu.end());
                std::sort(gpu.begin(), gpu.end());
                bool ok = cpu == gpu;
                if (ctx.my_rank() == 0) printf("%s ReduceByKey(KeyFirst, OnSecond<plus>) n=%zu distinct=%zu\n", ok ? "PASS" : "FAIL", n, cpu.size());
                if (!ok) g_failures++;
            }
            // ---- TeraSort: Sort of 100-byte Records by their 10-byte key (examples/terasort/terasort.cpp:186-200) ----
            {
                const size_t nr = n / 4;
                auto recs = api::Generate(ctx, nr, [](size_t i) {
                                              Record r;
                                              uint64_t a = splitmix64(2 * i + 42), b = splitmix64(2 * i + 43);
                                              memcpy(r.key, &a, 8);
                                              memcpy(r.key + 8, &b, 2);
                                              for (size_t j = 0; j < 90; ++j) r.value[j] = static_cast<uint8_t>((i * 131 + j * 7) & 0xff);
                                              return r;
                                          }).Cache().Keep(2);
                common::StatsTimerStart t_cpu;
                std::vector<Record> cpu = recs.Sort().AllGather();
                t_cpu.Stop();
                common::StatsTimerStart t_gpu;
                std::vector<Record> gpu = thrill_gpu::Sort(recs).AllGather();
                t_gpu.Stop();
                bool ok = cpu == gpu && cpu.size() == nr;
                if (ctx.my_rank() == 0)
                    printf("%s TeraSort Records n=%zu cpu=%.3fs gpu=%.3fs\n", ok ? "PASS" : "FAIL", nr, t_cpu.SecondsDouble(), t_gpu.SecondsDouble());
                if (!ok) g_failures++;
            }
            // ---- PageRank (cfg5's pipeline, examples/page_rank/page_rank.hpp:70-139), 10 iterations: ReduceToIndex on the GPU
            //      against the stock operator; tolerance of the reference's own test (tests/examples/page_rank_test.cpp:94)
            {
                const size_t pages = 50000;
                std::vector<double> cpu = PageRankRun<false>(ctx, pages, 10);
                std::vector<double> gpu = PageRankRun<true>(ctx, pages, 10);
                bool ok = cpu.size() == pages && gpu.size() == pages;
                double maxdiff = 0, sum = 0;
                for (size_t i = 0; ok && i < pages; ++i) {
                    maxdiff = std::max(maxdiff, std::abs(cpu[i] - gpu[i]));
                    sum += gpu[i];
                }
                ok = ok && maxdiff < 1e-6;
                if (ctx.my_rank() == 0)
                    printf("%s PageRank pages=%zu iterations=10 max|cpu-gpu|=%.3g sum=%.6f\n", ok ? "PASS" : "FAIL", pages, maxdiff, sum);
                if (!ok) g_failures++;
            }
        });
    if (rc != 0) return rc;
    return g_failures.load() ? 1 : 0;
}
