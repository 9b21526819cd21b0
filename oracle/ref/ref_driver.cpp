/*******************************************************************************
 * oracle/ref/ref_driver.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A small program of OURS that links the UNMODIFIED reference library (thrill/thrill, built by
 * oracle/ref/Makefile from /root/reference into oracle/_ref/libthrill_ref.a) and runs the reference's
 * own stock operators — DIA<T>::Sort (api/sort.hpp:800) and ReducePair / ReduceByKey
 * (api/reduce_by_key.hpp:312,410) — on the deterministic inputs of SURVEY.md §8(d) or on a raw binary
 * file.  It is used (a) to pin the CPU restatement in oracle/thrill_oracle.c, (b) to generate the golden
 * fixtures under tests/golden/ (tests/golden/make_golden.py), (c) as the "reference" CPU baseline that
 * bench.py times next to the GPU path.  Harness shape follows benchmarks/api/sort.cpp:43-58 and
 * benchmarks/hashtable/reduce.cpp:43-58 (Cache+Keep+Size first so generation is excluded, timer around
 * the operator + Size()).
 *
 * usage: THRILL_NET=mock THRILL_LOCAL=1 THRILL_WORKERS_PER_HOST=W \
 *        thrill_ref_driver op=<sort_u64|reduce_f64|reduce_u64|reduce_to_index|terasort|word_count> n=N [gen=uniform|zipf|file]
 *                          [in=path] [out=path] [iters=K] [seed=S] [universe=U] [exact=0|1]
 * Prints "RESULT op=... n=... workers=... hw_threads=... iter=i time=SECONDS" per iteration.
 ******************************************************************************/

#include <thrill/api/all_gather.hpp>
#include <thrill/api/cache.hpp>
#include <thrill/api/gather.hpp>
#include <thrill/api/generate.hpp>
#include <thrill/api/read_binary.hpp>
#include <thrill/api/read_lines.hpp>
#include <thrill/api/reduce_by_key.hpp>
#include <thrill/api/reduce_to_index.hpp>
#include <thrill/api/size.hpp>
#include <thrill/api/sort.hpp>
#include <thrill/common/stats_timer.hpp>

#include <examples/word_count/random_text_writer.hpp>
#include <examples/word_count/word_count.hpp>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <utility>
#include <vector>

using namespace thrill; // NOLINT

// --- deterministic generators; the same arithmetic lives in oracle/thrill_oracle.c ---
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t r) { return static_cast<double>(r >> 11) * (1.0 / 9007199254740992.0); }

// cumulative Zipf table, probabilities built as common/zipf_distribution.hpp:119-140 (k^-s, normalised)
static std::vector<double> g_zipf_cdf;
static void build_zipf(size_t U, double s) {
    g_zipf_cdf.resize(U);
    double p_sum = 0.0;
    for (size_t k = 1; k <= U; ++k) p_sum += 1.0 / std::pow(static_cast<double>(k), s);
    double p_norm = 1.0 / p_sum, acc = 0.0;
    for (size_t k = 1; k <= U; ++k) {
        acc += (1.0 / std::pow(static_cast<double>(k), s)) * p_norm;
        g_zipf_cdf[k - 1] = acc;
    }
}
static inline uint64_t zipf_rank(double u) {
    // smallest k with cdf[k-1] > u, clamped to U
    size_t idx = std::upper_bound(g_zipf_cdf.begin(), g_zipf_cdf.end(), u) - g_zipf_cdf.begin();
    if (idx >= g_zipf_cdf.size()) idx = g_zipf_cdf.size() - 1;
    return idx + 1;
}

struct Record {
    uint8_t key[10];
    uint8_t value[90];
    bool operator < (const Record& b) const {
        return std::lexicographical_compare(key, key + 10, b.key, b.key + 10);
    }
} __attribute__ ((packed));
static_assert(sizeof(Record) == 100, "Record packing");

static inline Record make_record(uint64_t i, uint64_t seed) {
    Record r;
    uint64_t a = splitmix64(2 * i + seed), b = splitmix64(2 * i + 1 + seed);
    std::memcpy(r.key, &a, 8);
    std::memcpy(r.key + 8, &b, 2);
    // payload: 90 bytes derived from the record index (11 x splitmix64 words, truncated)
    for (int w = 0; w < 12; ++w) {
        uint64_t v = splitmix64(i * 12 + w + (seed << 32));
        int len = (w == 11) ? 2 : 8;
        std::memcpy(r.value + 8 * w, &v, len);
    }
    return r;
}

struct ReduceOut {
    uint64_t key, valbits, worker;
} __attribute__ ((packed));

static std::map<std::string, std::string> g_args;
static std::string arg(const char* k, const char* def) {
    auto it = g_args.find(k);
    return it == g_args.end() ? std::string(def) : it->second;
}

template <typename T>
static void write_file(const std::string& path, const std::vector<T>& v) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { perror("fopen out"); exit(2); }
    if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
    fclose(f);
}

int main(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
        const char* eq = strchr(argv[i], '=');
        if (!eq) { fprintf(stderr, "bad arg %s\n", argv[i]); return 2; }
        g_args[std::string(argv[i], eq - argv[i])] = std::string(eq + 1);
    }
    const std::string op = arg("op", "sort_u64"), gen = arg("gen", "uniform");
    const std::string in = arg("in", ""), out = arg("out", "");
    const uint64_t n = strtoull(arg("n", "1000000").c_str(), nullptr, 10);
    const uint64_t seed = strtoull(arg("seed", "42").c_str(), nullptr, 10);
    const uint64_t universe = strtoull(arg("universe", "67108864").c_str(), nullptr, 10);
    const int iters = atoi(arg("iters", "1").c_str());
    const bool exact = atoi(arg("exact", "0").c_str()) != 0;
    if (gen == "zipf") build_zipf(universe, 1.0);

    auto key_of = [=](uint64_t i) -> uint64_t {
                      if (gen == "zipf") return zipf_rank(u01(splitmix64(i + seed)));
                      if (op == "sort_u64") return splitmix64(i + seed);
                      return 1 + splitmix64(i + seed) % universe;
                  };
    auto val_of = [=](uint64_t i) -> double {
                      uint64_t r = splitmix64(i + seed + (1ull << 40));
                      return exact ? static_cast<double>(r % 1024) : u01(r);
                  };

    return api::Run(
        [&](api::Context& ctx) {
            unsigned hw = std::thread::hardware_concurrency();
            auto report = [&](int it, double secs) {
                              if (ctx.my_rank() == 0) {
                                  printf("RESULT op=%s gen=%s n=%llu workers=%zu hw_threads=%u iter=%d time=%.6f\n",
                                         op.c_str(), gen.c_str(), (unsigned long long)n,
                                         ctx.num_workers(), hw, it, secs);
                                  fflush(stdout);
                              }
                          };
            if (op == "sort_u64") {
                auto input = (gen == "file")
                             ? api::ReadBinary<uint64_t>(ctx, in).Cache()
                             : api::Generate(ctx, n, [&](size_t i) { return key_of(i); }).Cache();
                input.Keep(iters + 1);
                input.Size();
                for (int it = 0; it < iters; ++it) {
                    ctx.net.Barrier();
                    common::StatsTimerStart timer;
                    auto sorted = input.Sort();
                    if (out.empty() || it + 1 < iters) {
                        sorted.Size();
                        ctx.net.Barrier();
                        timer.Stop();
                    }
                    else {
                        std::vector<uint64_t> all = sorted.Gather(0);
                        timer.Stop();
                        if (ctx.my_rank() == 0) write_file(out, all);
                    }
                    report(it, timer.SecondsDouble());
                }
            }
            else if (op == "reduce_f64" || op == "reduce_u64") {
                using PairF = std::pair<uint64_t, double>;
                using PairU = std::pair<uint64_t, uint64_t>;
                if (op == "reduce_f64") {
                    auto input = (gen == "file")
                                 ? api::ReadBinary<PairF>(ctx, in).Cache()
                                 : api::Generate(ctx, n, [&](size_t i) {
                                                     return PairF(key_of(i), val_of(i));
                                                 }).Cache();
                    input.Keep(iters + 1);
                    input.Size();
                    for (int it = 0; it < iters; ++it) {
                        ctx.net.Barrier();
                        common::StatsTimerStart timer;
                        auto red = input.ReducePair(std::plus<double>());
                        if (out.empty() || it + 1 < iters) {
                            red.Size();
                            ctx.net.Barrier();
                            timer.Stop();
                        }
                        else {
                            size_t me = ctx.my_rank();
                            std::vector<ReduceOut> all = red.Map([me](const PairF& p) {
                                                                     ReduceOut o;
                                                                     o.key = p.first;
                                                                     std::memcpy(&o.valbits, &p.second, 8);
                                                                     o.worker = me;
                                                                     return o;
                                                                 }).Gather(0);
                            timer.Stop();
                            if (ctx.my_rank() == 0) write_file(out, all);
                        }
                        report(it, timer.SecondsDouble());
                    }
                }
                else {
                    auto input = (gen == "file")
                                 ? api::ReadBinary<PairU>(ctx, in).Cache()
                                 : api::Generate(ctx, n, [&](size_t i) {
                                                     return PairU(key_of(i), splitmix64(i + seed + (1ull << 40)) % 1024);
                                                 }).Cache();
                    input.Keep(iters + 1);
                    input.Size();
                    for (int it = 0; it < iters; ++it) {
                        ctx.net.Barrier();
                        common::StatsTimerStart timer;
                        auto red = input.ReducePair(std::plus<uint64_t>());
                        if (out.empty() || it + 1 < iters) {
                            red.Size();
                            ctx.net.Barrier();
                            timer.Stop();
                        }
                        else {
                            size_t me = ctx.my_rank();
                            std::vector<ReduceOut> all = red.Map([me](const PairU& p) {
                                                                     ReduceOut o;
                                                                     o.key = p.first;
                                                                     o.valbits = p.second;
                                                                     o.worker = me;
                                                                     return o;
                                                                 }).Gather(0);
                            timer.Stop();
                            if (ctx.my_rank() == 0) write_file(out, all);
                        }
                        report(it, timer.SecondsDouble());
                    }
                }
            }
            else if (op == "reduce_to_index") {
                // DIA<pair<u64 index, double>>::ReduceToIndex(.first, sum of .second, size = universe) — the PageRank step
                // (examples/page_rank/page_rank.hpp:125-135); neutral element = PairF() = (0, 0.0)
                using PairF = std::pair<uint64_t, double>;
                const size_t result_size = static_cast<size_t>(std::stoull(arg("universe", "1024")));
                auto input = api::Generate(ctx, n, [&](size_t i) {
                                               return PairF(key_of(i) % result_size, val_of(i));
                                           }).Cache();
                input.Keep(iters + 1);
                input.Size();
                for (int it = 0; it < iters; ++it) {
                    ctx.net.Barrier();
                    common::StatsTimerStart timer;
                    auto red = input.ReduceToIndex(
                        [](const PairF& p) { return static_cast<size_t>(p.first); },
                        [](const PairF& a, const PairF& b) { return PairF(a.first, a.second + b.second); },
                        result_size);
                    if (out.empty() || it + 1 < iters) {
                        red.Size();
                        ctx.net.Barrier();
                        timer.Stop();
                    }
                    else {
                        std::vector<PairF> all = red.Gather(0);
                        timer.Stop();
                        if (ctx.my_rank() == 0) write_file(out, all);
                    }
                    report(it, timer.SecondsDouble());
                }
            }
            else if (op == "terasort") {
                auto input = (gen == "file")
                             ? api::ReadBinary<Record>(ctx, in).Cache()
                             : api::Generate(ctx, n, [&](size_t i) { return make_record(i, seed); }).Cache();
                input.Keep(iters + 1);
                input.Size();
                for (int it = 0; it < iters; ++it) {
                    ctx.net.Barrier();
                    common::StatsTimerStart timer;
                    auto sorted = input.Sort();
                    if (out.empty() || it + 1 < iters) {
                        sorted.Size();
                        ctx.net.Barrier();
                        timer.Stop();
                    }
                    else {
                        std::vector<Record> all = sorted.Gather(0);
                        timer.Stop();
                        if (ctx.my_rank() == 0) write_file(out, all);
                    }
                    report(it, timer.SecondsDouble());
                }
            }
            else if (op == "word_count") {
                // BASELINE.json configs[0] (plumbing): the reference's own examples/word_count/word_count.hpp on a text file (in=...)
                // or on n lines of 10 random words each (random_text_writer.hpp, as word_count_run -g does); out = "word count" lines
                using examples::word_count::WordCountPair;
                common::StatsTimerStart timer;
                std::vector<WordCountPair> all;
                if (gen == "file") {
                    all = examples::word_count::WordCount(api::ReadLines(ctx, in)).Gather(0);
                }
                else {
                    std::default_random_engine rng(static_cast<unsigned>(seed));
                    auto lines = api::Generate(ctx, n, [&](size_t) { return examples::word_count::RandomTextWriterGenerate(10, rng); });
                    all = examples::word_count::WordCount(lines).Gather(0);
                }
                timer.Stop();
                if (ctx.my_rank() == 0) {
                    std::sort(all.begin(), all.end());
                    size_t words = 0;
                    for (auto& wc : all) words += wc.second;
                    if (!out.empty()) {
                        FILE* f = fopen(out.c_str(), "w");
                        for (auto& wc : all) fprintf(f, "%s %zu\n", wc.first.c_str(), wc.second);
                        fclose(f);
                    }
                    printf("WORDCOUNT distinct=%zu words=%zu\n", all.size(), words);
                }
                report(0, timer.SecondsDouble());
            }
            else {
                if (ctx.my_rank() == 0) fprintf(stderr, "unknown op %s\n", op.c_str());
            }
        });
}
