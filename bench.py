#!/usr/bin/env python
"""bench.py — Sort keys/s (headline) and ReduceByKey records/s of the B200-native Thrill hot path.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one synthetic batch (SURVEY.md §8d generators, seed 42):
  sort   : DIA<uint64_t>::Sort of 1e8 uniform keys per GPU          (BASELINE.json configs[1]; weak scaling)
  reduce : ReducePair<uint64_t,double>(plus) of 1.25e8 Zipf(s=1, U=2^26) records per GPU (configs[2] per-GPU share)
`value` is the device-resident whole-job Sort throughput (inputs already in HBM); `e2e` is the same operator
through the reference-facing call (host Blocks in pinned memory -> tg_sort_file -> tg_fetch_output, H2D and D2H
inside the timed region).  `roofline` is measured live with CUDA events around every launch of the dominant
kernel; `cpu_baseline` is the UNMODIFIED reference (oracle/_ref/thrill_ref_driver) timed on this box's host
cores.  The oracle / reference are only ever the checker or the baseline here, never the measured product.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 42
SORT_N_PER_GPU = 100000000
REDUCE_N_PER_GPU = 125000000
ZIPF_UNIVERSE = 1 << 26
SORT_PASS_BYTES_PER_KEY = 16.0          # one onesweep pass: read 8 + write 8
REDUCE_PASS_BYTES_PER_RECORD = 32.0     # one hash-digit partition pass: read 16 + write 16


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json (measured)"
    return 6650.0, "B200_PROFILING.md fallback"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        # median over the samples taken under load (the upper half of the observed clocks)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def ref_driver_path():
    return os.path.join(ROOT, "oracle", "_ref", "thrill_ref_driver")


_REF_WORKERS = {}


def reference_workers(op, extra=()):
    """The reference's throughput is not monotone in its worker count (p^2 streams, one thread per worker):
    calibrate on a small sample and keep the best of a few counts up to all host threads."""
    if op in _REF_WORKERS:
        return _REF_WORKERS[op]
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (ncpu, ncpu // 2, ncpu // 4, 32, 16, 8) if 1 <= c <= ncpu}, reverse=True)
    best, best_t = cands[0], None
    for c in cands:
        try:
            t = _run_reference_once(op, 20000000 if op == "sort_u64" else 5000000, 2, c, extra)[-1]
        except Exception:
            continue
        if best_t is None or t < best_t:
            best, best_t = c, t
    _REF_WORKERS[op] = best
    return best


def _run_reference_once(op, n, iters, workers, extra=()):
    env = dict(os.environ, THRILL_NET="mock", THRILL_LOCAL="1", THRILL_WORKERS_PER_HOST=str(workers), THRILL_LOG="")
    args = [ref_driver_path(), "op=%s" % op, "n=%d" % n, "iters=%d" % iters, "seed=%d" % SEED] + list(extra)
    res = subprocess.run(args, env=env, capture_output=True, text=True, timeout=3000)
    if res.returncode != 0:
        raise RuntimeError("thrill_ref_driver failed (%d): %s" % (res.returncode, res.stderr[-1500:]))
    return [float(l.rsplit("time=", 1)[1]) for l in res.stdout.splitlines() if l.startswith("RESULT")]


def run_reference(op, n, iters, extra=()):
    """time the unmodified reference on the host cores; returns (per-iteration seconds, workers used)"""
    workers = reference_workers(op)
    return _run_reference_once(op, n, iters, workers, extra), workers


def run_oracle_port_sort(n):
    """fallback CPU baseline when the reference binary is absent: the single-threaded C restatement"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    keys = O.gen_sort_uniform(0, n, SEED)
    t0 = time.time()
    O.lib().to_sort_items(keys.ctypes.data, n, C.byref(O.U64_DESC))
    return time.time() - t0


def cpu_baseline_sort(n_sample):
    if os.path.exists(ref_driver_path()):
        times, workers = run_reference("sort_u64", n_sample, 3)
        t = statistics.median(times[1:]) if len(times) > 1 else times[0]
        return {"value": n_sample / t, "unit": "keys/s", "cores": workers, "kind": "reference",
                "sample": "thrill_ref_driver Generate(splitmix64).Cache -> Sort().Size() of %d u64 keys, "
                          "THRILL_WORKERS_PER_HOST=%d (best of a calibration over worker counts up to %d host threads), median of iterations 2-3 (%.3f s)" % (n_sample, workers, os.cpu_count() or 1, t)}
    t = run_oracle_port_sort(n_sample // 10)
    return {"value": (n_sample // 10) / t, "unit": "keys/s", "cores": 1, "kind": "port",
            "sample": "oracle/thrill_oracle.c to_sort_items on %d keys, 1 thread (%.3f s)" % (n_sample // 10, t)}


def cpu_baseline_reduce(n_sample):
    """the unmodified reference's ReducePair<uint64_t,double>(plus) on the same Zipf(1.0, 2^26) generator, bounded sample"""
    if not os.path.exists(ref_driver_path()):
        return None
    extra = ("gen=zipf", "universe=%d" % ZIPF_UNIVERSE)
    workers = reference_workers("reduce_f64", extra)
    times = _run_reference_once("reduce_f64", n_sample, 3, workers, extra)
    t = statistics.median(times[1:]) if len(times) > 1 else times[0]
    return {"value": n_sample / t, "unit": "records/s", "cores": workers, "kind": "reference",
            "sample": "thrill_ref_driver Generate(Zipf s=1 U=2^26).Cache -> ReducePair(plus<double>).Size() of %d records, "
                      "THRILL_WORKERS_PER_HOST=%d, median of iterations 2-3 (%.3f s)" % (n_sample, workers, t)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = args.n
    if not os.path.exists(ref_driver_path()):
        t = run_oracle_port_sort(n // 10)
        value, cores, kind, sample = (n // 10) / t, 1, "port", "oracle C port on %d keys" % (n // 10)
        ms = t * 1e3
    else:
        times, cores = run_reference("sort_u64", n, args.warmup + args.steps)
        timed = times[args.warmup:]
        ms = 1e3 * sum(timed) / len(timed)
        value, kind = n / (ms / 1e3), "reference"
        sample = ("unmodified thrill/thrill Sort() of %d uniform u64 keys per step on %d host threads (mock net, "
                  "1 host); at --gpus>1 the GPU arm sorts %d keys per GPU, the CPU arm keeps this bounded sample"
                  % (n, cores, n))
    line = {"impl": "reference", "metric": "sort_keys_per_s", "value": value, "unit": "keys/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "sort_uniform_u64_1e8_per_gpu", "keys_per_step": n, "generator": "splitmix64(i+42)"},
            "cpu_baseline": {"value": value, "unit": "keys/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "keys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=SORT_N_PER_GPU, help="sort keys per GPU")
    ap.add_argument("--reduce-n", type=int, default=REDUCE_N_PER_GPU, help="reduce records per GPU (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)

    from thrill_b200 import api, capi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (args.gpus, args.gpus))
    ctx = api.Context.from_env(rng_seed=SEED)
    tg = ctx.tg
    L = tg.L
    n = args.n
    K, W = args.steps, max(args.warmup, 3)
    hbm_peak, peak_src = measured_peaks()
    sampler = ClockSampler(tg.device)

    # ------------------------------------------------------------------ Sort, device resident ----------
    desc = capi.u64_desc()
    d_in = tg.alloc(n * 8)
    sampler.start()
    step_ms = []
    launches0 = None
    for it in range(W + K):
        tg.ck(L.tg_gen_sort_uniform(tg.h, d_in, rank * n, n, SEED))
        if it == W:
            tg.profile_enable(True)
            launches0 = tg.launches()
        tg.barrier()
        tg.timer_start()
        out_p, out_n = C.c_void_p(), C.c_size_t()
        tg.ck(L.tg_sort(tg.h, C.byref(desc), d_in, n, SEED + it, C.byref(out_p), C.byref(out_n)))
        ms = tg.timer_stop()
        ms = max_over_ranks(ms, world)
        if it >= W:
            step_ms.append(ms)
    sort_launches = tg.launches() - launches0
    part_ms, part_cnt = tg.profile_get(capi.K_PARTITION)
    hist_ms, hist_cnt = tg.profile_get(capi.K_RADIX_HIST)
    merge_ms, merge_cnt = tg.profile_get(capi.K_MERGE)
    fix_ms, fix_cnt = tg.profile_get(capi.K_FIXUP)
    segc_ms, segc_cnt = tg.profile_get(capi.K_SEGCOUNT)
    xchg_ms, xchg_cnt = tg.profile_get(capi.K_EXCHANGE)
    tg.profile_enable(False)
    # cheap parity properties on the last result (outside the timed region)
    ok_sorted = tg.is_sorted(desc, out_p.value, out_n.value)
    total_out = sum_over_ranks(float(out_n.value), world)
    if not ok_sorted or int(total_out) != n * world:
        raise SystemExit("bench: sort result failed its property check (sorted=%s, items=%d)" % (ok_sorted, total_out))
    ms_per_step = sum(step_ms) / len(step_ms)
    value = n * world / (ms_per_step / 1e3)
    pass_launch_ms = part_ms / max(part_cnt, 1)
    achieved = SORT_PASS_BYTES_PER_KEY * n / (pass_launch_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": "tgp::partition_kernel<1,256,16,3,RadixDigit,SEG> (one stable 8-bit partition pass: "
                                          "read 8 B + write 8 B per key)",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "peak_source": peak_src,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on 1e8 keys, from the committed
                # `ncu --set full` capture profiles/r1e_segmented_pass.txt (834.7 MB + 782.6 MB); scaled to this run's n
                "traffic": (834.701824e6 + 782.626560e6) * (n / 1e8), "traffic_source": "profiles/r1e_segmented_pass.txt (ncu, n=1e8)",
                "algorithmic_bytes_per_launch": SORT_PASS_BYTES_PER_KEY * n, "launch_ms": pass_launch_ms,
                "launches_timed": part_cnt,
                "step_share": {"partition_ms": part_ms / K, "partition_launches_per_step": part_cnt / K,
                               "radix_hist_ms": hist_ms / K, "segment_count_ms": segc_ms / K, "finishing_pass_ms": fix_ms / K,
                               "merge_ms": merge_ms / K, "nccl_alltoallv_ms": xchg_ms / K, "step_ms": ms_per_step},
                "prefix_sort_fallbacks": int(L.tg_prefix_sort_fallbacks(tg.h))}

    # ------------------------------------------------------------------ Sort, end to end (host Files) ---
    host_in = tg.host_alloc(n * 8)
    host_out = tg.host_alloc(n * 8 + (n // 4) * 8)        # a worker may receive more than it sent (eps = 0.1)
    tg.ck(L.tg_gen_sort_uniform(tg.h, d_in, rank * n, n, SEED))
    tg.ck(L.tg_download(tg.h, host_in.ctypes.data, d_in, n * 8))
    tg.sync()
    dia = api.DIA(ctx, host_in.view(np.uint64))
    e2e_s = []
    for it in range(2 + K):
        tg.barrier()
        t0 = time.perf_counter()
        res = dia.Sort(_pinned_out=host_out)               # tg_sort_file + tg_fetch_output (synchronises)
        dt = max_over_ranks(time.perf_counter() - t0, world)
        if it >= 2:
            e2e_s.append(dt)
    e2e_value = n * world / (sum(e2e_s) / len(e2e_s))
    if not bool(np.all(res.items[1:] >= res.items[:-1])):
        raise SystemExit("bench: e2e sort result is not sorted")
    e2e = {"value": e2e_value, "unit": "keys/s", "h2d_bytes_per_step": n * 8 * world,
           "d2h_bytes_per_step": int(total_out) * 8, "ms_per_step": 1e3 * sum(e2e_s) / len(e2e_s),
           "call": "thrill_b200.api.DIA.Sort -> tg_sort_file + tg_fetch_output over 1 MiB pinned Blocks"}
    tg.free(d_in)
    tg.host_free(host_in); tg.host_free(host_out)

    # ------------------------------------------------------------------ ReduceByKey (extra) -------------
    extra = {}
    rn = args.reduce_n
    if rn:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        cdf = zipf_cdf_numpy(ZIPF_UNIVERSE)
        d_cdf = tg.to_device(cdf)
        d_rin = tg.alloc(rn * 16)
        kvd = capi.KVDesc(16, capi.OP_SUM_F64)
        r_ms = []
        for it in range(2 + 3):
            tg.ck(L.tg_gen_reduce_zipf(tg.h, d_rin, rank * rn, rn, SEED, d_cdf, ZIPF_UNIVERSE, 0))
            if it == 2:
                tg.profile_enable(True)
            tg.barrier()
            tg.timer_start()
            rp, rcount = C.c_void_p(), C.c_size_t()
            tg.ck(L.tg_reduce_by_key(tg.h, C.byref(kvd), d_rin, rn, C.byref(rp), C.byref(rcount)))
            ms = max_over_ranks(tg.timer_stop(), world)
            if it >= 2:
                r_ms.append(ms)
        agg_ms, agg_cnt = tg.profile_get(capi.K_AGGREGATE)
        cmp_ms, cmp_cnt = tg.profile_get(capi.K_COMPACT)
        rpart_ms, rpart_cnt = tg.profile_get(capi.K_PARTITION)
        rhist_ms, rhist_cnt = tg.profile_get(capi.K_RADIX_HIST)
        rsegc_ms, rsegc_cnt = tg.profile_get(capi.K_SEGCOUNT)
        rx_ms, rx_cnt = tg.profile_get(capi.K_EXCHANGE)
        tg.profile_enable(False)
        r_step = sum(r_ms) / len(r_ms)
        # dominant kernel: the stable partition pass over 16-byte records (the first two launches of a step move all
        # rn records: read 16 + write 16 bytes each); later launches of a multi-GPU step move fewer
        part_launch = rpart_ms / max(rpart_cnt, 1)
        r_ach = REDUCE_PASS_BYTES_PER_RECORD * rn / (part_launch / 1e3) / 1e9 if world == 1 else None
        extra = {"reduce_records_per_s": rn * world / (r_step / 1e3), "reduce_ms_per_step": r_step,
                 "reduce_distinct_out": int(sum_over_ranks(float(rcount.value), world)),
                 "reduce_config": {"workload": "reduce_pair_u64_f64_zipf_s1_U2^26", "records_per_gpu": rn},
                 "reduce_roofline": {"bound": "hbm", "kernel": "tgp::partition_kernel<2,256,8,3,HashLevelDigit,SEG> (one hash-digit pass)",
                                     "achieved": r_ach, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": (r_ach / hbm_peak) if r_ach else None, "launch_ms": part_launch,
                                     "step_share": {"partition_ms": rpart_ms / 3, "partition_launches_per_step": rpart_cnt / 3,
                                                    "count_ms": (rhist_ms + rsegc_ms) / 3, "aggregate_ms": agg_ms / 3,
                                                    "compact_ms": cmp_ms / 3, "nccl_alltoallv_ms": rx_ms / 3, "step_ms": r_step}}}
        # ReduceByKey-uniform: the same operator where there is little to reduce (distinct ~ 0.45 n).  Guarded:
        # an extra must never cost the headline line.
        try:
            u_ms = []
            for it in range(1 + 2):
                tg.ck(L.tg_gen_reduce_uniform(tg.h, d_rin, rank * rn, rn, SEED, ZIPF_UNIVERSE, 0))
                tg.barrier()
                tg.timer_start()
                rp2, rc2 = C.c_void_p(), C.c_size_t()
                tg.ck(L.tg_reduce_by_key(tg.h, C.byref(kvd), d_rin, rn, C.byref(rp2), C.byref(rc2)))
                ms = max_over_ranks(tg.timer_stop(), world)
                if it >= 1:
                    u_ms.append(ms)
            u_step = sum(u_ms) / len(u_ms)
            extra["reduce_uniform"] = {"records_per_s": rn * world / (u_step / 1e3), "ms_per_step": u_step,
                                       "distinct_out": int(sum_over_ranks(float(rc2.value), world)),
                                       "workload": "reduce_pair_u64_f64_uniform_U2^26"}
        except Exception as e:          # noqa: BLE001
            extra["reduce_uniform"] = {"error": str(e)[:200]}
        tg.free(d_rin); tg.free(d_cdf)

    clocks = sampler.stop()

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sort(n)
        if rn and extra:
            extra["reduce_cpu_baseline"] = cpu_baseline_reduce(min(rn, 25000000))

    if rank == 0:
        line = {"metric": "sort_keys_per_s", "value": value, "unit": "keys/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic",
                "config": {"workload": "sort_uniform_u64_1e8_per_gpu", "keys_per_gpu": n, "generator": "splitmix64(i+42)",
                           "l2": "inputs (0.8 GB per GPU) larger than the 126 MB L2; regenerated on the device before every step",
                           "exchange": "none (1 GPU)" if world == 1 else "NCCL Alltoallv (ncclSend/ncclRecv group)"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(sort_launches),
                "clocks": clocks, "extra": extra}
        print(json.dumps(line))
    ctx.close()
    return 0


def zipf_cdf_numpy(universe, s=1.0):
    """cumulative Zipf table, probabilities as common/zipf_distribution.hpp:119-140 (k^-s normalised).
    Sequential accumulation like the oracle (np.cumsum is sequential in float64)."""
    k = np.arange(1, universe + 1, dtype=np.float64)
    prob = 1.0 / np.power(k, s)
    p_sum = 0.0
    # the reference sums sequentially; math.fsum-free sequential sum to keep the same rounding
    p_sum = float(np.cumsum(prob)[-1])
    return np.cumsum(prob * (1.0 / p_sum))


if __name__ == "__main__":
    sys.exit(main())
