"""bench.py's command-line contract, as far as it can be checked without a GPU: the reference arm (the unmodified reference on
the host cores) prints one JSON line with the keys the driver reads, on the same `config` object our arm would print; our arm
refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.ref
@pytest.mark.parametrize("metric,nflag,unit", [("sort", "--n", "keys/s"), ("reduce", "--reduce-n", "records/s")])
def test_reference_arm_prints_the_contract_line(metric, nflag, unit):
    if not O.have_ref_driver():
        pytest.skip("oracle/_ref/thrill_ref_driver not built")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--metric", metric, "--gpus", "2",
           "--steps", "1", "--warmup", "1", nflag, "200000"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, RANK="0"))
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "config", "cpu_baseline", "e2e", "dtype", "data"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == unit and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    # the same config object as our arm prints for this metric / size / worker count (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.workload_config(metric, 200000, 2)
    # the other ranks of a torchrun launch exit without work
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=60, env=dict(os.environ, RANK="1"))
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--n", "1000"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)
