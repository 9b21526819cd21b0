"""N>1 on real GPUs: launches tests/multi_gpu_worker.py with one process per GPU (torch.distributed.run) and
requires bit-exact parity with the reference outputs.  Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("pipeline", ["classify", "merge"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_collective_operators_on_n_gpus(world, pipeline):
    """pipeline = order of the multi-worker sort: classify/scatter -> exchange -> sort (default, the reference's order,
    api/sort.hpp:615-742) or sort -> boundaries -> exchange -> merge of the received runs (TG_SORT_PIPELINE=merge)"""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world + (10 if pipeline == "merge" else 0)),
           os.path.join(HERE, "multi_gpu_worker.py")]
    env = dict(os.environ, TG_SORT_PIPELINE=pipeline)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0 and "MULTI_GPU_PARITY_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-5000:]


def test_collective_operators_nccl_exchange():
    """TG_EXCHANGE=nccl: the two-step exchange (local partition, grouped ncclSend/ncclRecv) that is used where peers cannot map
    each other's windows"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(HERE, "multi_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, TG_EXCHANGE="nccl"))
    assert res.returncode == 0 and "MULTI_GPU_PARITY_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-5000:]
