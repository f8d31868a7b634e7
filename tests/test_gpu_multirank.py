"""N > 1 on real GPUs: the sharded engine (NCCL) must reproduce the single-GPU epochs.  Needs >= 2 GPUs
(skipped on the one-GPU box; run with `gpurun --gpus 2`).  The CPU counterpart is tests/test_distributed_host.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("backend", ["tensor", "simt"])
@pytest.mark.parametrize("shard_smoothing", [False, True], ids=["replicated_k3", "row_sharded_k3"])
def test_two_ranks_equal_one(shard_smoothing, backend):
    """Two NCCL ranks with half of the samples each == the float64 oracle (winners outside the near-tie gate, update to
    1e-5) == one GPU with all samples; `tensor` runs the tcgen05 search (selective FLAG + REFINE passes from the
    second epoch on) on every rank."""
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ)
    env["MR_BACKEND"] = backend
    env["DBGSOM_K3_SHARD_MIN_WORK"] = "0" if shard_smoothing else str(1 << 62)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731" if shard_smoothing else "29732", os.path.join(HERE, "_multirank_worker.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MULTIRANK_OK" in res.stdout
