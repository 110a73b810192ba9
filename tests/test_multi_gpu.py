"""Multi-GPU coverage of parallel.PeerAdam (the gradient all-reduce + sharded Adam + parameter
all-gather kernel over NVLink peer memory): runs tests/multi_gpu/peer_adam_check.py under torchrun
on two GPUs, unicast and NVLS forms.  Skipped on boxes with fewer than two GPUs (tools/gpu_multi.sh
runs the same check on the multi-GPU visits)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("nvls", ["0", "1"])
def test_peer_adam_matches_allreduce_plus_fused_adam(nvls):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, GAGS_B200_NVLS=nvls)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "multi_gpu", "peer_adam_check.py")]
    out = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("-> OK") == 2, out.stdout[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("nvls", ["0", "1"])
def test_sparse_peer_adam_matches_allreduce_plus_fused_adam(nvls):
    """parallel.SparsePeerAdam (row-sparse exchange + row-sparse Adam on every rank), both forms."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, GAGS_B200_NVLS=nvls)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29534",
           os.path.join(ROOT, "tests", "multi_gpu", "sparse_peer_adam_check.py")]
    out = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("-> OK") == 8, out.stdout[-2000:]
