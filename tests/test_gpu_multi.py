"""Multi-GPU numerical test (needs >= 2 B200s: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`; skipped on
a 1-GPU box).  The comparison itself lives in tests/dist_numeric_check.py (run under torchrun with NCCL)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_gradients_and_update_equal_single_gpu_on_the_concatenated_batch():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dist_numeric_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "dist_numeric_check ok" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("env", [{}, {"KMBART_PEER_TAIL": "kernel"}, {"KMBART_PEER_SIGNAL": "kernel"}],
                         ids=["copy_engines", "one_kernel_tail", "kernel_flags"])
def test_peer_memory_gradient_exchange_equals_rank_ordered_mean(env):
    """csrc/peer_exchange.cu against an all-gather + rank-ordered sum on ragged regions: bit-identical on every rank
    (tests/dist_peer_exchange_check.py), for both transports and both ways of moving the flags."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "tests", "dist_peer_exchange_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, **env})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "PEER_EXCHANGE_OK" in r.stdout
