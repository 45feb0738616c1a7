"""Multi-GPU parity (needs >= 2 CUDA devices): sample-sharded Hessian + NCCL all-reduce, distributed
inverses, row-sharded sweep with in-library NCCL for the SSR statistics, against the single-GPU path."""

import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_layer_matches_single_gpu(exchange):
    """exchange: how the row-sharded SSR sweep all-reduces its per-block statistics -- 'p2p' = the one-shot all-reduce over
    peer-memory mailboxes fused into the fold / gather kernels (csrc/comm.cu), 'nccl' = the in-library NCCL communicator."""
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29611" if exchange == "p2p" else "29612",
           os.path.join(HERE, "mgpu_worker.py")]
    env = dict(os.environ, TQ_P2P_ALLREDUCE="1" if exchange == "p2p" else "0", TQ_EXPECT_P2P="1" if exchange == "p2p" else "0")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "MGPU_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_sharded_whole_model_matches_reference():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29613", os.path.join(HERE, "mgpu_model_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "MGPU_MODEL_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
