"""Multi-GPU (NCCL) checks; skipped with fewer than 2 GPUs."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_bench_two_ranks_nccl():
    """bench.py under torchrun, 2 ranks: shards + one NCCL all-gather; JSON contract fields present."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "1", "--batch", "8"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["config"]["global_batch"] == 16 and line["value"] > 0
    assert line["gpu_launches"] > 0 and line["scaling"] == "weak" and line["e2e"]["value"] > 0


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_train_mode_syncbn_two_ranks():
    """Sharded train-mode forward (global batch 4 over 2 ranks): BatchNorm statistics all-reduced over NCCL reproduce the
    unsharded reference's running statistics, guiding points and losses."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "tests", "multi_train_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "SYNCBN_OK" in out.stdout, out.stderr[-3000:]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_training_backward_two_ranks():
    """Data-parallel training step (global batch 4 over 2 ranks): SyncBN forward + backward sums over NCCL inside the library, one
    all-reduce of the flat gradient; the averaged gradients equal those of the unsharded batch on one GPU."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tests", "multi_backward_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "SHARDED_BACKWARD_OK" in out.stdout, out.stderr[-3000:]
