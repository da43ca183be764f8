"""Multi-rank parity on real NCCL (skipped on boxes with fewer than 2 GPUs): under torchrun with 2 (and, when present,
4 / 8) ranks the all-reduced K, the sharded pre-products, every shard of a / var(a), the pick and the whole forward
search are bit-identical to a single-GPU evaluation of the same data set (tests/multirank_worker.py asserts it)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_equals_single_gpu(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29530 + world), os.path.join(ROOT, "tests", "multirank_worker.py"), "3300", "40000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTIRANK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
