"""The reference's `ngpu` made live behind the C ABI (eg_init_multi: one process, one host thread per GPU, NCCL): the
five hot-path exports and the store API give, bit for bit, the single-GPU results on 2 (4, 8) GPUs.  Self-skipping on a
box with one GPU.  Each case runs in a fresh interpreter (the multi-GPU set is process-wide state)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()





@pytest.mark.parametrize("ngpu,n,L", [(2, 300, 5000), (2, 1100, 9001), (4, 515, 3000), (8, 300, 700)])
def test_exports_on_a_multi_gpu_set_equal_one_gpu(ngpu, n, L, tmp_path):
    if _ngpu() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    script = os.path.join(ROOT, "tests", "multi_abi_worker.py")
    r = subprocess.run([sys.executable, script, str(ngpu), str(n), str(L)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "MULTI_ABI_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
