"""Column-sharded pre-products (what each rank runs at N GPUs) timed per shard on one GPU."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eagleeverything_b200 import device, _lib
n = int(os.environ.get("SW_N", 10000)); world = int(os.environ.get("SW_WORLD", 2))
lib = device.init(0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
S = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); S = (S + S.T) * (0.5 / n ** 0.5); S.diagonal().add_(2.0)
V = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); V = (V + V.T) * (0.5 / n ** 0.5); V.diagonal().add_(1.5)
tmp = torch.empty(n * n, dtype=torch.float64, device="cuda")
Wp = torch.zeros(lib.eg_scan_wp_elems(n), dtype=torch.float64, device="cuda")
cuts = [min(n, int(round(n * ((1.0 + 3.0 * r / world) ** 0.5 - 1.0) / 32.0)) * 32) for r in range(world)] + [n]
vp = lambda t: C.c_void_p(t.data_ptr())
for rep in range(2):
    for r in range(world):
        c0, c1 = cuts[r], cuts[r + 1]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        _lib.check(lib.eg_dev_scan_prepare_cols(vp(S), vp(V), n, c0, c1, 1, vp(tmp), vp(Wp), None))
        e1.record(); torch.cuda.synchronize()
        ms, ops = C.c_double(), C.c_double(); lib.eg_last_prep_kernels(C.byref(ms), C.byref(ops))
        print(f"rep {rep} shard {r} cols [{c0},{c1}): {e0.elapsed_time(e1):.2f} ms (prep_i8 kernels {ms.value:.2f} ms, {ops.value / ms.value / 1e9:.0f} TOP/s)", flush=True)
