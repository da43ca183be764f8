"""Pre-products W = S (V S) at n = 10,000 (config 3) with the single-CTA and the CTA-pair kernel: same bits, kernel time.
Usage (GPU box): python scripts/prep_pair_probe.py [n]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eagleeverything_b200 import _lib, device  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
lib = device.init(0)
g = torch.Generator(device="cuda"); g.manual_seed(1234)
S = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); S = (S + S.T) * (0.5 / n ** 0.5); S.diagonal().add_(2.0)
V = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); V = (V + V.T) * (0.5 / n ** 0.5); V.diagonal().add_(1.5)
a = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
out = {}
for pair in ("0", "1", "0", "1"):
    os.environ["EAGLE_PREP_PAIR"] = pair
    ms = []
    for rep in range(3):
        Wp = device.scan_prepare(S, V, a, n)
        torch.cuda.synchronize()
        k_ms, k_ops = C.c_double(), C.c_double()
        _lib.check(lib.eg_last_prep_kernels(C.byref(k_ms), C.byref(k_ops)))
        ms.append(round(k_ms.value, 2))
    out.setdefault(pair, Wp.clone())
    print(f"pair={pair} kernel ms {ms}  int8 TOP/s {k_ops.value / (ms[-1] * 1e-3) / 1e12:.0f}", flush=True)
Kpad = (n + 31) // 32 * 32
A, B = (out[k][: Kpad * n].view(n, Kpad).T[:n, :] for k in ("0", "1"))
print("upper triangles bit-identical:", bool(torch.equal(torch.triu(A), torch.triu(B))))
