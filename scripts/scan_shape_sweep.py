"""Times the int8 scan kernel at the c3 shape for several super-tile shapes (EAGLE_SI_MSUP x EAGLE_SI_GSUP)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eagleeverything_b200 import device, synth, _lib
n = int(os.environ.get("SW_N", 10000)); L = int(os.environ.get("SW_L", 250000))
lib = device.init(0)
img = device.synth_ascii(L, n, synth.GENO_SEED)          # Mt image
tt, _ = device.decode(img, n + 1, L, n)
del img
g = torch.Generator(device="cuda"); g.manual_seed(1)
S = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); S = (S + S.T) * (0.5 / n ** 0.5); S.diagonal().add_(2.0)
V = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); V = (V + V.T) * (0.5 / n ** 0.5); V.diagonal().add_(1.5)
a = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
Wp = device.scan_prepare(S, V, a, n)
ref = None
for shape in os.environ.get("SW_SHAPES", "16x9,37x4,74x2,148x1,8x18,24x6").split(","):
    m, gs = shape.split("x")
    os.environ["EAGLE_SI_MSUP"], os.environ["EAGLE_SI_GSUP"] = m, gs
    for _ in range(2):
        oa, ov = device.scan(tt, L, n, Wp)
    torch.cuda.synchronize()
    ms, ops = C.c_double(), C.c_double()
    lib.eg_last_scan_kernel(C.byref(ms), C.byref(ops))
    if ref is None: ref = ov.clone()
    print(f"shape {shape:>6}: {ms.value:8.2f} ms  {ops.value / ms.value / 1e9:7.1f} TOP/s  identical={bool(torch.equal(ov, ref))}", flush=True)
