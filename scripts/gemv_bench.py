"""a = Mt * v at a given shape: FP64 kernel vs exact DP4A kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eagleeverything_b200 import device, synth
n = int(os.environ.get("SW_N", 10000)); L = int(os.environ.get("SW_L", 1000000))
device.init(0)
img = device.synth_ascii(L, n, synth.GENO_SEED)
tt, _ = device.decode(img, n + 1, L, n)
del img
x = torch.randn(n, dtype=torch.float64, device="cuda")
def timed(fn, rep=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(rep): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rep
res = {}
for mode in ("i8", "f64"):
    os.environ["EAGLE_GEMV_MODE"] = mode
    ms = timed(lambda: device.gemv_i8(tt, L, n, x, 1.5))
    res[mode] = device.gemv_i8(tt, L, n, x, 1.5)
    print(f"gemv {mode}: {ms:.3f} ms  ({L * tt.stride(0) / ms / 1e6:.0f} GB/s of store)", flush=True)
ref = -1.5 * (tt[:4096, :n].double() @ x)   # the store holds the negated genotype values
for mode in res:
    print(mode, "max rel err vs torch on 4096 rows:", ((res[mode][:4096] - ref).abs() / ref.abs().clamp_min(1e-6)).max().item())
