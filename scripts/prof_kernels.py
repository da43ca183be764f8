"""Launches each hot kernel twice on profile-sized inputs (for `ncu -k regex:... -c 6`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from eagleeverything_b200 import device, synth

n = int(os.environ.get("PROF_N", 2000))
L = int(os.environ.get("PROF_L", 500000))
Ls = int(os.environ.get("PROF_LSCAN", 148 * 128))
device.init(0)
img = device.synth_ascii(n, L, synth.GENO_SEED)
st = None
def timed(name, fn, reps=2):
    out = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); out = fn(); e1.record(); e1.synchronize()
    print(f"{name}: {e0.elapsed_time(e1):.3f} ms", flush=True)
    return out
st, err = timed("decode", lambda: device.decode(img, L + 1, n, L))
C32 = torch.zeros((n, n), dtype=torch.int32, device="cuda")
timed("syrk", lambda: device.syrk(st, n, L, C32=C32, zero=True))
tt = device.transpose(st[:, :], n, Ls) if Ls < L else device.transpose(st, n, L)
g = torch.Generator(device="cuda"); g.manual_seed(1)
S = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); S = (S + S.T) * (0.5 / n ** 0.5); S.diagonal().add_(2.0)
V = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); V = (V + V.T) * (0.5 / n ** 0.5); V.diagonal().add_(1.5)
a = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
Wp = device.scan_prepare(S, V, a, n)
timed("scan", lambda: device.scan(tt, min(Ls, L), n, Wp))
torch.cuda.synchronize()
print("ok")
