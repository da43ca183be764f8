"""One launch of each hot kernel at the BASELINE config-3 shape (n=10,000 x L=1,000,000) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eagleeverything_b200 import device, synth
n, L = 10000, 1000000
device.init(0)
img = device.synth_ascii(n, L, synth.GENO_SEED)
stk, err = device.decode_kb(img, L + 1, n, L)
C32 = device.syrk_kb(stk, n, L)
tt = device.transpose_kb(stk, n, L)
del stk, img
g = torch.Generator(device="cuda"); g.manual_seed(1)
S = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); S = (S + S.T) * (0.5 / n ** 0.5); S.diagonal().add_(2.0)
V = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); V = (V + V.T) * (0.5 / n ** 0.5); V.diagonal().add_(1.5)
a = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
Wp = device.scan_prepare(S, V, a, n)
oa, ov = device.scan(tt, L, n, Wp)
torch.cuda.synchronize()
print("ok", float(ov[0]))
