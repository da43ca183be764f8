"""The scan pre-products W = S (V S) at n = 10,000 (BASELINE config 3): int8-slice path and cuBLAS path, timed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eagleeverything_b200 import device
n = int(os.environ.get("SW_N", 10000))
device.init(0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
S = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); S = (S + S.T) * (0.5 / n ** 0.5); S.diagonal().add_(2.0)
V = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); V = (V + V.T) * (0.5 / n ** 0.5); V.diagonal().add_(1.5)
a = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
tmp = torch.empty(n * n, dtype=torch.float64, device="cuda")
res = {}
for mode in os.environ.get("SW_MODES", "i8,f64").split(","):
    os.environ["EAGLE_PREP_MODE"] = mode
    Wp = device.scan_prepare(S, V, a, n, tmp=tmp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(int(os.environ.get("SW_REP", 3))):
        Wp = device.scan_prepare(S, V, a, n, tmp=tmp)
    e1.record(); torch.cuda.synchronize()
    res[mode] = Wp.clone()
    print(f"prepare {mode}: {e0.elapsed_time(e1) / int(os.environ.get('SW_REP', 3)):.2f} ms", flush=True)
if len(res) == 2:
    d = (res["i8"] - res["f64"]).abs().max().item(); m = res["f64"].abs().max().item()
    print(f"max |U_i8 - U_f64| = {d:.3e} (max |U| = {m:.3e})")
