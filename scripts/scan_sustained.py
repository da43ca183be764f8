"""Sustained (power-capped) behaviour of the int8 scan kernels next to a cuBLASLt int8 GEMM:
back-to-back launches for ~2 s each, nvidia-smi clocks / power sampled every 100 ms."""
import sys, os, subprocess, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eagleeverything_b200 import device, synth
n = int(os.environ.get("SW_N", 10000)); L = int(os.environ.get("SW_L", 250000)); REP = int(os.environ.get("SW_REP", 24))
lib = device.init(0)
img = device.synth_ascii(L, n, synth.GENO_SEED)
tt, _ = device.decode(img, n + 1, L, n)
del img
g = torch.Generator(device="cuda"); g.manual_seed(1)
S = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); S = (S + S.T) * (0.5 / n ** 0.5); S.diagonal().add_(2.0)
V = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); V = (V + V.T) * (0.5 / n ** 0.5); V.diagonal().add_(1.5)
a = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
Wp = device.scan_prepare(S, V, a, n)
del S, V

class Smi:
    def __enter__(self):
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                                   "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
        time.sleep(0.3)
        return self
    def __exit__(self, *a):
        self.p.terminate(); out, _ = self.p.communicate(timeout=5)
        rows = [[x.strip() for x in l.split(",")] for l in out.strip().splitlines() if l.count(",") >= 2]
        rows = rows[len(rows) // 3:]   # the loaded part
        clk = sorted(float(r[0]) for r in rows); pw = sorted(float(r[1]) for r in rows)
        self.s = f"sm {clk[len(clk)//2]:.0f} MHz, {pw[len(pw)//2]:.0f} W, power_cap {sum(r[2].startswith('Active') for r in rows)}/{len(rows)}"

def timed(fn, rep):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    with Smi() as smi:
        e0.record()
        for _ in range(rep): fn()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rep, smi.s

for name, env in [("pair", {"EAGLE_SI_PAIR": "1"}), ("single", {"EAGLE_SI_PAIR": "0"}),
                  ("pair,noflow", {"EAGLE_SI_PAIR": "1", "EAGLE_SCAN_FLOWCTL": "0"})]:
    for k in ("EAGLE_SI_PAIR", "EAGLE_SCAN_FLOWCTL"): os.environ.pop(k, None)
    os.environ.update(env)
    ms, s = timed(lambda: device.scan(tt, L, n, Wp), REP)
    kms, ops = C.c_double(), C.c_double(); lib.eg_last_scan_kernel(C.byref(kms), C.byref(ops))
    print(f"scan {name:12s}: {ms:7.2f} ms/scan (kernel {kms.value:6.2f} ms, {ops.value / kms.value / 1e9:6.0f} TOP/s) | {s}", flush=True)
N = 16384
ai = torch.randint(-1, 2, (N, N), dtype=torch.int8, device="cuda"); bi = torch.randint(-1, 2, (N, N), dtype=torch.int8, device="cuda")
ms, s = timed(lambda: torch._int_mm(ai, bi), 600)
print(f"cuBLASLt int8 {N}^3 : {ms:7.3f} ms  {2.0 * N ** 3 / ms / 1e9:6.0f} TOP/s | {s}", flush=True)
bi2 = torch.randint(-128, 128, (N, N), dtype=torch.int8, device="cuda")
ms, s = timed(lambda: torch._int_mm(ai, bi2), 600)
print(f"cuBLASLt int8 {N}^3 (B full-range bytes): {ms:7.3f} ms  {2.0 * N ** 3 / ms / 1e9:6.0f} TOP/s | {s}", flush=True)
