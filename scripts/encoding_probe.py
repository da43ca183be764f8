"""How much of the scan's / SYRK's time is operand energy?  The same kernels on genotype images whose class frequencies
are permuted: (a) the synthetic data's 55 / 35 / 10 % as stored today (AA -> +1, AB -> 0, BB -> -1 = 0xFF),
(b) the frequent class stored as 0x00, (c) bytes 0x00 / 0x01 only (what raw codes would give for most markers)."""
import ctypes as C
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from eagleeverything_b200 import _lib, device

lib = device.init(0)
n, L = 10000, int(os.environ.get("PROBE_L", "200000"))
g = torch.Generator(device="cuda"); g.manual_seed(1)
S = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); S = (S + S.T) * (0.5 / n ** 0.5); S.diagonal().add_(2.0)
V = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); V = (V + V.T) * (0.5 / n ** 0.5); V.diagonal().add_(1.5)
a = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
Wp = device.scan_prepare(S, V, a, n)
del S, V


def image(freq):          # freq: probabilities of the characters '0', '1', '2'
    u = torch.rand((n, L), device="cuda", generator=g)
    img = torch.full((n, L + 1), ord("\n"), dtype=torch.uint8, device="cuda")
    codes = (u > freq[0]).to(torch.uint8) + (u > freq[0] + freq[1]).to(torch.uint8)
    img[:, :L] = codes + ord("0")
    return torch.cat([img.reshape(-1), torch.zeros(64, dtype=torch.uint8, device="cuda")])


for name, freq in (("a: '0' 55% '1' 35% '2' 10%  (+1 / 0 / 0xFF, as stored today)", (0.55, 0.35, 0.10)),
                   ("b: '1' 55% '0' 35% '2' 10%  (frequent class = 0x00)", (0.35, 0.55, 0.10)),
                   ("c: '1' 55% '0' 45%          (bytes 0x00 / 0x01 only)", (0.45, 0.55, 0.0)),
                   ("d: '2' 55% '1' 35% '0' 10%  (frequent class = 0xFF)", (0.10, 0.35, 0.55))):
    img = image(freq)
    kb, err = device.decode_kb(img, L + 1, n, L)
    tT = device.transpose_kb(kb, n, L)
    C32 = torch.empty((n, n), dtype=torch.int32, device="cuda")
    del img
    for what in ("scan", "syrk"):
        f = (lambda: device.scan(tT, L, n, Wp)) if what == "scan" else (lambda: device.syrk_kb(kb, n, L, C32=C32))
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        reps = 60 if what == "scan" else 300
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f()
        e1.record(); e1.synchronize()
        print(f"{name:62s} {what}: {e0.elapsed_time(e1) / reps:8.3f} ms per call over {reps} back-to-back calls", flush=True)
    del kb, tT, C32
    torch.cuda.empty_cache()
    time.sleep(2)
