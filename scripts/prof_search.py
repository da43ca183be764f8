"""A short forward search (n = 10,000 individuals, few markers) for profiling the kernels of the eigenbasis route:
sec_roots / sec_lowner / sec_transform (secular solve), pi_slice + prep_i8_kernel (the one n^3 product), eig_fold."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from eagleeverything_b200 import am, device, synth

device.init(0)
n, L = int(os.environ.get("PROF_N", "10000")), int(os.environ.get("PROF_L", "40960"))
img = device.synth_ascii(n, L, 20261018)
kb, _ = device.decode_kb(img, L + 1, n, L)
tT = device.transpose_kb(kb, n, L)
rng = np.random.default_rng(7)
y = 10.0 + rng.standard_normal(n)
for b, j in zip([1.0, 0.8, 0.6], [L // 5, L // 2, 4 * L // 5]):
    y = y + b * device.extract_col(kb, n, j, kblocked=True).cpu().numpy().astype(np.float64)
t0 = time.perf_counter()
r = am.AM_resident(kb, tT, n, L, y, maxit=int(os.environ.get("PROF_MAXIT", "3")))
print("picked", r["all_picked"], "seconds", r["seconds"], "secular", r["secular"], "wall", round(time.perf_counter() - t0, 3))
