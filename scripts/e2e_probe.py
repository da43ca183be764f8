"""Per-call wall times of the host-level ABI sequence of bench.py's e2e leg, several steps in a row."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eagleeverything_b200 import _lib, device, synth
n = int(os.environ.get("SW_N", 10000)); L = int(os.environ.get("SW_L", 1000000))
lib = device.init(0)
img = device.synth_ascii(n, L, synth.GENO_SEED)
nb = n * (L + 1)
img_h = torch.empty(nb + 64, dtype=torch.uint8, pin_memory=True); img_h[:nb].copy_(img[:nb]); img_h[nb:].zero_()
del img; torch.cuda.empty_cache()
S, V, a = synth.scan_inputs(n)
S_h = torch.from_numpy(S).pin_memory(); V_h = torch.from_numpy(V).pin_memory(); a_h = torch.from_numpy(a).pin_memory()
K_h = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
oa = torch.empty(L, dtype=torch.float64, pin_memory=True); ov = torch.empty(L, dtype=torch.float64, pin_memory=True)
vp = C.c_void_p; dp = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
for step in range(int(os.environ.get("STEPS", 10))):
    t = [time.perf_counter()]
    h, ht = vp(), vp()
    _lib.check(lib.eg_store_from_host_ascii(vp(img_h.data_ptr()), n, L, 0, L, C.byref(h))); t.append(time.perf_counter())
    _lib.check(lib.eg_store_mmt(h, None, 0, dp(K_h))); t.append(time.perf_counter())
    _lib.check(lib.eg_store_transpose(h, C.byref(ht))); t.append(time.perf_counter())
    _lib.check(lib.eg_store_a_and_vara(ht, None, 0, dp(S_h), dp(V_h), dp(a_h), dp(oa), dp(ov))); t.append(time.perf_counter())
    lib.eg_store_free(h); lib.eg_store_free(ht); torch.cuda.synchronize(); t.append(time.perf_counter())
    names = ["from_host_ascii", "mmt", "transpose", "a_and_vara", "free"]
    print(f"step {step}: total {1e3 * (t[-1] - t[0]):8.1f} ms  " + "  ".join(f"{k} {1e3 * (t[i + 1] - t[i]):7.1f}" for i, k in enumerate(names)), flush=True)
