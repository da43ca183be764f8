"""The n x n algebra of one AM() forward iteration on the device at n individuals:
K^(1/2), K^(-1/2) (first iteration only), H, P, a_hat, V -- device-level entry points, CUDA-event timing --
next to numpy/LAPACK on the host cores for the same functions (the oracle's restatements) when SW_CPU=1."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eagleeverything_b200 import device, _lib
n = int(os.environ.get("SW_N", 10000)); q = int(os.environ.get("SW_Q", 3))
lib = device.init(0)
g = torch.Generator(device="cuda"); g.manual_seed(5)
M = (torch.randint(0, 3, (n, 4 * n), device="cuda", generator=g, dtype=torch.int8).double() - 1.0)
K = M @ M.T; K = K / K.max(); K.diagonal().add_(0.95); K = ((K + K.T) * 0.5).contiguous()
del M
X = torch.cat([torch.ones(n, 1, dtype=torch.float64, device="cuda"), torch.randn(n, q - 1, dtype=torch.float64, device="cuda", generator=g)], 1).T.contiguous()  # column-major n x q
y = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
ve, vg = 0.7, 1.3
nn = n * n
new = lambda k: torch.empty(k, dtype=torch.float64, device="cuda")
sq, inv, tmp, H, P, V, D, small, a, t_n = new(nn), new(nn), new(nn), new(nn), new(nn), new(nn), new(nn), new(4 * n * q + 3 * q * q + 16), new(n), new(n)
vp = lambda t: C.c_void_p(t.data_ptr())
def timed(name, fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(f"{name:32s} {e0.elapsed_time(e1):9.1f} ms", flush=True)
    return e0.elapsed_time(e1)
not_pd, tr = C.c_int(0), C.c_double(0)
t = {}
t["sqrt_and_sqrtinv"] = timed("calculateMMt_sqrt_and_sqrtinv", lambda: _lib.check(lib.eg_dev_sqrt_and_sqrtinv(vp(K), n, vp(sq), vp(inv), vp(tmp), C.byref(not_pd), C.byref(tr), None)))
print("   positive definite:", not bool(not_pd.value), " trace(sqrt %*% invsqrt) =", tr.value)
t["H"] = timed("calculateH", lambda: _lib.check(lib.eg_dev_calculateH(vp(K), n, ve, vg, vp(H), None)))
t["P"] = timed("calculateP", lambda: _lib.check(lib.eg_dev_calculateP(vp(H), vp(X), n, q, vp(P), vp(small), None)))
t["a"] = timed("calculate_reduced_a", lambda: _lib.check(lib.eg_dev_calculate_reduced_a(vg, vp(P), vp(sq), vp(y), n, vp(t_n), vp(a), None)))
t["V"] = timed("calculate_reduced_vara", lambda: _lib.check(lib.eg_dev_calculate_reduced_vara(vp(X), q, ve, vg, vp(sq), n, vp(V), vp(D), vp(small), None)))
t["eigen"] = timed("eigen(K, symmetric=TRUE)", lambda: (tmp.copy_(K.view(-1)), _lib.check(lib.eg_dev_eigen_sym(vp(tmp), n, vp(t_n.new_empty(n)), None))))
print(f"per iteration (H + P + a + V): {t['H'] + t['P'] + t['a'] + t['V']:.1f} ms; first iteration adds {t['sqrt_and_sqrtinv']:.1f} ms")
Km = sq.view(n, n) @ sq.view(n, n)
print("check: |sqrt^2 - K|_max / |K|_max =", ((Km - K).abs().max() / K.abs().max()).item())
if os.environ.get("SW_CPU"):
    from oracle import am_driver as am
    Kh, Xh, yh = K.cpu().numpy(), X.T.contiguous().cpu().numpy(), y.cpu().numpy()
    t0 = time.perf_counter(); s_, i_ = am.calculateMMt_sqrt_and_sqrtinv(Kh); t1 = time.perf_counter()
    Hh = am.calculateH(Kh, ve, vg); Ph = am.calculateP(Hh, Xh); ah = am.calculate_reduced_a(vg, Ph, s_, yh); Vh = am.calculate_reduced_vara(Xh, ve, vg, Kh, s_); t2 = time.perf_counter()
    print(f"host (numpy/LAPACK, {os.cpu_count()} cores): sqrt_and_sqrtinv {1e3 * (t1 - t0):.0f} ms, per iteration {1e3 * (t2 - t1):.0f} ms")
    print("parity vs host: V", float(np.abs(V.view(n, n).T.cpu().numpy() - Vh).max() / np.abs(Vh).max()), " a", float(np.abs(a.cpu().numpy() - ah).max() / np.abs(ah).max()))
