"""GPU parity of the n x n algebra between two scans against the oracle's
restatements of the R functions (oracle/am_driver.py, which follow R/calculateMMt_sqrt_and_sqrtinv.R,
calculateH.R, calculateP.R, calculate_reduced_a.R, calculate_reduced_vara.R line by line).

Tolerance: 1e-9 of the largest entry of each result (these are LAPACK-shaped FP64 computations whose
operation order differs between libraries; the scan that consumes them is held to 1e-9)."""
import numpy as np
import pytest

from eagleeverything_b200 import synth

pytestmark = pytest.mark.gpu


def close(a, b, tol=1e-9):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() <= tol * np.abs(b).max()


@pytest.fixture(scope="module", params=[(150, 1), (333, 3), (1000, 2)], ids=["n150q1", "n333q3", "n1000q2"])
def problem(request):
    n, q = request.param
    rng = np.random.default_rng(n)
    G = synth.genotypes(n, 4 * n, seed=n).astype(np.float64) - 1.0
    MMt = G @ G.T
    K = MMt / MMt.max() + 0.95 * np.eye(n)                     # calcMMt.R:13
    X = np.column_stack([np.ones(n)] + [rng.integers(-1, 2, n).astype(float) for _ in range(q - 1)])
    y = rng.standard_normal(n) + 0.5 * G[:, 7]
    return dict(n=n, q=q, K=K, X=X, y=y, ve=0.7, vg=1.3)


def test_sqrt_and_sqrtinv(problem):
    from eagleeverything_b200 import api
    from oracle import am_driver as am
    K = problem["K"]
    res = api.calculateMMt_sqrt_and_sqrtinv(K)
    sq, inv = am.calculateMMt_sqrt_and_sqrtinv(K)
    assert close(res["sqrt_MMt"], sq) and close(res["inverse_sqrt_MMt"], inv)
    assert np.array_equal(res["inverse_sqrt_MMt"], res["inverse_sqrt_MMt"].T)          # chol2inv returns a mirrored triangle
    assert close(res["sqrt_MMt"] @ res["sqrt_MMt"], K) and close(res["sqrt_MMt"] @ res["inverse_sqrt_MMt"], np.eye(len(K)))


def test_not_positive_definite_and_asymmetric(problem):
    from eagleeverything_b200 import _lib, api
    K = problem["K"].copy()
    K[1, :] = K[0, :]; K[:, 1] = K[:, 0]; K[1, 1] = K[0, 0]      # two identical individuals: singular
    said = []
    assert api.calculateMMt_sqrt_and_sqrtinv(K, message=said.append) is None
    assert "not positive definite" in said[0] and "terminated with errors" in said[-1]
    K2 = problem["K"].copy()
    K2[0, 1] += 1e-12
    with pytest.raises(_lib.EagleGpuError, match="not a symmetric matrix"):
        api.calculateMMt_sqrt_and_sqrtinv(K2)


def test_H_P_reduced_a_reduced_vara(problem):
    from eagleeverything_b200 import api
    from oracle import am_driver as am
    K, X, y, ve, vg = (problem[k] for k in ("K", "X", "y", "ve", "vg"))
    H = api.calculateH(K, ve, vg)
    assert np.array_equal(H, am.calculateH(K, ve, vg))
    said = []
    assert api.calculateH(K, -1.0, vg, message=said.append) is None and "VarE cannot be negative" in said[0]
    P = api.calculateP(H, X)
    Pref = am.calculateP(H, X)
    assert close(P, Pref)
    assert np.abs(P @ X).max() <= 1e-9 * np.abs(P).max() * np.abs(X).max() * len(K)   # P annihilates the fixed effects
    sq, _ = am.calculateMMt_sqrt_and_sqrtinv(K)
    a = api.calculate_reduced_a(vg, Pref, sq, y)
    assert a.shape == (len(K), 1) and close(a.reshape(-1), am.calculate_reduced_a(vg, Pref, sq, y))
    V = api.calculate_reduced_vara(X, ve, vg, K, sq)
    assert close(V, am.calculate_reduced_vara(X, ve, vg, K, sq))


def test_scan_inputs_chain_picks_the_same_locus(problem):
    """find_qtl.R:5-45 end to end: GPU algebra -> GPU scan against oracle algebra -> oracle scan."""
    from eagleeverything_b200 import api
    from oracle import am_driver as am
    from oracle import np_oracle as npo
    K, X, y, ve, vg, n = (problem[k] for k in ("K", "X", "y", "ve", "vg", "n"))
    H = api.calculateH(K, ve, vg)
    P = api.calculateP(H, X)
    r = api.calculateMMt_sqrt_and_sqrtinv(K)
    hat_a = api.calculate_reduced_a(vg, P, r["sqrt_MMt"], y)
    V = api.calculate_reduced_vara(X, ve, vg, K, r["sqrt_MMt"])
    S0, V0, a0 = am.scan_inputs(K, K, X, y, ve, vg)
    assert close(r["inverse_sqrt_MMt"], S0) and close(V, V0) and close(hat_a.reshape(-1), np.asarray(a0).reshape(-1))
    G = synth.genotypes(n, 2000, seed=n + 1)
    Mt = (G.T.astype(np.float64) - 1.0)
    def scan(S, Vv, a):
        W = S @ (Vv @ S)
        aa = Mt @ (S @ np.asarray(a).reshape(-1))
        return am.pick_locus(aa, np.einsum("ij,jk,ik->i", Mt, W, Mt))[0]
    assert scan(r["inverse_sqrt_MMt"], V, hat_a) == scan(S0, V0, a0)


def test_forward_search_with_device_algebra(demo, synth_small):
    """The restated AM() loop with BOTH the hot path and the n x n algebra on the device: the selected-QTL sequence
    and the extBIC trace of the shipped demo data (golden) and of a synthetic set."""
    from eagleeverything_b200 import api
    from oracle import am_driver as am
    from oracle import eagle_oracle as eo
    z = demo["z"]
    r = am.AM(api, demo["geno"], z["trait1"], algebra=api)
    assert r["selected"] == list(z["am1_selected"]) and r["all_picked"] == list(z["am1_all_picked"])
    np.testing.assert_allclose(r["extBIC"], z["am1_extBIC"], rtol=1e-9)
    y, _ = synth.phenotype(synth_small["G"])
    rg = am.AM(api, synth_small["geno"], y, maxit=6, algebra=api)
    ro = am.AM(eo, synth_small["geno"], y, maxit=6)
    assert rg["all_picked"] == ro["all_picked"] and rg["selected"] == ro["selected"]


def test_emma_eigendecompositions(problem):
    """emma.eigen.L.wo.Z / emma.eigen.R.wo.Z: eigenvalues to 1e-9; eigenvectors are compared through what EMMA
    uses them for -- eta^2 = (U' y)^2 -- and as projectors, both invariant to the sign LAPACK happens to pick."""
    from eagleeverything_b200 import api
    from oracle import am_driver as am
    K, X, y, n, q = (problem[k] for k in ("K", "X", "y", "n", "q"))
    rl = api.emma_eigen_L_wo_Z(K)
    xi, Uo = am.r_eigen_sym(K)
    assert close(rl["values"], xi) and np.all(np.diff(rl["values"]) <= 0)
    assert close(rl["vectors"] @ np.diag(rl["values"]) @ rl["vectors"].T, K)
    assert close(api.emma_eigen_L_wo_Z(K, vectors=False)["values"], xi)
    rr = api.emma_eigen_R_wo_Z(K, X)
    lam, U = am.emma_eigen_R_wo_Z(K, X)
    assert rr["values"].shape == (n - q,) and rr["vectors"].shape == (n, n - q)
    assert close(rr["values"], lam)
    assert np.abs(np.abs(np.sum(rr["vectors"] * U, axis=0)) - 1.0).max() < 1e-6       # same vectors up to sign
    eg, eo_ = (rr["vectors"].T @ y) ** 2, (U.T @ y) ** 2
    assert np.abs(eg - eo_).max() <= 1e-7 * eo_.max()
    assert np.abs(rr["vectors"].T @ X).max() < 1e-8 * n                                # orthogonal to the fixed effects


# ---------------------------------------------------------------------------------------------------------------------
# The same algebra in the basis of eigen(K) (csrc/eigbasis.cu, csrc/secular.cuh): what am.AM_resident runs.
def test_secular_emma_eigen_on_device(problem):
    """eg_emma_eigen_R_wo_Z_eigbasis (q rank-one compressions solved by their secular equations, O(q n^2)) against the
    oracle's dense restatement of R/emma_eigen_R_wo_Z.R:7-20 + R/emma_REMLE.R:40: eigenvalues, and eta^2 through the
    order-free sums EMMA's likelihoods consume."""
    import ctypes as C
    from eagleeverything_b200 import _lib
    from oracle import am_driver as am
    lib = _lib.require_gpu()
    K, X, y, n, q = (problem[k] for k in ("K", "X", "y", "n", "q"))
    xi, U = am.r_eigen_sym(K)
    Xt, yt = np.asfortranarray(U.T @ X), np.ascontiguousarray(U.T @ y)
    vals, etas, st = np.empty(n - q), np.empty(n - q), (C.c_int64 * 4)()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    _lib.check(lib.eg_emma_eigen_R_wo_Z_eigbasis(dp(xi), dp(Xt), dp(yt), n, q, dp(vals), dp(etas), st))
    lam, Ur = am.emma_eigen_R_wo_Z(K, X)
    assert st[0] == q and st[2] < 60 and np.all(np.diff(vals) <= 0)
    np.testing.assert_allclose(vals, lam, rtol=0, atol=1e-12 * np.abs(xi).max())
    e_ref = (Ur.T @ y) ** 2
    for d in (1e-5, 1e-2, 1.0, 1e2):
        np.testing.assert_allclose((etas ** 2 / (vals + d)).sum(), (e_ref / (lam + d)).sum(), rtol=1e-11)
        np.testing.assert_allclose((etas ** 2 / (vals + d) ** 2).sum(), (e_ref / (lam + d) ** 2).sum(), rtol=1e-11)
    # a design matrix without full column rank is refused
    Xd = np.asfortranarray(np.column_stack([Xt[:, 0], 3.0 * Xt[:, 0]]))
    v2, e2 = np.empty(n - 2), np.empty(n - 2)
    assert lib.eg_emma_eigen_R_wo_Z_eigbasis(dp(xi), dp(Xd), dp(yt), n, 2, dp(v2), dp(e2), None) == _lib.EG_ERR_ARG
    assert b"rank deficient" in lib.eg_last_error()


@pytest.mark.parametrize("mode", ["auto", "i8", "f64"])
def test_scan_rhs_from_the_eigenbasis_equals_the_dense_route(problem, mode, monkeypatch):
    """eg_dev_scan_prepare_eig (W = U diag(w) U^T - E E^T: one digit-slice SYRK on the int8 tensor cores, or DSYRK) against
    eg_dev_scan_prepare fed with the oracle's dense S, V, a_hat (R/find_qtl.R:5-45): the packed right-hand side of the scan
    to 1e-9 of its largest entry, and a / var(a) of a marker panel."""
    import ctypes as C
    import torch
    from eagleeverything_b200 import _lib, am as pam, device
    from oracle import am_driver as am
    if mode != "auto":
        monkeypatch.setenv("EAGLE_PREP_MODE", mode)
    lib = device.init(0)
    K, X, y, ve, vg, n, q = (problem[k] for k in ("K", "X", "y", "ve", "vg", "n", "q"))
    S, V, hat_a = am.scan_inputs(K, K, X, y, ve, vg)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    Wp_ref = device.scan_prepare(cu(S), cu(V), cu(np.asarray(hat_a).reshape(-1)), n)
    xi, U = am.r_eigen_sym(K)
    w, Et, vt = pam.eigbasis_inputs(xi, U.T @ X, U.T @ y, ve, vg)
    dU, dUt = cu(U.T.copy()), cu(U.copy())       # row-major torch tensors: U^T row-major == U column-major
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    Wp = torch.empty(lib.eg_scan_wp_elems(n), dtype=torch.float64, device="cuda")
    work = torch.empty(n * q, dtype=torch.float64, device="cuda")
    work2 = torch.empty(n * n, dtype=torch.float64, device="cuda")
    d_w, d_Et, d_vt = cu(w), cu(Et.T.copy()), cu(vt)         # n x q column-major = (q, n) row-major
    _lib.check(lib.eg_dev_scan_prepare_eig(p(dU), p(dUt), n, p(d_w), p(d_Et), q, p(d_vt), p(work), p(work2), p(Wp), None))
    torch.cuda.synchronize()
    a, b = Wp.cpu().numpy(), Wp_ref.cpu().numpy()
    assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max()
    L = 1500
    G = synth.genotypes(n, L, seed=n + 3)
    img = torch.from_numpy(np.concatenate([synth.ascii_image(G).reshape(-1), np.zeros(64, np.uint8)])).cuda()
    kb, _ = device.decode_kb(img, L + 1, n, L)
    tT = device.transpose_kb(kb, n, L)
    a1, v1 = device.scan(tT, L, n, Wp)
    a0, v0 = device.scan(tT, L, n, Wp_ref)
    a1, v1, a0, v0 = (t.cpu().numpy() for t in (a1, v1, a0, v0))
    assert np.abs(a1 - a0).max() <= 1e-9 * np.abs(a0).max()
    assert np.abs(v1 - v0).max() <= 1e-9 * np.abs(v0).max()


def test_dense_and_eigenbasis_searches_agree(synth_small):
    """am.AM_resident (eigenbasis) against am.AM_resident_dense (R's formulas one for one on the device)."""
    import torch
    from eagleeverything_b200 import am as pam, device
    device.init(0)
    s = synth_small
    y, _ = synth.phenotype(s["G"])
    img = torch.from_numpy(np.concatenate([synth.ascii_image(s["G"]).reshape(-1), np.zeros(64, np.uint8)])).cuda()
    kb, _ = device.decode_kb(img, s["L"] + 1, s["n"], s["L"])
    tT = device.transpose_kb(kb, s["n"], s["L"])
    r1 = pam.AM_resident(kb, tT, s["n"], s["L"], y, maxit=6)
    r0 = pam.AM_resident_dense(kb, tT, s["n"], s["L"], y, maxit=6)
    assert r1["all_picked"] == r0["all_picked"] and r1["selected"] == r0["selected"]
    np.testing.assert_allclose(r1["extBIC"], r0["extBIC"], rtol=1e-9)
    assert r1["secular"]["max_root_iterations"] < 60


def test_emma_eigen_R_resident_with_many_fixed_effects():
    """q close to n / 2: the scratch layout of eg_dev_emma_eigen_R_wo_Z stays inside the documented buffer sizes."""
    import ctypes as C
    import torch
    from eagleeverything_b200 import _lib, device
    from oracle import am_driver as am
    lib = device.init(0)
    n, q = 12, 6
    rng = np.random.default_rng(0)
    A = rng.standard_normal((n, 3 * n))
    K = A @ A.T / (3 * n) + 0.95 * np.eye(n)
    X = np.column_stack([np.ones(n), rng.standard_normal((n, q - 1))])
    y = rng.standard_normal(n)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    f64 = dict(dtype=torch.float64, device="cuda")
    guard = 7.0
    dK, dX, dy = cu(K), cu(X.T.copy()), cu(y)
    vals, etas, U = torch.empty(n, **f64), torch.empty(n, **f64), torch.empty(n * n, **f64)
    w1, w2 = torch.full((n * n + 64,), guard, **f64), torch.full((n * n + 64,), guard, **f64)
    small = torch.full((2 * n * q + 2 * q * q + 64,), guard, **f64)
    _lib.check(lib.eg_dev_emma_eigen_R_wo_Z(p(dK), p(dX), p(dy), n, q, p(vals), p(etas), p(U), p(w1), p(w2), p(small), None))
    torch.cuda.synchronize()
    for buf, size in ((w1, n * n), (w2, n * n), (small, 2 * n * q + 2 * q * q)):
        assert bool((buf[size:] == guard).all().item()), "scratch overrun"
    lam, Ur = am.emma_eigen_R_wo_Z(K, X)
    np.testing.assert_allclose(vals[: n - q].cpu().numpy(), lam, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(etas[: n - q].cpu().numpy() ** 2, (Ur.T @ y) ** 2, rtol=1e-7, atol=1e-10)


def test_scan_from_the_cached_projection_equals_the_contraction(problem):
    """B = M^T U once (eg_dev_project_i8: the scan's int8 contraction in projection mode), then
    var(a)_j = sum_k w_k B_jk^2 - sum_c (E_c^T m_j)^2 and a_j = m_j^T v (eg_dev_bscan + eg_dev_gemv_i8) against the scan on
    the explicit W (eg_dev_scan_prepare_eig + eg_dev_scan), and B itself against an FP64 product."""
    import ctypes as C
    import torch
    from eagleeverything_b200 import _lib, am as pam, device
    from oracle import am_driver as am
    lib = device.init(0)
    K, X, y, ve, vg, n, q = (problem[k] for k in ("K", "X", "y", "ve", "vg", "n", "q"))
    xi, U = am.r_eigen_sym(K)
    w, Et, vt = pam.eigbasis_inputs(xi, U.T @ X, U.T @ y, ve, vg)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    f64 = dict(dtype=torch.float64, device="cuda")
    dU, dUt = cu(U.T.copy()), cu(U.copy())
    L = 1700
    G = synth.genotypes(n, L, seed=n + 9)
    G[:, 40] = G[:, 1500]                                                   # identical markers far apart
    img = torch.from_numpy(np.concatenate([synth.ascii_image(G).reshape(-1), np.zeros(64, np.uint8)])).cuda()
    kb, _ = device.decode_kb(img, L + 1, n, L)
    tT = device.transpose_kb(kb, n, L)
    # explicit W route
    d_w, d_Et, d_vt = cu(w), cu(Et.T.copy()), cu(vt)
    Wp = torch.empty(lib.eg_scan_wp_elems(n), **f64)
    _lib.check(lib.eg_dev_scan_prepare_eig(p(dU), p(dUt), n, p(d_w), p(d_Et), q, p(d_vt), p(torch.empty(n * q, **f64)),
                                           p(torch.empty(n * n, **f64)), p(Wp), None))
    a0, v0 = device.scan(tT, L, n, Wp)
    # cached projection route
    ldb = (n + 1) // 2 * 2
    B = torch.full((L, ldb), float("nan"), **f64)
    _lib.check(lib.eg_dev_project_i8(p(tT), L, n, tT.stride(0), p(dU), p(B), ldb, None))
    torch.cuda.synchronize()
    Mt = 1.0 - G.T.astype(np.float64)                                        # the store holds the negated genotype
    Bref = Mt @ U
    Bh = B.cpu().numpy()[:, :n]
    assert np.abs(Bh - Bref).max() <= 4 * n * np.finfo(float).eps * (np.abs(Mt) @ np.abs(U)).max()
    d_E, d_v = torch.empty((q, n), **f64), torch.empty(n, **f64)
    _lib.check(lib.eg_dev_eigbasis_apply(p(dU), n, p(d_Et), q, 0, p(d_E), None))
    _lib.check(lib.eg_dev_eigbasis_apply(p(dU), n, p(d_vt), 1, 0, p(d_v), None))
    e = torch.empty((q, L), **f64)
    a1, v1, tmpL = torch.empty(L, **f64), torch.empty(L, **f64), torch.empty(L, **f64)
    for c in range(q):
        _lib.check(lib.eg_dev_gemv_i8(p(tT), L, n, tT.stride(0), p(d_E[c]), 1.0, p(e[c]), None))
    _lib.check(lib.eg_dev_gemv_i8(p(tT), L, n, tT.stride(0), p(d_v), 1.0, p(a1), None))
    _lib.check(lib.eg_dev_bscan(p(B), L, n, ldb, p(d_w), p(e), q, p(tmpL), p(v1), None))
    a0, v0, a1, v1 = (t.cpu().numpy() for t in (a0, v0, a1, v1))
    E = U @ Et
    cond_v = ((np.abs(Mt) @ np.abs(U)) ** 2 * w).sum(1) + ((Mt @ E) ** 2).sum(1)
    assert (np.abs(v1 - v0) <= 1e-9 * np.abs(v0) + 4 * (n + 10) * np.finfo(float).eps * cond_v).all()
    assert np.abs(a1 - a0).max() <= 1e-12 * np.abs(a0).max()
    assert v1[40] == v1[1500] and a1[40] == a1[1500]                        # identical rows, identical bits


def test_search_routes_agree(synth_small):
    """am.AM_resident with the cached projection against the contraction per iteration, without and with Z."""
    import torch
    from eagleeverything_b200 import am as pam, device
    device.init(0)
    s = synth_small
    y, _ = synth.phenotype(s["G"])
    img = torch.from_numpy(np.concatenate([synth.ascii_image(s["G"]).reshape(-1), np.zeros(64, np.uint8)])).cuda()
    kb, _ = device.decode_kb(img, s["L"] + 1, s["n"], s["L"])
    tT = device.transpose_kb(kb, s["n"], s["L"])
    r1 = pam.AM_resident(kb, tT, s["n"], s["L"], y, maxit=6, bcache=True)
    r0 = pam.AM_resident(kb, tT, s["n"], s["L"], y, maxit=6, bcache=False)
    assert "cached projection" in r1["scan_route"] and "contraction" in r0["scan_route"]
    assert r1["all_picked"] == r0["all_picked"] and r1["selected"] == r0["selected"]
    np.testing.assert_allclose(r1["extBIC"], r0["extBIC"], rtol=1e-12)
    rng = np.random.default_rng(2)
    idx = np.concatenate([np.arange(s["n"]), rng.integers(0, s["n"], 90)])
    yz = y[idx] + 0.3 * rng.standard_normal(len(idx))
    z1 = pam.AM_resident(kb, tT, s["n"], s["L"], yz, maxit=4, Z=idx, bcache=True)
    z0 = pam.AM_resident(kb, tT, s["n"], s["L"], yz, maxit=4, Z=idx, bcache=False)
    assert z1["all_picked"] == z0["all_picked"]
    np.testing.assert_allclose(z1["extBIC"], z0["extBIC"], rtol=1e-12)
