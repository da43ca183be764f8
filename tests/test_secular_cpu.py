"""CPU checks of the eigenbasis route of the forward search (csrc/secular.cuh, am.eigbasis_inputs).

The arithmetic and host orchestration of the secular solver live in a host + device header; here it is compiled with a
plain-loop back end (tests/csrc/secular_host.cpp -- the CUDA back end of csrc/eigbasis.cu runs the same functions inside
its kernels) and checked against a dense LAPACK eigendecomposition of S (K + I) S, i.e. against the oracle's
restatement of R/emma_eigen_R_wo_Z.R; the eigenbasis form of the scan's inputs is checked against the oracle's
restatement of R's dense formulas (find_qtl.R:5-45), and a whole forward search assembled from these pieces reproduces
the golden demo results.  The GPU suite repeats the comparisons with the kernels (tests/test_gpu_algebra.py, test_am.py).
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from eagleeverything_b200 import am
from oracle import am_driver as oam

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def sec():
    src = os.path.join(ROOT, "tests", "csrc", "secular_host.cpp")
    hdr = os.path.join(ROOT, "eagleeverything_b200", "csrc", "secular.cuh")
    out = os.path.join(ROOT, "tests", "_build", "libsecular_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src])
    lib = C.CDLL(out)

    def compress(xi, Xt, yt):
        n, q = Xt.shape
        Xf = np.asfortranarray(Xt, dtype=np.float64)
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        yt = np.ascontiguousarray(yt, dtype=np.float64)
        vals, etas, st = np.empty(n - q), np.empty(n - q), (C.c_int64 * 4)()
        rc = lib.th_secular_compress(C.c_int64(n), q, xi.ctypes.data_as(_dp), Xf.ctypes.data_as(_dp), yt.ctypes.data_as(_dp),
                                     vals.ctypes.data_as(_dp), etas.ctypes.data_as(_dp), st)
        return rc, vals, etas, dict(steps=st[0], deflated=st[1], max_iters=st[2], roots=st[3])
    return compress


def _spectral_sums(vals, etasq, deltas=(1e-5, 1e-2, 1.0, 1e2)):
    """what EMMA's likelihoods consume (R/emma_REMLE.R:44-76): order-free sums over the eigenpairs"""
    return np.array([[(etasq / (vals + d)).sum(), (etasq / (vals + d) ** 2).sum(), np.log(vals + d).sum(), (1 / (vals + d)).sum()]
                     for d in deltas])


def _K_of(G):
    M = G.astype(np.float64) - 1.0
    MMt = M @ M.T
    return MMt / MMt.max() + np.diag(np.full(G.shape[0], 0.95))


@pytest.mark.parametrize("n,L,q", [(40, 300, 1), (150, 900, 3), (333, 2000, 6), (260, 30, 3)])
def test_secular_equals_dense_eigen(sec, n, L, q):
    rng = np.random.default_rng(n + q)
    G = rng.integers(0, 3, (n, L))
    if n == 150:
        G[7] = G[8]          # identical individuals: repeated eigenvalues of K
    K = _K_of(G)
    X = np.column_stack([np.ones(n)] + [G[:, j].astype(float) - 1 for j in range(q - 1)])
    y = rng.standard_normal(n) + (G[:, 3] - 1)
    xi, U = oam.r_eigen_sym(K)
    rc, vals, etas, st = sec(xi, U.T @ X, U.T @ y)
    assert rc == 0 and st["steps"] == q and st["max_iters"] < 60
    lam, Ur = oam.emma_eigen_R_wo_Z(K, X)                     # R/emma_eigen_R_wo_Z.R:7-20 (dense)
    np.testing.assert_allclose(np.sort(vals), np.sort(lam), rtol=0, atol=1e-12 * np.abs(xi).max())
    np.testing.assert_allclose(_spectral_sums(vals, etas ** 2), _spectral_sums(lam, (Ur.T @ y) ** 2), rtol=1e-11)
    assert vals[0] >= vals[-1]                                # decreasing, as eigen()
    if L < n:
        assert st["deflated"] > 0                             # n > L: K has a repeated eigenvalue 0.95


def test_secular_graded_and_clustered_poles(sec):
    rng = np.random.default_rng(5)
    n = 300
    xi = np.sort(np.concatenate([0.95 + 10 ** rng.uniform(-14, -1, n // 2), 0.95 + rng.uniform(0, 50, n // 2)]))
    xi[40:48] = xi[40]
    for q in (1, 4):
        Xt = rng.standard_normal((n, q)) * 10 ** rng.uniform(-16, 0, (n, 1))
        yt = rng.standard_normal(n)
        rc, vals, etas, st = sec(xi, Xt, yt)
        assert rc == 0
        Q, _ = np.linalg.qr(Xt)
        P = np.eye(n) - Q @ Q.T
        T = P @ np.diag(xi) @ P
        w, Vv = np.linalg.eigh((T + T.T) / 2)
        keep = np.argsort(np.linalg.norm(Q.T @ Vv, axis=0))[: n - q]
        np.testing.assert_allclose(np.sort(vals), np.sort(w[keep]), rtol=0, atol=1e-12 * xi.max())
        np.testing.assert_allclose(_spectral_sums(vals, etas ** 2), _spectral_sums(w[keep], (Vv[:, keep].T @ yt) ** 2), rtol=1e-10)
        np.testing.assert_allclose((etas ** 2).sum(), yt @ P @ yt, rtol=1e-12)


def test_secular_rank_deficient_design_is_refused(sec):
    rng = np.random.default_rng(2)
    n = 60
    xi = 0.95 + rng.uniform(0, 3, n)
    x = rng.standard_normal(n)
    rc, *_ = sec(xi, np.column_stack([x, 2 * x]), rng.standard_normal(n))
    assert rc == 1


def test_eigbasis_inputs_equal_the_dense_formulas(demo):
    """W = S V S and v = S a_hat (src/calculate_a_and_vara_rcpp.cpp:90,97-98) from R's dense formulas (oracle) against
    the eigenbasis form the resident search feeds the scan with."""
    G = demo["G"]
    n = demo["n"]
    K = _K_of(G)
    y = np.asarray(demo["z"]["trait1"], dtype=np.float64)
    X = np.column_stack([np.ones(n), G[:, 2206].astype(float) - 1, G[:, 4502].astype(float) - 1])
    ve, vg = 0.8, 19.7
    S, V, hat_a = oam.scan_inputs(K, oam.r_chol2inv_chol(K), X, y, ve, vg)
    W_ref, v_ref = S @ (V @ S), S @ hat_a
    xi, U = oam.r_eigen_sym(K)
    w, Et, vt = am.eigbasis_inputs(xi, U.T @ X, U.T @ y, ve, vg)
    E = U @ Et
    W = (U * w) @ U.T - E @ E.T
    assert np.abs(W - W_ref).max() <= 1e-9 * np.abs(W_ref).max()
    assert np.abs(U @ vt - v_ref).max() <= 1e-9 * np.abs(v_ref).max()
    # and the quantities the scan derives from them, for every demo marker
    M = G.astype(np.float64) - 1.0
    vara, vara_ref = np.einsum("ij,ij->j", M, W @ M), np.einsum("ij,ij->j", M, W_ref @ M)
    a, a_ref = M.T @ (U @ vt), M.T @ v_ref
    big = np.abs(vara_ref) > 1e-6 * np.abs(vara_ref).max()
    np.testing.assert_allclose(vara[big], vara_ref[big], rtol=1e-8)
    np.testing.assert_allclose(a[big], a_ref[big], rtol=1e-8, atol=1e-9 * np.abs(a_ref).max())


def test_forward_search_in_the_eigenbasis_reproduces_the_golden_demo(sec, demo):
    """AM_resident's arithmetic with numpy standing in for the device products: secular EMMA + eigenbasis inputs give the
    golden selected loci and extBIC trace (tests/golden/demo.npz)."""
    z, G, n, L = demo["z"], demo["G"], demo["n"], demo["L"]
    M = G.astype(np.float64) - 1.0
    xi, U = oam.r_eigen_sym(_K_of(G))

    def search(y, X0):
        y = np.asarray(y, dtype=np.float64)
        X = X0.copy()
        emma = am._Emma(None, dict(emma_eigen_s=0.0))
        emma._xi = xi
        yt, picked, ext = U.T @ y, [], []
        for _ in range(20):
            q = X.shape[1]
            rc, lam, et, _st = sec(xi, U.T @ X, yt)
            assert rc == 0
            emma._last = (q, lam, et * et)
            vc = emma.REMLE(y, X)
            ml = emma.MLE(y, X, llim=-100.0, ulim=100.0)
            ext.append(-2 * ml["ML"] + (q + 1) * math.log(n) + 2 * am._lchoose(L, q - 1))
            if int(np.flatnonzero(np.asarray(ext) == min(ext))[0]) != len(ext) - 1:
                break
            w, Et, vt = am.eigbasis_inputs(xi, U.T @ X, yt, vc["ve"], vc["vg"])
            E = U @ Et
            T = U.T @ M
            vara = (w[:, None] * T * T).sum(0) - ((E.T @ M) ** 2).sum(0)
            a = M.T @ (U @ vt)
            picked.append(oam.pick_locus(a, vara)[0])
            X = np.column_stack([X, M[:, picked[-1] - 1]])
        return picked, ext

    picked, ext = search(z["trait1"], np.ones((n, 1)))
    assert picked == list(z["am1_all_picked"])
    np.testing.assert_allclose(ext, z["am1_extBIC"], rtol=1e-8)
    picked2, ext2 = search(z["trait2"], np.column_stack([np.ones(n), z["pc1"], z["pc2"]]))
    assert picked2 == list(z["am2_all_picked"])
    np.testing.assert_allclose(ext2, z["am2_extBIC"], rtol=1e-8)


# ---------------------------------------------------------------------------------------------------------------------
# Repeated measures (Z): EMMA's Z branches and the Z-aware scan inputs in the basis of eigen(C^1/2 K C^1/2), C = Z'Z
def _z_problem(seed=3, t=70, L=500, nrec=110, q=2):
    rng = np.random.default_rng(seed)
    G = rng.integers(0, 3, (t, L))
    idx = np.concatenate([np.arange(t), rng.integers(0, t, nrec - t)])
    rng.shuffle(idx)
    Z = np.zeros((nrec, t))
    Z[np.arange(nrec), idx] = 1
    X = np.column_stack([np.ones(nrec)] + [rng.standard_normal(nrec) for _ in range(q - 1)])
    y = rng.standard_normal(nrec) + 0.9 * (Z @ (G[:, 11] - 1.0))
    return G, Z, idx, X, y


def _z_basis(K, idx, t):
    cnt = np.bincount(idx, minlength=t).astype(np.float64)
    sc = np.sqrt(cnt)
    xi, Hm = oam.r_eigen_sym(K * sc[:, None] * sc[None, :])
    return cnt, xi, sc[:, None] * Hm              # A = C^1/2 H


def _z_coords(A, cnt, idx, v):
    top = A.T @ (np.bincount(idx, weights=v, minlength=len(cnt)) / cnt)
    return top, v - ((A @ top) / cnt)[idx]


def test_emma_with_Z_through_the_secular_solve_equals_the_dense_restatement(sec):
    G, Z, idx, X, y = _z_problem()
    t, q, nrec = G.shape[0], X.shape[1], len(y)
    K = _K_of(G)
    cnt, xi, A = _z_basis(K, idx, t)
    tops = [_z_coords(A, cnt, idx, X[:, c]) for c in range(q)]
    yt, ry = _z_coords(A, cnt, idx, y)
    Xt = np.column_stack([a for a, _ in tops])
    _, Rr = np.linalg.qr(np.column_stack([r for _, r in tops] + [ry]))
    rc, vals, etas, _st = sec(np.concatenate([xi, np.zeros(q + 1)]), np.vstack([Xt, Rr[:, :q]]), np.concatenate([yt, Rr[:, q]]))
    assert rc == 0
    lam, e1sq, e2sq = vals[: t - q], etas[: t - q] ** 2, float((etas[t - q:] ** 2).sum())
    lam_ref, U_ref = oam.emma_eigen_R_w_Z(Z, K, X)                     # R/emma_eigen_R_w_Z.R:2-23 (dense, non-symmetric eigen)
    np.testing.assert_allclose(np.sort(lam), np.sort(lam_ref), rtol=1e-9, atol=1e-11)
    e_ref = U_ref.T @ y
    np.testing.assert_allclose(e2sq, (e_ref[t - q:] ** 2).sum(), rtol=1e-9)
    np.testing.assert_allclose(np.sort(xi), np.sort(oam.emma_eigen_L_w_Z(Z, K)), rtol=1e-9)
    r, ro = am.emma_w_Z(nrec, t, q, lam, e1sq, e2sq), oam.emma_REMLE_Z(y, X, K, Z)
    m = am.emma_w_Z(nrec, t, q, lam, e1sq, e2sq, xi=xi, llim=-100.0, ulim=100.0)
    mo = oam.emma_MLE_Z(y, X, K, Z, llim=-100.0, ulim=100.0)
    for k in ("REML", "delta", "ve", "vg"):
        np.testing.assert_allclose(r[k], ro[k], rtol=1e-7)
    for k in ("ML", "delta", "ve", "vg"):
        np.testing.assert_allclose(m[k], mo[k], rtol=1e-7)


def test_Z_aware_scan_inputs_equal_the_dense_model():
    G, Z, idx, X, y = _z_problem(seed=8, q=3)
    t = G.shape[0]
    K = _K_of(G)
    cnt, xi, A = _z_basis(K, idx, t)
    Xt = np.column_stack([_z_coords(A, cnt, idx, X[:, c])[0] for c in range(X.shape[1])])
    yt = _z_coords(A, cnt, idx, y)[0]
    ve, vg = 0.6, 1.4
    w, Et, vt = am.eigbasis_inputs(xi, Xt, yt, ve, vg, XtX=X.T @ X, Xty=X.T @ y)
    E = A @ Et
    W, v = (A * w) @ A.T - E @ E.T, A @ vt
    W_ref, v_ref = oam.scan_inputs_Z(K, Z, X, y, ve, vg)                 # H = ve I + vg Z K Z', dense
    assert np.abs(W - W_ref).max() <= 1e-9 * np.abs(W_ref).max()
    assert np.abs(v - v_ref).max() <= 1e-9 * np.abs(v_ref).max()
    # Z = I falls back to the formulas without Z
    n = t
    Xi, yi = X[:n], y[:n]
    xi0, U0 = oam.r_eigen_sym(K)
    a0 = am.eigbasis_inputs(xi0, U0.T @ Xi, U0.T @ yi, ve, vg)
    a1 = am.eigbasis_inputs(xi0, U0.T @ Xi, U0.T @ yi, ve, vg, XtX=Xi.T @ Xi, Xty=Xi.T @ yi)
    for p0, p1 in zip(a0, a1):
        np.testing.assert_allclose(p0, p1, rtol=1e-9, atol=1e-12)
    with pytest.raises(ValueError):
        am.z_index(np.delete(Z, 3, axis=1), t)
    assert np.array_equal(am.z_index(Z, t), idx)
