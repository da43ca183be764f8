"""Restatement of the R side of AM() (the caller of the hot path).  TEST INFRASTRUCTURE ONLY.

R is not installed in this image, so the forward-selection driver that surrounds the three
hot-path exports is restated in numpy so that a full multi-locus search can be run with EITHER
the CPU oracle or the GPU library plugged in as `backend` (an object exposing
calculateMMt_rcpp / calculate_a_and_vara_rcpp / extract_geno_rcpp with the reference's
argument lists).  In a real deployment none of this exists: the untouched R package calls the
Rcpp exports.  Paths below are relative to /root/reference/MyPackage/Eagle/R/.

Single trait, no Z matrix (Z never reaches find_qtl in this snapshot: AM.R:450-452,
find_qtl.R:1-2, calculateP.R:22-25).
"""
from __future__ import annotations

import math
import struct

import numpy as np
from scipy.linalg import lapack
from scipy.special import gammaln

#: R's NA_real_ (NaN with low word 1954): the reference tests it with R_IsNA, which a plain NaN fails
NA = struct.unpack("<d", struct.pack("<Q", 0x7FF00000000007A2))[0]


# ----------------------------------------------------------------------------- base-R helpers
def r_eigen_sym(A):
    """eigen(A, symmetric=TRUE): eigenvalues in DEcreasing order."""
    w, v = np.linalg.eigh(A)
    return w[::-1].copy(), v[:, ::-1].copy()


def r_chol2inv_chol(A):
    """chol2inv(chol(A)) == LAPACK dpotrf + dpotri."""
    c, info = lapack.dpotrf(np.asfortranarray(A), lower=0)
    if info != 0:
        raise np.linalg.LinAlgError(f"chol: leading minor of order {info} is not positive definite")
    inv, info = lapack.dpotri(c, lower=0)
    if info != 0:
        raise np.linalg.LinAlgError("dpotri failed")
    iu = np.triu_indices_from(inv, 1)
    inv[(iu[1], iu[0])] = inv[iu]
    return inv


def r_uniroot(f, lower, upper, tol=np.finfo(float).eps ** 0.25, maxiter=1000):
    """uniroot() -> R_zeroin2: the Brent/Dekker 'zeroin' of Forsythe, Malcolm & Moler
    (published algorithm, netlib zeroin.c), default tol = .Machine$double.eps^0.25."""
    a, b = float(lower), float(upper)
    fa, fb = f(a), f(b)
    c, fc = a, fa
    EPS = np.finfo(float).eps
    if fa == 0.0:
        return a
    if fb == 0.0:
        return b
    it = maxiter + 1
    while it > 0:
        it -= 1
        prev_step = b - a
        if abs(fc) < abs(fb):
            a, b, c = b, c, b
            fa, fb, fc = fb, fc, fb
        tol_act = 2 * EPS * abs(b) + tol / 2
        new_step = (c - b) / 2
        if abs(new_step) <= tol_act or fb == 0.0:
            return b
        if abs(prev_step) >= tol_act and abs(fa) > abs(fb):
            cb = c - b
            if a == c:
                t1 = fb / fa
                p = cb * t1
                q = 1.0 - t1
            else:
                q = fa / fc
                t1 = fb / fc
                t2 = fb / fa
                p = t2 * (cb * q * (q - t1) - (b - a) * (t1 - 1.0))
                q = (q - 1.0) * (t1 - 1.0) * (t2 - 1.0)
            if p > 0:
                q = -q
            else:
                p = -p
            if p < (0.75 * cb * q - abs(tol_act * q) / 2) and p < abs(prev_step * q / 2):
                new_step = p / q
        if abs(new_step) < tol_act:
            new_step = tol_act if new_step > 0 else -tol_act
        a, fa = b, fb
        b += new_step
        fb = f(b)
        if (fb > 0 and fc > 0) or (fb < 0 and fc < 0):
            c, fc = a, fa
    return b


def lchoose(n, k):
    return gammaln(n + 1) - gammaln(k + 1) - gammaln(n - k + 1)


# ----------------------------------------------------------------------------- EMMA (emma_*.R)
def emma_eigen_R_wo_Z(K, X):
    """emma_eigen_R_wo_Z.R:4-20"""
    n, q = X.shape
    dn = np.eye(n)
    S = dn - X @ np.linalg.inv(X.T @ X) @ X.T
    w, v = r_eigen_sym(S @ (K + dn) @ S)
    return w[: n - q] - 1.0, v[:, : n - q]


def _grid_opt(dLL, logdelta, llim, ulim, esp, LLfun, dLLfun):
    """emma_REMLE.R:54-76 / emma_MLE.R:34-56 (identical control flow)."""
    m = len(logdelta)
    opt_ld, opt_ll = [], []
    if dLL[0] < esp:
        opt_ld.append(llim)
        opt_ll.append(LLfun(llim))
    if dLL[m - 2] > 0 - esp:  # R: dLL[m - 1], 1-based
        opt_ld.append(ulim)
        opt_ll.append(LLfun(ulim))
    for i in range(m - 1):
        if dLL[i] * dLL[i + 1] < 0 - esp * esp and dLL[i] > 0 and dLL[i + 1] < 0:
            r = r_uniroot(dLLfun, logdelta[i], logdelta[i + 1])
            opt_ld.append(r)
            opt_ll.append(LLfun(r))
    k = int(np.argmax(opt_ll))  # which.max: first maximum
    return math.exp(opt_ld[k]), opt_ll[k]


def emma_REMLE(y, X, K, ngrids=100, llim=-10.0, ulim=10.0, esp=1e-10):
    """emma_REMLE.R:27-131 (Z = NULL branch)."""
    n, q = len(y), X.shape[1]
    if np.linalg.det(X.T @ X) == 0:
        return dict(REML=0.0, delta=0.0, ve=0.0, vg=0.0)
    lam, U = emma_eigen_R_wo_Z(K, X)
    etas = U.T @ y
    etasq = etas * etas
    logdelta = np.arange(ngrids + 1) / ngrids * (ulim - llim) + llim
    delta = np.exp(logdelta)
    Lambdas = lam[:, None] + delta[None, :]
    Etasq = etasq[:, None]
    dLL = 0.5 * delta * ((n - q) * (Etasq / (Lambdas * Lambdas)).sum(0) / (Etasq / Lambdas).sum(0)
                         - (1.0 / Lambdas).sum(0))

    def LLfun(ld):  # emma.delta.REML.LL.wo.Z  :144-150
        nq = len(etas)
        d = math.exp(ld)
        return 0.5 * (nq * (math.log(nq / (2 * math.pi)) - 1 - math.log((etasq / (lam + d)).sum()))
                      - np.log(lam + d).sum())

    def dLLfun(ld):  # emma.delta.REML.dLL.wo.Z :134-142
        nq = len(etas)
        d = math.exp(ld)
        ldel = lam + d
        return 0.5 * (nq * (etasq / (ldel * ldel)).sum() / (etasq / ldel).sum() - (1.0 / ldel).sum())

    maxdelta, maxLL = _grid_opt(dLL, logdelta, llim, ulim, esp, LLfun, dLLfun)
    maxva = (etasq / (lam + maxdelta)).sum() / (n - q)
    return dict(REML=maxLL, delta=maxdelta, ve=maxva * maxdelta, vg=maxva)


def emma_MLE(y, X, K, ngrids=100, llim=-10.0, ulim=10.0, esp=1e-10):
    """emma_MLE.R:2-117 (Z = NULL branch)."""
    n, q = len(y), X.shape[1]
    if np.linalg.det(X.T @ X) == 0:
        return dict(ML=0.0, delta=0.0, ve=0.0, vg=0.0)
    xi, _ = r_eigen_sym(K)  # emma_eigen_L_wo_Z.R:9
    lam, U = emma_eigen_R_wo_Z(K, X)
    etas = U.T @ y
    etasq = etas * etas
    logdelta = np.arange(ngrids + 1) / ngrids * (ulim - llim) + llim
    delta = np.exp(logdelta)
    Lambdas = lam[:, None] + delta[None, :]
    Xis = xi[:, None] + delta[None, :]
    Etasq = etasq[:, None]
    dLL = 0.5 * delta * (n * (Etasq / (Lambdas * Lambdas)).sum(0) / (Etasq / Lambdas).sum(0)
                         - (1.0 / Xis).sum(0))

    def LLfun(ld):  # emma_delta_ML_LL_wo_Z.R:2-8
        nn = len(xi)
        d = math.exp(ld)
        return 0.5 * (nn * (math.log(nn / (2 * math.pi)) - 1 - math.log((etasq / (lam + d)).sum()))
                      - np.log(xi + d).sum())

    def dLLfun(ld):  # emma_misc.R:10-18
        nn = len(xi)
        d = math.exp(ld)
        ldel = lam + d
        return 0.5 * (nn * (etasq / (ldel * ldel)).sum() / (etasq / ldel).sum() - (1.0 / (xi + d)).sum())

    maxdelta, maxLL = _grid_opt(dLL, logdelta, llim, ulim, esp, LLfun, dLLfun)
    maxva = (etasq / (lam + maxdelta)).sum() / n
    return dict(ML=maxLL, delta=maxdelta, ve=maxva * maxdelta, vg=maxva)


# ----------------------------------------------------------------------------- EMMA with an incidence matrix Z
def _r_eigen_nonsym(A):
    """eigen(A, symmetric=FALSE): eigenvalues by decreasing modulus; for the matrices of emma.eigen.*.w.Z (similar to
    symmetric positive semi-definite ones) they are real; the reference keeps Re() (emma_eigen_R_w_Z.R:14-17)."""
    w, v = np.linalg.eig(A)
    w, v = np.real(w), np.real(v)
    o = np.argsort(-np.abs(w), kind="stable")
    return w[o], v[:, o]


def emma_eigen_L_w_Z(Z, K, complete=True):
    """emma_eigen_L_w_Z.R:2-14 (values only are consumed by emma.MLE)."""
    if not complete:
        vids = Z.sum(0) > 0
        Z, K = Z[:, vids], K[np.ix_(vids, vids)]
    w, _ = _r_eigen_nonsym(K @ (Z.T @ Z))
    return w


def emma_eigen_R_w_Z(Z, K, X, complete=True):
    """emma_eigen_R_w_Z.R:2-23 -> (values[1:(t-q)], vectors n x (n-q): the first t-q pair with the values)."""
    if not complete:
        vids = Z.sum(0) > 0
        Z, K = Z[:, vids], K[np.ix_(vids, vids)]
    n, t = Z.shape
    q = X.shape[1]
    SZ = Z - X @ np.linalg.solve(X.T @ X, X.T @ Z)
    w, v = _r_eigen_nonsym(K @ (Z.T @ SZ))
    qrX, _ = np.linalg.qr(X)
    Q, _ = np.linalg.qr(np.column_stack([SZ @ v[:, : t - q], qrX]), mode="complete")
    return w[: t - q], Q[:, list(range(t - q)) + list(range(t, n))]


def _emma_w_Z(y, X, K, Z, ngrids, llim, ulim, esp, ml):
    """The Z branches of emma_REMLE.R:78-117 and emma_MLE.R:57-105 (ml: the ML form, with eigen(K Z'Z) and n in
    place of n - q).  Note: emma.MLE passes `ngpu` in the position of `complete` (emma_MLE.R:60,63), i.e. complete = 0:
    individuals without a record are dropped there; with every individual measured the two forms coincide."""
    n, q = len(y), X.shape[1]
    t = K.shape[0]
    if np.linalg.det(X.T @ X) == 0:
        return dict(ML=0.0, delta=0.0, ve=0.0, vg=0.0) if ml else dict(REML=0.0, delta=0.0, ve=0.0, vg=0.0)
    lam, U = emma_eigen_R_w_Z(Z, K, X, complete=not ml)
    xi = emma_eigen_L_w_Z(Z, K, complete=False) if ml else None
    etas = U.T @ y
    e1, e2sq = etas[: t - q], float((etas[t - q:] ** 2).sum())
    e1sq = e1 * e1
    logdelta = np.arange(ngrids + 1) / ngrids * (ulim - llim) + llim
    delta = np.exp(logdelta)
    Lambdas = lam[:, None] + delta[None, :]
    E = e1sq[:, None]
    if ml:
        Xis = xi[:, None] + delta[None, :]
        dLL = 0.5 * delta * (n * ((E / (Lambdas * Lambdas)).sum(0) + e2sq / (delta * delta)) / ((E / Lambdas).sum(0) + e2sq / delta)
                             - ((1.0 / Xis).sum(0) + (n - t) / delta))

        def LLfun(ld):   # emma.delta.ML.LL.w.Z (emma_misc.R)
            d = math.exp(ld)
            return 0.5 * (n * (math.log(n / (2 * math.pi)) - 1 - math.log((e1sq / (lam + d)).sum() + e2sq / d))
                          - (np.log(xi + d).sum() + (n - len(xi)) * ld))

        def dLLfun(ld):  # emma_delta_ML_dLL_w_Z.R:1-10
            d = math.exp(ld)
            ldel = lam + d
            return 0.5 * (n * ((e1sq / (ldel * ldel)).sum() + e2sq / (d * d)) / ((e1sq / ldel).sum() + e2sq / d)
                          - ((1.0 / (xi + d)).sum() + (n - len(xi)) / d))
        denom = n
    else:
        dLL = 0.5 * delta * ((n - q) * ((E / (Lambdas * Lambdas)).sum(0) + e2sq / (delta * delta)) / ((E / Lambdas).sum(0) + e2sq / delta)
                             - ((1.0 / Lambdas).sum(0) + (n - t) / delta))

        def LLfun(ld):   # emma.delta.REML.LL.w.Z
            tq = len(e1)
            nq = n - t + tq
            d = math.exp(ld)
            return 0.5 * (nq * (math.log(nq / (2 * math.pi)) - 1 - math.log((e1sq / (lam + d)).sum() + e2sq / d))
                          - (np.log(lam + d).sum() + (n - t) * ld))

        def dLLfun(ld):  # emma.delta.REML.dLL.w.Z
            tq = len(e1)
            nq = n - t + tq
            d = math.exp(ld)
            ldel = lam + d
            return 0.5 * (nq * ((e1sq / (ldel * ldel)).sum() + e2sq / (d * d)) / ((e1sq / ldel).sum() + e2sq / d)
                          - ((1.0 / ldel).sum() + (n - t) / d))
        denom = n - q
    maxdelta, maxLL = _grid_opt(dLL, logdelta, llim, ulim, esp, LLfun, dLLfun)
    maxva = ((e1sq / (lam + maxdelta)).sum() + e2sq / maxdelta) / denom
    return {("ML" if ml else "REML"): maxLL, "delta": maxdelta, "ve": maxva * maxdelta, "vg": maxva}


def emma_REMLE_Z(y, X, K, Z, ngrids=100, llim=-10.0, ulim=10.0, esp=1e-10):
    return _emma_w_Z(y, X, K, Z, ngrids, llim, ulim, esp, ml=False)


def emma_MLE_Z(y, X, K, Z, ngrids=100, llim=-10.0, ulim=10.0, esp=1e-10):
    return _emma_w_Z(y, X, K, Z, ngrids, llim, ulim, esp, ml=True)


def scan_inputs_Z(K, Z, X, y, ve, vg):
    """What a Z-aware find_qtl feeds the scan with (NOT in the reference snapshot, where Z never
    reaches find_qtl: AM.R:450-452, find_qtl.R:1-2, calculateP.R:22-25).  Model: y = X b + Z u + e, u ~ N(0, vg K).
    H = ve I + vg Z K Z' (calculateH.R:36 with Z K Z' for K), P as calculateP.R:27-28; the marker scores are
    a = vg M' Z' P y and var(a) = vg^2 diag(M' Z' P Z M), i.e. the scan of calculate_a_and_vara_rcpp with
    v = vg Z' P y and W = vg^2 Z' P Z (t x t) in place of S a_hat and S V S.  Dense evaluation."""
    n = Z.shape[0]
    H = ve * np.eye(n) + vg * (Z @ K @ Z.T)
    P = calculateP(H, X)
    return vg * vg * (Z.T @ P @ Z), vg * (Z.T @ (P @ y))


def AM_Z(backend_M, y, X0, Z, L, maxit=20):
    """The forward search with repeated measures: AM()'s loop control (AM.R:395-492), EMMA with Z as the reference has it
    (AM.R:428,436), the Z-aware scan inputs above; dense numpy throughout.  backend_M: the t x L matrix of -1/0/1
    genotypes."""
    y = np.asarray(y, dtype=np.float64)
    M = np.asarray(backend_M, dtype=np.float64)
    n = len(y)
    X = np.ones((n, 1)) if X0 is None else np.asarray(X0, dtype=np.float64)
    MMt = M @ M.T
    K = MMt / MMt.max() + np.diag(np.full(M.shape[0], 0.95))
    picked, extBIC, vc = [], [], None
    itnum, cont = 1, True
    while cont:
        vc = emma_REMLE_Z(y, X, K, Z)
        ml = emma_MLE_Z(y, X, K, Z, llim=-100.0, ulim=100.0)
        extBIC.append(-2 * ml["ML"] + (X.shape[1] + 1) * math.log(n) + 2 * lchoose(L, X.shape[1] - 1))
        if int(np.flatnonzero(np.asarray(extBIC) == min(extBIC))[0]) == len(extBIC) - 1:
            W, v = scan_inputs_Z(K, Z, X, y, vc["ve"], vc["vg"])
            a = M.T @ v
            vara = np.einsum("ij,ij->j", M, W @ M)
            picked.append(pick_locus(a, vara)[0])
            X = np.column_stack([X, Z @ M[:, picked[-1] - 1]])
        else:
            cont = False
        itnum += 1
        if itnum > maxit:
            cont = False
    sel = picked if itnum > maxit else picked[:-1] if picked else picked
    return dict(all_picked=picked, selected=sel, extBIC=extBIC, vc=vc)


# ----------------------------------------------------------------------------- n x n algebra
def calculateH(MMt, varE, varG):
    """calculateH.R:36"""
    return varE * np.eye(MMt.shape[0]) + varG * MMt


def calculateP(H, X):
    """calculateP.R:27-28"""
    Hinv = r_chol2inv_chol(H)
    return Hinv - Hinv @ X @ np.linalg.inv(X.T @ Hinv @ X) @ X.T @ Hinv


def is_positive_definite(A, tol=1e-8):
    """matrixcalc::is.positive.definite (third-party, un-vendored; published behaviour):
    exact symmetry required, eigenvalues with |ev| < tol set to 0, all must be > 0."""
    if not np.array_equal(A, A.T):
        raise ValueError("argument x is not a symmetric matrix")
    ev = np.linalg.eigvalsh(A)
    ev[np.abs(ev) < tol] = 0.0
    return bool(np.all(ev > 0))


def calculateMMt_sqrt_and_sqrtinv(MMt):
    """calculateMMt_sqrt_and_sqrtinv.R:15-31"""
    if not is_positive_definite(MMt):
        raise ValueError("M %*% t(M) is not positive definite")
    w, v = r_eigen_sym(MMt)
    sq = v @ np.diag(np.sqrt(w)) @ v.T
    return sq, r_chol2inv_chol(sq)


def calculate_reduced_a(varG, P, MMtsqrt, y):
    """calculate_reduced_a.R:31  --  `varG * MMtsqrt %*% P %*% y` is left-associative."""
    return ((varG * MMtsqrt) @ P) @ y


def calculate_reduced_vara(X, varE, varG, invMMt, MMtsqrt):
    """calculate_reduced_vara.R:21-35"""
    n = invMMt.shape[0]
    Ze = MMtsqrt
    R1 = np.linalg.inv(varE * np.eye(n))
    G1 = np.linalg.inv(varG * np.eye(n))
    A = X.T @ R1 @ X
    B = X.T @ R1 @ Ze
    Cc = Ze.T @ R1 @ X
    D = Ze.T @ R1 @ Ze + G1
    D1 = np.linalg.inv(D)
    return varG * np.eye(n) - (D1 + D1 @ Cc @ np.linalg.inv(A - B @ D1 @ Cc) @ B @ D1)


# ----------------------------------------------------------------------------- the hot-path callers
def calcMMt(backend, geno, availmemGb, ncpu, selected_loci):
    """calcMMt.R:5-13 + calculateMMt.R:24-27"""
    sel = np.asarray(selected_loci, dtype=np.float64)
    if not np.any(np.isnan(sel)):
        sel = sel - 1
    MMt = backend.calculateMMt_rcpp(geno["asciifileM"], availmemGb, ncpu, sel, geno["dim_of_ascii_M"], True, None)
    MMt = np.asarray(MMt)
    return MMt / MMt.max() + np.diag(np.full(MMt.shape[0], 0.95))


def scan_inputs(MMt, invMMt, X, y, ve, vg, algebra=None):
    """find_qtl.R:5-45: everything the scan export consumes (S = K^-1/2, V, a_hat).
    `algebra`: an object exposing the five R-named functions (eagleeverything_b200.api, so that a test can drive
    the device implementation through the same loop); default: the restatements in this module."""
    if algebra is not None:
        H = algebra.calculateH(MMt, ve, vg)
        P = algebra.calculateP(H, X)
        r = algebra.calculateMMt_sqrt_and_sqrtinv(MMt)
        sq, sqinv = r["sqrt_MMt"], r["inverse_sqrt_MMt"]
        hat_a = np.asarray(algebra.calculate_reduced_a(vg, P, sq, y)).reshape(-1)
        return sqinv, algebra.calculate_reduced_vara(X, ve, vg, invMMt, sq), hat_a
    H = calculateH(MMt, ve, vg)
    P = calculateP(H, X)
    sq, sqinv = calculateMMt_sqrt_and_sqrtinv(MMt)
    hat_a = calculate_reduced_a(vg, P, sq, y)
    var_hat_a = calculate_reduced_vara(X, ve, vg, invMMt, sq)
    return sqinv, var_hat_a, hat_a


def pick_locus(a, vara):
    """find_qtl.R:71-83: tsq = a^2/vara; first index of max (NaN ignored); returns 1-based index."""
    with np.errstate(divide="ignore", invalid="ignore"):
        tsq = (np.asarray(a).reshape(-1) ** 2) / np.asarray(vara).reshape(-1)
    mx = np.nanmax(tsq)
    return int(np.flatnonzero(tsq == mx)[0]) + 1, tsq


def find_qtl(backend, geno, availmemGb, selected_loci, MMt, invMMt, ve, vg, X, y, trace=None, algebra=None):
    S, V, hat_a = scan_inputs(MMt, invMMt, X, y, ve, vg, algebra)
    n, L = geno["dim_of_ascii_M"]
    sel = np.asarray(selected_loci, dtype=np.float64)
    if not np.any(np.isnan(sel)):  # calculate_a_and_vara.R:23 (never true under AM(): AM.R:260)
        sel = sel - 1
    res = backend.calculate_a_and_vara_rcpp(geno["asciifileMt"], sel, S, V, availmemGb, (L, n), hat_a, True, None)
    idx, tsq = pick_locus(res["a"], res["vara"])
    if trace is not None:
        trace.append(dict(S=S, V=V, hat_a=hat_a, a=np.asarray(res["a"]).reshape(-1).copy(),
                          vara=np.asarray(res["vara"]).reshape(-1).copy(), tsq=tsq, picked=idx))
    return idx


def calc_extBIC(y, X, MMt, L):
    """calc_extBIC.R:6-9"""
    res_p = emma_MLE(y, X, MMt, llim=-100.0, ulim=100.0)
    BIC = -2 * res_p["ML"] + (X.shape[1] + 1) * math.log(len(y))
    return BIC + 2 * lchoose(L, X.shape[1] - 1)


def AM(backend, geno, y, X0=None, availmemGb=8, ncpu=1, maxit=20, keep_trace=False, algebra=None):
    """AM.R:260, 395-504.  geno = dict(asciifileM, asciifileMt, dim_of_ascii_M=(n, L));
    y = trait vector (no NAs); X0 = design matrix before marker effects (default: intercept).
    Returns dict(selected=[1-based loci], extBIC=[...], trace=[per-iteration scan inputs/outputs])."""
    y = np.asarray(y, dtype=np.float64)
    n, L = geno["dim_of_ascii_M"]
    X = np.ones((n, 1)) if X0 is None else np.asarray(X0, dtype=np.float64)
    selected_loci = [NA]      # AM.R:260
    new_selected_locus = NA   # AM.R:261
    extBIC = []
    trace = [] if keep_trace else None
    itnum = 1
    cont = True
    MMt = invMMt = None
    while cont:
        if not math.isnan(new_selected_locus):  # constructX.R:11-20
            col = backend.extract_geno_rcpp(geno["asciifileM"], availmemGb, int(new_selected_locus) - 1, (n, L))
            X = np.column_stack([X, np.asarray(col, dtype=np.float64)])
        if itnum == 1:  # AM.R:414-423
            MMt = calcMMt(backend, geno, availmemGb, ncpu, selected_loci)
            invMMt = r_chol2inv_chol(MMt)
        vc = emma_REMLE(y, X, MMt)                      # AM.R:428
        extBIC.append(calc_extBIC(y, X, MMt, L))         # AM.R:436
        if int(np.flatnonzero(np.asarray(extBIC) == min(extBIC))[0]) == len(extBIC) - 1:  # AM.R:448
            new_selected_locus = find_qtl(backend, geno, availmemGb, selected_loci, MMt, invMMt,
                                          vc["ve"], vc["vg"], X, y, trace, algebra)
            selected_loci.append(new_selected_locus)     # AM.R:455
        else:
            cont = False
        itnum += 1
        if itnum > maxit:                                # AM.R:465
            cont = False
    if itnum > maxit:                                    # AM.R:477-481
        final = selected_loci
    elif len(selected_loci) > 1:                         # AM.R:485-492
        final = selected_loci[:-1]
    else:
        final = selected_loci
    return dict(selected=[int(s) for s in final if not math.isnan(s)],
                all_picked=[int(s) for s in selected_loci if not math.isnan(s)],
                extBIC=extBIC, trace=trace, vc=vc)
