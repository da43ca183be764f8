"""AM(): the multi-locus forward search that calls the hot path (reference: R/AM.R:260, 395-504).

R is not installed in this image, so the loop an R user runs -- AM() -> emma.REMLE -> calc_extBIC -> find_qtl ->
calculate_a_and_vara -> which.max -- is mirrored here on top of this package's own mirrors of the Rcpp exports and of
the R algebra functions (api.py), i.e. on top of libeaglegpu's C ABI: every n x n operation (the three
eigendecompositions of EMMA, K^1/2 and K^-1/2, H, P, the reduced BLUP and its variance matrix), M.Mt, the genotype
column extraction and the a / var(a) scan run on the device; what stays on the host is what stays in R when the library
is dropped into the package -- the loop itself, EMMA's one-dimensional likelihood search over delta on n-vectors
(R/emma_REMLE.R:44-76, R/emma_MLE.R:34-56, uniroot) and the extended BIC (R/calc_extBIC.R:6-9).

Two things the R code recomputes in every iteration although their input never changes after iteration 1
(K = MMt/max(MMt) + 0.95 I is fixed: R/AM.R:414-423) are computed once here: eigen(K) inside emma.MLE and
K^1/2, K^-1/2 inside find_qtl.  Same inputs, same outputs; nothing else is reordered.

Single trait, no Z matrix (Z never reaches find_qtl in the reference snapshot: R/AM.R:450-452, R/find_qtl.R:1-2).
bench.py --search times this loop at BASELINE config 3; tests/test_am.py checks it against the golden demo results.
"""
from __future__ import annotations

import math
import time

import numpy as np

from . import api

NA = api.NA_REAL


# ----------------------------------------------------------------------------- genotype back ends
class FileGeno:
    """The R route: M.ascii / Mt.ascii on disk, the three Rcpp exports called with the reference's arguments."""

    def __init__(self, asciifileM, asciifileMt, dim_of_ascii_M, availmemGb=8.0, ncpu=1):
        self.M, self.Mt = asciifileM, asciifileMt
        self.n, self.L = int(dim_of_ascii_M[0]), int(dim_of_ascii_M[1])
        self.mem, self.ncpu = availmemGb, ncpu

    def mmt(self, selected0):
        return api.calculateMMt_rcpp(self.M, self.mem, self.ncpu, selected0, (self.n, self.L), True, None)

    def a_and_vara(self, selected0, S, V, a):
        r = api.calculate_a_and_vara_rcpp(self.Mt, selected0, S, V, self.mem, (self.L, self.n), a, True, None)
        return r["a"].reshape(-1), r["vara"].reshape(-1)

    def extract(self, locus0):
        return api.extract_geno_rcpp(self.M, self.mem, locus0, (self.n, self.L))


class ResidentGeno:
    """Stores already resident in HBM (api.GenotypeStore handles for M and Mt): no files involved."""

    def __init__(self, M_store, Mt_store):
        i = M_store.info()
        self.Ms, self.Mts, self.n, self.L = M_store, Mt_store, i["rows"], i["cols"]

    @staticmethod
    def _idx(selected0):
        s = np.atleast_1d(np.asarray(selected0, dtype=np.float64))
        return [] if np.isnan(s[0]) else [int(v) for v in s]

    def mmt(self, selected0):
        return self.Ms.mmt(self._idx(selected0))

    def a_and_vara(self, selected0, S, V, a):
        return self.Mts.a_and_vara(S, V, a, self._idx(selected0))

    def extract(self, locus0):
        return self.Ms.extract_col(locus0)


# ----------------------------------------------------------------------------- base-R pieces that stay on the host
def _uniroot(f, lower, upper, tol=np.finfo(float).eps ** 0.25, maxiter=1000):
    """uniroot() = R_zeroin2: Brent / Dekker 'zeroin' (Forsythe, Malcolm & Moler; netlib zeroin.c), R's default tol."""
    a, b = float(lower), float(upper)
    fa, fb = f(a), f(b)
    c, fc = a, fa
    eps = np.finfo(float).eps
    if fa == 0.0:
        return a
    if fb == 0.0:
        return b
    for _ in range(maxiter + 1):
        prev = b - a
        if abs(fc) < abs(fb):
            a, b, c = b, c, b
            fa, fb, fc = fb, fc, fb
        tol_act = 2 * eps * abs(b) + tol / 2
        step = (c - b) / 2
        if abs(step) <= tol_act or fb == 0.0:
            return b
        if abs(prev) >= tol_act and abs(fa) > abs(fb):
            cb = c - b
            if a == c:
                t1 = fb / fa
                p, q = cb * t1, 1.0 - t1
            else:
                q0, t1, t2 = fa / fc, fb / fc, fb / fa
                p = t2 * (cb * q0 * (q0 - t1) - (b - a) * (t1 - 1.0))
                q = (q0 - 1.0) * (t1 - 1.0) * (t2 - 1.0)
            if p > 0:
                q = -q
            else:
                p = -p
            if p < 0.75 * cb * q - abs(tol_act * q) / 2 and p < abs(prev * q / 2):
                step = p / q
        if abs(step) < tol_act:
            step = tol_act if step > 0 else -tol_act
        a, fa = b, fb
        b += step
        fb = f(b)
        if (fb > 0 and fc > 0) or (fb < 0 and fc < 0):
            c, fc = a, fa
    return b


def _lchoose(n, k):
    return math.lgamma(n + 1) - math.lgamma(k + 1) - math.lgamma(n - k + 1)


def _delta_search(dLL, logdelta, llim, ulim, esp, LL, dLLf):
    """R/emma_REMLE.R:54-76 and R/emma_MLE.R:34-56: boundary candidates, then a root of dLL in every grid cell where
    it changes sign from + to -; the first maximum of the likelihood among the candidates."""
    m = len(logdelta)
    cand, ll = [], []
    if dLL[0] < esp:
        cand.append(llim)
        ll.append(LL(llim))
    if dLL[m - 2] > 0 - esp:
        cand.append(ulim)
        ll.append(LL(ulim))
    for i in range(m - 1):
        if dLL[i] * dLL[i + 1] < 0 - esp * esp and dLL[i] > 0 and dLL[i + 1] < 0:
            r = _uniroot(dLLf, logdelta[i], logdelta[i + 1])
            cand.append(r)
            ll.append(LL(r))
    k = int(np.argmax(ll))
    return math.exp(cand[k]), ll[k]


class _Emma:
    """emma.REMLE / emma.MLE without Z (R/emma_REMLE.R:27-131, R/emma_MLE.R:2-117): the eigendecompositions on the
    device (api.emma_eigen_R_wo_Z / emma_eigen_L_wo_Z), the search over delta here."""

    def __init__(self, K, stats):
        self.K, self.stats = K, stats
        self._xi = None      # eigen(K): K is fixed after iteration 1
        self._last = None    # (X id, lam, etasq): REMLE and MLE of one iteration share eigen(S(K+I)S)

    def _eig_R(self, y, X):
        key = X.shape[1]
        if self._last is None or self._last[0] != key:
            t0 = time.perf_counter()
            r = api.emma_eigen_R_wo_Z(self.K, X)
            etas = r["vectors"].T @ y
            self._last = (key, r["values"], etas * etas)
            self.stats["emma_eigen_s"] += time.perf_counter() - t0
        return self._last[1], self._last[2]

    def _xi_K(self):
        if self._xi is None:
            t0 = time.perf_counter()
            self._xi = api.emma_eigen_L_wo_Z(self.K, vectors=False)["values"]
            self.stats["emma_eigen_s"] += time.perf_counter() - t0
        return self._xi

    @staticmethod
    def _grid(ngrids, llim, ulim):
        logdelta = np.arange(ngrids + 1) / ngrids * (ulim - llim) + llim
        return logdelta, np.exp(logdelta)

    def REMLE(self, y, X, ngrids=100, llim=-10.0, ulim=10.0, esp=1e-10):
        n, q = len(y), X.shape[1]
        if np.linalg.det(X.T @ X) == 0:
            return dict(REML=0.0, delta=0.0, ve=0.0, vg=0.0)
        lam, etasq = self._eig_R(y, X)
        logdelta, delta = self._grid(ngrids, llim, ulim)
        Lam = lam[:, None] + delta[None, :]
        E = etasq[:, None]
        dLL = 0.5 * delta * ((n - q) * (E / (Lam * Lam)).sum(0) / (E / Lam).sum(0) - (1.0 / Lam).sum(0))
        nq = len(etasq)

        def LL(ld):
            d = math.exp(ld)
            return 0.5 * (nq * (math.log(nq / (2 * math.pi)) - 1 - math.log((etasq / (lam + d)).sum())) - np.log(lam + d).sum())

        def dLLf(ld):
            ldel = lam + math.exp(ld)
            return 0.5 * (nq * (etasq / (ldel * ldel)).sum() / (etasq / ldel).sum() - (1.0 / ldel).sum())

        maxdelta, maxLL = _delta_search(dLL, logdelta, llim, ulim, esp, LL, dLLf)
        maxva = (etasq / (lam + maxdelta)).sum() / (n - q)
        return dict(REML=maxLL, delta=maxdelta, ve=maxva * maxdelta, vg=maxva)

    def MLE(self, y, X, ngrids=100, llim=-10.0, ulim=10.0, esp=1e-10):
        n = len(y)
        if np.linalg.det(X.T @ X) == 0:
            return dict(ML=0.0, delta=0.0, ve=0.0, vg=0.0)
        xi = self._xi_K()
        lam, etasq = self._eig_R(y, X)
        logdelta, delta = self._grid(ngrids, llim, ulim)
        Lam = lam[:, None] + delta[None, :]
        Xis = xi[:, None] + delta[None, :]
        E = etasq[:, None]
        dLL = 0.5 * delta * (n * (E / (Lam * Lam)).sum(0) / (E / Lam).sum(0) - (1.0 / Xis).sum(0))
        nn = len(xi)

        def LL(ld):
            d = math.exp(ld)
            return 0.5 * (nn * (math.log(nn / (2 * math.pi)) - 1 - math.log((etasq / (lam + d)).sum())) - np.log(xi + d).sum())

        def dLLf(ld):
            d = math.exp(ld)
            ldel = lam + d
            return 0.5 * (nn * (etasq / (ldel * ldel)).sum() / (etasq / ldel).sum() - (1.0 / (xi + d)).sum())

        maxdelta, maxLL = _delta_search(dLL, logdelta, llim, ulim, esp, LL, dLLf)
        maxva = (etasq / (lam + maxdelta)).sum() / n
        return dict(ML=maxLL, delta=maxdelta, ve=maxva * maxdelta, vg=maxva)


def emma_w_Z(n, t, q, lam, e1sq, e2sq, xi=None, ngrids=100, llim=-10.0, ulim=10.0, esp=1e-10):
    """The Z branches of emma.REMLE (R/emma_REMLE.R:78-131; xi is None) and emma.MLE (R/emma_MLE.R:57-117; xi =
    eigen(K Z'Z)$values) on what the eigen step delivers: lam = the t - q eigenvalues emma.eigen.R.w.Z keeps
    (R/emma_eigen_R_w_Z.R:13-21), e1sq = the squares of the matching etas, e2sq = the sum of squares of the n - t other
    etas.  n records, t individuals, q fixed effects."""
    logdelta, delta = _Emma._grid(ngrids, llim, ulim)
    Lam = lam[:, None] + delta[None, :]
    E = e1sq[:, None]
    ml = xi is not None
    if ml:
        Xis = xi[:, None] + delta[None, :]
        nt = n - len(xi)
        dLL = 0.5 * delta * (n * ((E / (Lam * Lam)).sum(0) + e2sq / (delta * delta)) / ((E / Lam).sum(0) + e2sq / delta)
                             - ((1.0 / Xis).sum(0) + nt / delta))

        def LL(ld):
            d = math.exp(ld)
            return 0.5 * (n * (math.log(n / (2 * math.pi)) - 1 - math.log((e1sq / (lam + d)).sum() + e2sq / d))
                          - (np.log(xi + d).sum() + nt * ld))

        def dLLf(ld):
            d = math.exp(ld)
            ldel = lam + d
            return 0.5 * (n * ((e1sq / (ldel * ldel)).sum() + e2sq / (d * d)) / ((e1sq / ldel).sum() + e2sq / d)
                          - ((1.0 / (xi + d)).sum() + nt / d))
        denom = n
    else:
        nq = n - t + len(lam)
        dLL = 0.5 * delta * ((n - q) * ((E / (Lam * Lam)).sum(0) + e2sq / (delta * delta)) / ((E / Lam).sum(0) + e2sq / delta)
                             - ((1.0 / Lam).sum(0) + (n - t) / delta))

        def LL(ld):
            d = math.exp(ld)
            return 0.5 * (nq * (math.log(nq / (2 * math.pi)) - 1 - math.log((e1sq / (lam + d)).sum() + e2sq / d))
                          - (np.log(lam + d).sum() + (n - t) * ld))

        def dLLf(ld):
            d = math.exp(ld)
            ldel = lam + d
            return 0.5 * (nq * ((e1sq / (ldel * ldel)).sum() + e2sq / (d * d)) / ((e1sq / ldel).sum() + e2sq / d)
                          - ((1.0 / ldel).sum() + (n - t) / d))
        denom = n - q
    maxdelta, maxLL = _delta_search(dLL, logdelta, llim, ulim, esp, LL, dLLf)
    maxva = ((e1sq / (lam + maxdelta)).sum() + e2sq / maxdelta) / denom
    return {("ML" if ml else "REML"): maxLL, "delta": maxdelta, "ve": maxva * maxdelta, "vg": maxva}


def z_index(Z, t):
    """Z (n records x t individuals, 0/1 with exactly one 1 per row: R/ReadZmat.R:47-85) as the index of each record's
    individual; an index vector passes through.  Every individual needs at least one record here (emma.MLE drops the
    others, emma.REMLE does not: R/emma_MLE.R:60-63 pass `ngpu` in the position of `complete`)."""
    Z = np.asarray(Z)
    if Z.ndim == 2:
        if Z.shape[1] != t or not np.all((Z == 0) | (Z == 1)) or not np.all(Z.sum(1) == 1):
            raise ValueError("Z must be a 0/1 matrix with one 1 per row and one column per individual")
        idx = Z.argmax(1).astype(np.int64)
    else:
        idx = Z.astype(np.int64)
    if idx.min() < 0 or idx.max() >= t or len(np.unique(idx)) != t:
        raise ValueError("every individual needs at least one record in Z")
    return idx


# ----------------------------------------------------------------------------- the loop
def AM(geno, y, X0=None, maxit=20, message=None):
    """R/AM.R:260, 395-504.  geno: FileGeno or ResidentGeno; y: trait (no NAs); X0: fixed-effects design matrix before
    any marker (default: the intercept).  -> dict(selected = 1-based loci of the final model, all_picked, extBIC,
    vc = last variance components, seconds = where the time went)."""
    say = message or (lambda s: None)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n, L = geno.n, geno.L
    X = np.ones((n, 1)) if X0 is None else np.asarray(X0, dtype=np.float64).reshape(n, -1)
    stats = dict(mmt_s=0.0, emma_eigen_s=0.0, emma_search_s=0.0, algebra_s=0.0, scan_s=0.0, extract_s=0.0, total_s=0.0)
    t_all = time.perf_counter()
    selected = [NA]            # AM.R:260
    new_locus = NA             # AM.R:261
    extBIC = []
    itnum, cont = 1, True
    K = emma = roots = vc = None
    while cont:
        if not math.isnan(new_locus):    # constructX.R:11-20: the picked marker's genotypes join the fixed effects
            t0 = time.perf_counter()
            X = np.column_stack([X, np.asarray(geno.extract(int(new_locus) - 1), dtype=np.float64)])
            stats["extract_s"] += time.perf_counter() - t0
        if itnum == 1:                   # AM.R:414-423, calcMMt.R:5-13: K = MMt / max(MMt) + 0.95 I
            t0 = time.perf_counter()
            MMt = np.asarray(geno.mmt(np.asarray(selected)))
            K = MMt / MMt.max() + np.diag(np.full(n, 0.95))
            del MMt
            stats["mmt_s"] += time.perf_counter() - t0
            emma = _Emma(K, stats)
        t0 = time.perf_counter()
        e0 = stats["emma_eigen_s"]
        vc = emma.REMLE(y, X)                                                   # AM.R:428
        ml = emma.MLE(y, X, llim=-100.0, ulim=100.0)                             # calc_extBIC.R:6
        stats["emma_search_s"] += time.perf_counter() - t0 - (stats["emma_eigen_s"] - e0)
        bic = -2 * ml["ML"] + (X.shape[1] + 1) * math.log(n)                     # calc_extBIC.R:7-9
        extBIC.append(bic + 2 * _lchoose(L, X.shape[1] - 1))
        say(f" iteration {itnum}: extBIC = {extBIC[-1]:.4f}")
        if int(np.flatnonzero(np.asarray(extBIC) == min(extBIC))[0]) == len(extBIC) - 1:   # AM.R:448
            # find_qtl.R:5-83
            t0 = time.perf_counter()
            if roots is None:            # calculateMMt_sqrt_and_sqrtinv.R: K is the same in every iteration
                roots = api.calculateMMt_sqrt_and_sqrtinv(K, message=message)
                if roots is None:
                    raise ValueError("M %*% t(M) is not positive definite")
            H = api.calculateH(K, vc["ve"], vc["vg"], message=message)
            P = api.calculateP(H, X)
            del H
            hat_a = api.calculate_reduced_a(vc["vg"], P, roots["sqrt_MMt"], y).reshape(-1)
            del P
            V = api.calculate_reduced_vara(X, vc["ve"], vc["vg"], K, roots["sqrt_MMt"])
            stats["algebra_s"] += time.perf_counter() - t0
            t0 = time.perf_counter()
            sel = np.asarray(selected, dtype=np.float64)
            if not np.any(np.isnan(sel)):                                        # calculate_a_and_vara.R:23
                sel = sel - 1
            a, vara = geno.a_and_vara(sel, roots["inverse_sqrt_MMt"], V, hat_a)
            del V
            with np.errstate(divide="ignore", invalid="ignore"):
                tsq = a * a / vara                                               # find_qtl.R:71
            new_locus = int(np.flatnonzero(tsq == np.nanmax(tsq))[0]) + 1         # :76-80 first maximum, NaN ignored
            stats["scan_s"] += time.perf_counter() - t0
            selected.append(new_locus)                                           # AM.R:455
            say(f" iteration {itnum}: picked locus {new_locus}")
        else:
            cont = False
        itnum += 1
        if itnum > maxit:                                                        # AM.R:465
            cont = False
    if itnum > maxit:                                                            # AM.R:477-481
        final = selected
    elif len(selected) > 1:                                                      # AM.R:485-492
        final = selected[:-1]
    else:
        final = selected
    stats["total_s"] = time.perf_counter() - t_all
    return dict(selected=[int(s) for s in final if not math.isnan(s)],
                all_picked=[int(s) for s in selected if not math.isnan(s)], extBIC=extBIC, vc=vc,
                iterations=itnum - 1, seconds={k: round(v, 4) for k, v in stats.items()})


# ----------------------------------------------------------------------------- everything resident, in the basis of eigen(K)
def eigbasis_inputs(xi, Xt, yt, ve, vg, XtX=None, Xty=None):
    """The inputs of the scan from eigenbasis quantities (host, O(n q^2)): with K = U diag(xi) U^T, Xt = U^T X, yt = U^T y,
         Dh = 1 / (ve + vg xi)                                  H^-1 = U diag(Dh) U^T                  (R/calculateH.R:36)
         C  = Xt^T Dh Xt = L L^T                                t(X) Hinv X                            (R/calculateP.R:28)
         W  = K^-1/2 V K^-1/2 = vg^2 P = U diag(vg^2 Dh) U^T - E E^T,   E = U Et,  Et = vg Dh Xt L^-T
                                              (R/calculate_reduced_vara.R:21-35, src/calculate_a_and_vara_rcpp.cpp:97-98)
         v  = K^-1/2 a_hat = vg P y = U vt,   vt = vg (Dh yt - Dh Xt C^-1 Xt^T Dh yt)
                                              (R/calculate_reduced_a.R:31, src/calculate_a_and_vara_rcpp.cpp:90)
    -> (w, Et, vt).  tests/test_gpu_algebra.py and tests/test_secular_cpu.py check W and v against R's dense formulas."""
    Dh = 1.0 / (ve + vg * xi)
    B = Dh[:, None] * Xt
    Cq = Xt.T @ B
    rhs = B.T @ yt
    if XtX is not None:
        # repeated measures (Z): H = ve I + vg Z K Z' has the eigenvalues ve + vg xi on the t directions Z C^-1/2 h_j
        # (xi, h_j: eigenpairs of C^1/2 K C^1/2, C = Z'Z) and ve on the other n - t; Xt, yt are the coordinates of X, y
        # on the first, XtX - Xt'Xt and Xty - Xt'yt what they carry on the others.  The scan then runs on
        # W = vg^2 Z' P Z = A diag(w) A' - E E' and v = vg Z' P y = A vt with A = C^1/2 [h_1 .. h_t] in place of U.
        Cq = Cq + (XtX - Xt.T @ Xt) / ve
        rhs = rhs + (Xty - Xt.T @ yt) / ve
    Lc = np.linalg.cholesky(Cq)
    Et = vg * np.linalg.solve(Lc, B.T).T
    vt = vg * (Dh * yt - B @ np.linalg.solve(Cq, rhs))
    return vg * vg * Dh, np.asfortranarray(Et), vt


def AM_resident(store_kb, storeT, n, L, y, X0=None, maxit=20, message=None, shard=None, Z=None, bcache="auto"):
    """See _AM_resident.  The host side of the loop only touches n-vectors and n x q panels: BLAS worker threads would
    gain nothing there, and their spin-waiting after every small product delays the thread that launches the kernels
    (measured at config 3: 0.3 - 0.8 s of the search with 16 OpenBLAS threads, 0.07 s with one)."""
    try:
        from threadpoolctl import threadpool_limits
    except ImportError:
        return _AM_resident(store_kb, storeT, n, L, y, X0, maxit, message, shard, Z, bcache)
    with threadpool_limits(limits=1):
        return _AM_resident(store_kb, storeT, n, L, y, X0, maxit, message, shard, Z, bcache)


def _AM_resident(store_kb, storeT, n, L, y, X0=None, maxit=20, message=None, shard=None, Z=None, bcache="auto"):
    """AM()'s forward search (R/AM.R:260, 395-504) with the genotypes resident in HBM and the n x n algebra of every
    iteration carried out in the basis of eigen(K) (csrc/eigbasis.cu): K = MMt/max(MMt) + 0.95 I never changes after the
    first iteration (R/AM.R:414-423), so it is decomposed ONCE; after that an iteration costs
      * EMMA's eigen(S (K + I) S) (R/emma_eigen_R_wo_Z.R:7-20): a secular solve, O(q n^2), no n x n matrix;
      * H, P, K^+-1/2, a_hat, V (R/find_qtl.R:5-43): n-vectors and n x q panels on the host (eigbasis_inputs);
      * the scan's right-hand side W: ONE n^3 product on the int8 tensor cores (eg_dev_scan_prepare_eig);
      * the a / var(a) scan and the pick (src/calculate_a_and_vara_rcpp.cpp, R/find_qtl.R:71-83).
    Same picks and extBIC trace as the dense route (AM_resident_dense / AM over the host ABI; tests/test_am.py).

    store_kb: K-blocked int8 M store of THIS rank's markers (device.decode_kb), storeT: the row-major Mt store of the same
    markers (device.transpose_kb); L: the number of markers of the whole data set.  shard (dist.Shard or None): marker
    shards over torch.distributed ranks -- partial M.Mt all-reduced (int32), scans sharded, the pick by the sharded
    first-maximum rule, the picked column broadcast by its owner; the n x n algebra is replicated.

    Z (repeated measures; BASELINE config 5): the records' incidence matrix (or each record's individual as an index
    vector).  y and X0 then have one row per RECORD, n stays the number of individuals.  EMMA's Z branches
    (R/emma_eigen_L_w_Z.R:2-14, R/emma_eigen_R_w_Z.R:2-23, R/emma_REMLE.R:78-131, R/emma_MLE.R:57-117) run on the
    eigenpairs of C^1/2 K C^1/2 (C = Z'Z, computed once) through the same secular solve, and the scan is fed the Z-aware
    H = ve I + vg Z K Z' (in the reference snapshot Z stops at EMMA and a non-square Z cannot
    pass find_qtl: R/AM.R:450-452, R/calculateP.R:22-25).

    bcache ("auto" / True / False): K -- hence U -- is the same in every iteration, so the projection B = M^T U of this
    rank's markers (L_local x n doubles: 80 GB at config 3) can be computed ONCE, by the scan's own int8 digit-slice
    contraction in projection mode (eg_dev_project_i8); every later scan is then
    var(a)_j = sum_k w_k B_jk^2 - sum_c (E_c^T m_j)^2, a_j = m_j^T v: one HBM-bound pass over B plus q + 1 exact int8
    matrix-vector products, instead of the n^2 L contraction of src/calculate_a_and_vara_rcpp.cpp:103-112 per iteration
    (and no n^3 product for W at all).  "auto": when B fits in the free device memory with 12 GB to spare and at least three
    iterations are allowed (the projection costs as much as two scans)."""
    import ctypes as C

    import torch

    from . import _lib, device
    lib = _lib.require_gpu()
    say = message or (lambda s: None)
    dev = store_kb.device
    f64 = dict(dtype=torch.float64, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())                                      # noqa: E731
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)             # noqa: E731
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))                       # noqa: E731
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    zidx = None if Z is None else z_index(Z, n)
    nrec = n if zidx is None else len(zidx)
    if len(y) != nrec:
        raise ValueError("y needs one value per record")
    X = np.ones((nrec, 1)) if X0 is None else np.asarray(X0, dtype=np.float64).reshape(nrec, -1)
    cnt = None if zidx is None else np.bincount(zidx, minlength=n).astype(np.float64)
    Lloc = storeT.shape[0]
    stats = dict(mmt_s=0.0, eigen_K_s=0.0, dsyevd_s=0.0, project_s=0.0, emma_eigen_s=0.0, emma_search_s=0.0, algebra_s=0.0, scan_s=0.0,
                 extract_s=0.0, total_s=0.0)

    def timed(key, t0):
        torch.cuda.synchronize()
        stats[key] += time.perf_counter() - t0

    t_all = time.perf_counter()
    kern = {}
    # ---- M M^T (AM.R:414-417), K (calcMMt.R:13), eigen(K) once
    t0 = time.perf_counter()
    U = torch.empty((n, n), **f64)           # K, then its eigenvectors (columns; column-major)
    C32 = device.syrk_kb(store_kb, n, Lloc)
    if shard is not None:
        shard.allreduce_mmt(C32)
    device.mmt_finalize(C32, n, out=U)
    del C32
    U.div_(U.max())
    U.diagonal().add_(0.95)
    timed("mmt_s", t0)
    t0 = time.perf_counter()
    vals = torch.empty(n, **f64)
    if zidx is not None:                                                         # C^1/2 K C^1/2 (same spectrum as K Z'Z)
        sc_d = torch.from_numpy(np.sqrt(cnt)).to(dev)
        U.mul_(sc_d[:, None]).mul_(sc_d[None, :])
    _lib.check(lib.eg_dev_eigen_sym(p(U), n, p(vals), st()))                     # values decreasing, as R's eigen()
    torch.cuda.synchronize()
    stats["dsyevd_s"] = round(time.perf_counter() - t0, 4)                       # the one library factorisation of the search
    xi = vals.cpu().numpy().copy()
    if not np.all(np.where(np.abs(xi) < 1e-8, 0.0, xi) > 0):                     # matrixcalc::is.positive.definite's screen
        raise ValueError("M %*% t(M) is not positive definite")                  # calculateMMt_sqrt_and_sqrtinv.R:15-23
    if zidx is not None:
        U.mul_(sc_d[None, :])                                                    # A = C^1/2 [h_1 .. h_t] (rows of the torch view = columns)
    ldb = (n + 1) // 2 * 2
    if bcache == "auto":
        free_b, _tot = torch.cuda.mem_get_info()
        use_b = maxit >= 3 and Lloc * ldb * 8 + (12 << 30) < free_b   # the projection costs two scans: pays from the third
    else:
        use_b = bool(bcache)
    if use_b:
        Ut = Wp = work2 = None
    else:
        Ut = torch.empty((n, n), **f64)
        _lib.check(lib.eg_dev_transpose_f64(p(U), n, p(Ut), st()))
        Wp = torch.empty(lib.eg_scan_wp_elems(n), **f64)
        work2 = torch.empty((n, n), **f64) if not lib.eg_prep_uses_i8(n) else None

    def apply_U(v, to_eigenbasis):
        d_in = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(dev)
        d_out = torch.empty(n, **f64)
        _lib.check(lib.eg_dev_eigbasis_apply(p(U), n, p(d_in), 1, 1 if to_eigenbasis else 0, p(d_out), st()))
        return d_out.cpu().numpy()

    def to_eig(v):
        """Coordinates of a vector on the eigen directions.  No Z: U^T v.  With Z (v has one entry per record): on the t
        directions Z C^-1/2 h_j, i.e. h_j^T C^-1/2 Z^T v = A^T C^-1 Z^T v; returns also the part of v outside their span."""
        if zidx is None:
            return apply_U(v, True)
        top = apply_U(np.bincount(zidx, weights=v, minlength=n) / cnt, True)
        return top, v - (apply_U(top, False) / cnt)[zidx]

    if zidx is None:
        yt = to_eig(y)
        Xt = np.column_stack([to_eig(X[:, c]) for c in range(X.shape[1])])
        Res = None
    else:
        yt, ry = to_eig(y)
        tops = [to_eig(X[:, c]) for c in range(X.shape[1])]
        Xt = np.column_stack([a for a, _ in tops])
        Res = [r for _, r in tops]                                               # residuals of the columns of X (and of y: ry)
    timed("eigen_K_s", t0)
    if use_b:
        t0 = time.perf_counter()
        Bc = torch.empty((Lloc, ldb), **f64)
        ev_p = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev_p[0].record()
        _lib.check(lib.eg_dev_project_i8(p(storeT), Lloc, n, storeT.stride(0), p(U), p(Bc), ldb, st()))
        ev_p[1].record()
        tmpL = torch.empty(Lloc, **f64)
        timed("project_s", t0)
        kern["project_ms"] = ev_p[0].elapsed_time(ev_p[1])
        kern["project_int8_tops"] = 2.0 * 7 * Lloc * n * n / (kern["project_ms"] * 1e-3) / 1e12   # 7 digit slices of the full U
    emma = _Emma(None, stats)
    emma._xi = xi
    bscan_ms, gemv_ms = [], []
    selected, new_locus, extBIC = [NA], NA, []
    itnum, cont, vc = 1, True, None
    sec_stats = []
    while cont:
        if not math.isnan(new_locus):                                            # constructX.R:11-20
            t0 = time.perf_counter()
            g = int(new_locus) - 1
            if shard is None:
                col = device.extract_col(store_kb, n, g, kblocked=True)
            else:
                col = shard.fetch_col(lambda j: device.extract_col(store_kb, n, j, kblocked=True), n, g, dev)
            xnew = col.cpu().numpy().astype(np.float64)
            if zidx is None:
                X = np.column_stack([X, xnew])
                Xt = np.column_stack([Xt, to_eig(xnew)])
            else:
                xnew = xnew[zidx]                                                # Z m: the marker's genotype of every record
                X = np.column_stack([X, xnew])
                top, r = to_eig(xnew)
                Xt = np.column_stack([Xt, top])
                Res.append(r)
            timed("extract_s", t0)
        q = X.shape[1]
        # ---- emma.REMLE and emma.MLE share eigen(S (K + I) S) of this iteration
        t0 = time.perf_counter()
        s4 = (C.c_int64 * 4)()
        if zidx is None:
            lam, et = np.empty(n - q), np.empty(n - q)
            Xtf = np.asfortranarray(Xt)
            _lib.check(lib.eg_emma_eigen_R_wo_Z_eigbasis(dp(xi), dp(Xtf), dp(yt), n, q, dp(lam), dp(et), s4))
            emma._last = (q, lam, et * et)
        else:
            # the n - t directions outside range(Z) carry the eigenvalue 0; X and y only reach the q + 1 of them that
            # their own residuals span: t + q + 1 poles describe the whole compression (R/emma_eigen_R_w_Z.R:10-21)
            Qr, Rr = np.linalg.qr(np.column_stack(Res + [ry]))
            m = n + q + 1
            xi_x = np.concatenate([xi, np.zeros(q + 1)])
            Xt_x = np.asfortranarray(np.vstack([Xt, Rr[:, :q]]))
            yt_x = np.concatenate([yt, Rr[:, q]])
            val_x, et_x = np.empty(m - q), np.empty(m - q)
            _lib.check(lib.eg_emma_eigen_R_wo_Z_eigbasis(dp(xi_x), dp(Xt_x), dp(yt_x), m, q, dp(val_x), dp(et_x), s4))
            lam, e1sq, e2sq = val_x[: n - q], et_x[: n - q] ** 2, float((et_x[n - q:] ** 2).sum())
        t2 = (C.c_double * 2)()
        lib.eg_last_secular_times(t2)
        sec_stats.append(list(s4) + [t2[0], t2[1], time.perf_counter() - t0])
        timed("emma_eigen_s", t0)
        t0 = time.perf_counter()
        if zidx is None:
            vc = emma.REMLE(y, X)                                                 # AM.R:428
            ml = emma.MLE(y, X, llim=-100.0, ulim=100.0)                         # calc_extBIC.R:6
        elif np.linalg.det(X.T @ X) == 0:
            vc, ml = dict(REML=0.0, delta=0.0, ve=0.0, vg=0.0), dict(ML=0.0, delta=0.0, ve=0.0, vg=0.0)
        else:
            vc = emma_w_Z(nrec, n, q, lam, e1sq, e2sq)
            ml = emma_w_Z(nrec, n, q, lam, e1sq, e2sq, xi=xi, llim=-100.0, ulim=100.0)
        stats["emma_search_s"] += time.perf_counter() - t0
        bic = -2 * ml["ML"] + (q + 1) * math.log(nrec)
        extBIC.append(bic + 2 * _lchoose(L, q - 1))
        say(f" iteration {itnum}: extBIC = {extBIC[-1]:.4f}")
        if int(np.flatnonzero(np.asarray(extBIC) == min(extBIC))[0]) == len(extBIC) - 1:   # AM.R:448
            t0 = time.perf_counter()
            if zidx is None:
                w, Et, vt = eigbasis_inputs(xi, Xt, yt, float(vc["ve"]), float(vc["vg"]))
            else:
                w, Et, vt = eigbasis_inputs(xi, Xt, yt, float(vc["ve"]), float(vc["vg"]), XtX=X.T @ X, Xty=X.T @ y)
            d_w = torch.from_numpy(w).to(dev)
            d_Et = torch.from_numpy(Et.T.copy()).to(dev)                         # q x n row-major = n x q column-major
            d_vt = torch.from_numpy(vt).to(dev)
            if use_b:
                d_E = torch.empty((q, n), **f64)                                 # E = U Et, column c = row c of the torch view
                d_v = torch.empty(n, **f64)
                _lib.check(lib.eg_dev_eigbasis_apply(p(U), n, p(d_Et), q, 0, p(d_E), st()))
                _lib.check(lib.eg_dev_eigbasis_apply(p(U), n, p(d_vt), 1, 0, p(d_v), st()))
                timed("algebra_s", t0)
                t0 = time.perf_counter()
                e = torch.empty((q, Lloc), **f64)
                a, vara = torch.empty(Lloc, **f64), torch.empty(Lloc, **f64)
                ev_b = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev_b[0].record()
                for c in range(q):                                               # e_c = Mt E_c: exact int8 digit products (DP4A)
                    _lib.check(lib.eg_dev_gemv_i8(p(storeT), Lloc, n, storeT.stride(0), p(d_E[c]), 1.0, p(e[c]), st()))
                _lib.check(lib.eg_dev_gemv_i8(p(storeT), Lloc, n, storeT.stride(0), p(d_v), 1.0, p(a), st()))
                ev_b[1].record()
                _lib.check(lib.eg_dev_bscan(p(Bc), Lloc, n, ldb, p(d_w), p(e), q, p(tmpL), p(vara), st()))
                ev_b[2].record()
                ev_b[2].synchronize()
                gemv_ms.append(ev_b[0].elapsed_time(ev_b[1]))
                bscan_ms.append(ev_b[1].elapsed_time(ev_b[2]))
            else:
                work = torch.empty(n * q, **f64)
                _lib.check(lib.eg_dev_scan_prepare_eig(p(U), p(Ut), n, p(d_w), p(d_Et), q, p(d_vt), p(work),
                                                       p(work2) if work2 is not None else None, p(Wp), st()))
                timed("algebra_s", t0)
                t0 = time.perf_counter()
                a, vara = device.scan(storeT, Lloc, n, Wp)
            best, idx = device.argmax_tsq(a, vara)
            if shard is None:
                new_locus = int(idx.item()) + 1
            else:
                new_locus = int(shard.global_argmax(best, idx)[1]) + 1
            del a, vara
            timed("scan_s", t0)
            selected.append(new_locus)
            say(f" iteration {itnum}: picked locus {new_locus}")
        else:
            cont = False
        itnum += 1
        if itnum > maxit:
            cont = False
    if itnum > maxit:
        final = selected
    elif len(selected) > 1:
        final = selected[:-1]
    else:
        final = selected
    torch.cuda.synchronize()
    stats["total_s"] = time.perf_counter() - t_all
    return dict(selected=[int(s) for s in final if not math.isnan(s)],
                all_picked=[int(s) for s in selected if not math.isnan(s)], extBIC=extBIC, vc=vc,
                iterations=itnum - 1, seconds={k: round(v, 4) for k, v in stats.items()},
                scan_route="cached projection B = M^T U (one HBM-bound pass per iteration)" if use_b else "n^2 L contraction per iteration",
                scan_kernels=(dict(kern, bscan_ms=[round(x, 3) for x in bscan_ms], gemv_ms=[round(x, 3) for x in gemv_ms],
                                   bscan_gbs=(8.0 * Lloc * n / (min(bscan_ms) * 1e-3) / 1e9) if bscan_ms else None,
                                   bscan_bytes="8 n bytes per marker (B read once); gemv: (q + 1) passes over the int8 Mt store")
                              if use_b else None),
                secular={"max_root_iterations": max(s[2] for s in sec_stats), "deflated_poles": sum(s[1] for s in sec_stats),
                         "roots": sum(s[3] for s in sec_stats),
                         "seconds_per_call": [dict(device_back_end=round(s[4], 4), library_call=round(s[5], 4), with_python=round(s[6], 4))
                                              for s in sec_stats]})


# ----------------------------------------------------------------------------- everything resident in HBM
def AM_resident_dense(store_kb, storeT, n, L, y, X0=None, maxit=20, message=None):
    """(Round-1 form, kept as a cross-check of AM_resident: R's dense formulas executed one for one on the device --
    cuSOLVER dsyevd / potrf / potri and cuBLAS products every iteration.)
    The same search with every n x n matrix living in HBM from the first iteration to the last (the B200-first form:
    nothing but n-vectors, the q-column design matrix and scalars crosses PCIe after the genotypes are resident).
    store_kb: the K-blocked int8 M store (device.decode_kb), storeT: the row-major Mt store (device.transpose_kb);
    torch owns the buffers, the device-level C ABI (eg_dev_*) does the work, EMMA's 1-D search stays on the host.
    Same results as AM() up to the summation order of two matrix-vector products (tests/test_am.py)."""
    import ctypes as C

    import torch

    from . import _lib, device
    lib = _lib.require_gpu()
    say = message or (lambda s: None)
    dev = store_kb.device
    f64 = dict(dtype=torch.float64, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())                                      # noqa: E731
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)             # noqa: E731
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    X = np.ones((n, 1)) if X0 is None else np.asarray(X0, dtype=np.float64).reshape(n, -1)
    stats = dict(mmt_s=0.0, emma_eigen_s=0.0, emma_search_s=0.0, algebra_s=0.0, scan_s=0.0, extract_s=0.0, total_s=0.0)

    def timed(key, t0):
        torch.cuda.synchronize()
        stats[key] += time.perf_counter() - t0

    t_all = time.perf_counter()
    y_d = torch.from_numpy(y).to(dev)
    K = torch.empty((n, n), **f64)       # column-major n x n buffers; K, its roots and V are symmetric
    U = torch.empty((n, n), **f64)
    w1 = torch.empty((n, n), **f64)
    w2 = torch.empty((n, n), **f64)
    vals = torch.empty(n, **f64)
    etas = torch.empty(n, **f64)
    tmp_n = torch.empty(n, **f64)
    hat_a = torch.empty(n, **f64)
    sq = inv = None
    xi = None
    selected, new_locus, extBIC = [NA], NA, []
    itnum, cont, vc = 1, True, None
    emma = _Emma(None, stats)
    while cont:
        if not math.isnan(new_locus):
            t0 = time.perf_counter()
            col = device.extract_col(store_kb, n, int(new_locus) - 1, kblocked=True)
            X = np.column_stack([X, col.cpu().numpy().astype(np.float64)])
            timed("extract_s", t0)
        q = X.shape[1]
        X_d = torch.from_numpy(np.asfortranarray(X).T.copy()).to(dev)          # q x n row-major = n x q column-major
        small = torch.empty(4 * n * q + 3 * q * q + 16, **f64)
        if itnum == 1:
            t0 = time.perf_counter()
            C32 = device.syrk_kb(store_kb, n, L)
            device.mmt_finalize(C32, n, out=K)
            del C32
            K.div_(K.max())                                                      # calcMMt.R:13  MMt/max(MMt) + 0.95 I
            K.diagonal().add_(0.95)
            timed("mmt_s", t0)
        # ---- emma.REMLE and emma.MLE share eigen(S (K + I) S) of this iteration; eigen(K) is computed once
        t0 = time.perf_counter()
        _lib.check(lib.eg_dev_emma_eigen_R_wo_Z(p(K), p(X_d), p(y_d), n, q, p(vals), p(etas), p(U), p(w1), p(w2), p(small), st()))
        lam = vals[: n - q].cpu().numpy()
        et = etas[: n - q].cpu().numpy()
        emma._last = (q, lam, et * et)
        if xi is None:
            U.copy_(K)
            _lib.check(lib.eg_dev_eigen_sym(p(U), n, p(vals), st()))
            xi = vals.cpu().numpy().copy()
            emma._xi = xi
        timed("emma_eigen_s", t0)
        t0 = time.perf_counter()
        vc = emma.REMLE(y, X)
        ml = emma.MLE(y, X, llim=-100.0, ulim=100.0)
        stats["emma_search_s"] += time.perf_counter() - t0
        bic = -2 * ml["ML"] + (q + 1) * math.log(n)
        extBIC.append(bic + 2 * _lchoose(L, q - 1))
        say(f" iteration {itnum}: extBIC = {extBIC[-1]:.4f}")
        if int(np.flatnonzero(np.asarray(extBIC) == min(extBIC))[0]) == len(extBIC) - 1:
            t0 = time.perf_counter()
            if sq is None:
                sq, inv = torch.empty((n, n), **f64), torch.empty((n, n), **f64)
                not_pd, tr = C.c_int(0), C.c_double(0.0)
                _lib.check(lib.eg_dev_sqrt_and_sqrtinv(p(K), n, p(sq), p(inv), p(w1), C.byref(not_pd), C.byref(tr), st()))
                if not_pd.value:
                    raise ValueError("M %*% t(M) is not positive definite")
            _lib.check(lib.eg_dev_calculateH(p(K), n, float(vc["ve"]), float(vc["vg"]), p(w1), st()))          # H in w1
            _lib.check(lib.eg_dev_calculateP(p(w1), p(X_d), n, q, p(w2), p(small), st()))                       # P in w2
            _lib.check(lib.eg_dev_calculate_reduced_a(float(vc["vg"]), p(w2), p(sq), p(y_d), n, p(tmp_n), p(hat_a), st()))
            _lib.check(lib.eg_dev_calculate_reduced_vara(p(X_d), q, float(vc["ve"]), float(vc["vg"]), p(sq), n, p(U), p(w1),
                                                         p(small), st()))                                         # V in U
            timed("algebra_s", t0)
            t0 = time.perf_counter()
            Wp = device.scan_prepare(inv, U, hat_a, n, tmp=w2.view(-1))
            a, vara = device.scan(storeT, L, n, Wp)
            best, idx = device.argmax_tsq(a, vara)
            new_locus = int(idx.item()) + 1
            del Wp, a, vara
            timed("scan_s", t0)
            selected.append(new_locus)
            say(f" iteration {itnum}: picked locus {new_locus}")
        else:
            cont = False
        itnum += 1
        if itnum > maxit:
            cont = False
    if itnum > maxit:
        final = selected
    elif len(selected) > 1:
        final = selected[:-1]
    else:
        final = selected
    torch.cuda.synchronize()
    stats["total_s"] = time.perf_counter() - t_all
    return dict(selected=[int(s) for s in final if not math.isnan(s)],
                all_picked=[int(s) for s in selected if not math.isnan(s)], extBIC=extBIC, vc=vc,
                iterations=itnum - 1, seconds={k: round(v, 4) for k, v in stats.items()})
