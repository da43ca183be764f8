"""CPU tests of the host-side pieces of eagleeverything_b200/am.py (the mirror of AM()'s forward search): EMMA's search over
delta, uniroot and the extended BIC against the oracle's restatements, with the eigendecompositions injected (on a GPU
box they come from the device; here from LAPACK, the same numbers the oracle uses)."""
import math

import numpy as np

from eagleeverything_b200 import am
from oracle import am_driver as oam


def _inject(K, X, y):
    e = am._Emma(K, dict(emma_eigen_s=0.0))
    lam, U = oam.emma_eigen_R_wo_Z(K, X)
    etas = U.T @ y
    e._last = (X.shape[1], lam, etas * etas)
    e._xi = oam.r_eigen_sym(K)[0]
    return e


def test_uniroot_is_the_same_zeroin():
    for f, lo, hi in [(lambda x: math.cos(x) - x, 0.0, 1.0), (lambda x: x ** 3 - 2 * x - 5, 2.0, 3.0),
                      (lambda x: math.exp(-x) - 0.3, -1.0, 4.0), (lambda x: x, -1.0, 0.0)]:
        assert am._uniroot(f, lo, hi) == oam.r_uniroot(f, lo, hi)


def test_emma_search_and_extbic_match_the_oracle(demo):
    z = demo["z"]
    n = demo["n"]
    G = demo["G"].astype(np.float64) - 1.0
    MMt = G @ G.T
    K = MMt / MMt.max() + np.diag(np.full(n, 0.95))
    for y, X in ((z["trait1"], np.ones((n, 1))),
                 (z["trait2"], np.column_stack([np.ones(n), z["pc1"], z["pc2"]])),
                 (z["trait1"], np.column_stack([np.ones(n), G[:, 2206], G[:, 4502]]))):
        y = np.asarray(y, dtype=np.float64)
        e = _inject(K, X, y)
        r, ro = e.REMLE(y, X), oam.emma_REMLE(y, X, K)
        m, mo = e.MLE(y, X, llim=-100.0, ulim=100.0), oam.emma_MLE(y, X, K, llim=-100.0, ulim=100.0)
        for k in ("REML", "delta", "ve", "vg"):
            assert r[k] == ro[k], k
        for k in ("ML", "delta", "ve", "vg"):
            assert m[k] == mo[k], k
        bic = -2 * m["ML"] + (X.shape[1] + 1) * math.log(n) + 2 * am._lchoose(demo["L"], X.shape[1] - 1)
        np.testing.assert_allclose(bic, oam.calc_extBIC(y, X, K, demo["L"]), rtol=1e-13)
    # a design matrix without full column rank: both return zeros (emma_REMLE.R:33-36)
    Xs = np.column_stack([np.ones(n), np.ones(n)])
    assert _inject(K, np.ones((n, 1)), y).REMLE(y, Xs) == oam.emma_REMLE(y, Xs, K) == dict(REML=0.0, delta=0.0, ve=0.0, vg=0.0)


def test_resident_geno_index_translation():
    assert am.ResidentGeno._idx([am.NA]) == [] and am.ResidentGeno._idx([am.NA, 5.0]) == []
    assert am.ResidentGeno._idx([3.0, 9.0]) == [3, 9]
