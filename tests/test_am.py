"""The product-side mirror of AM()'s forward search (eagleeverything_b200/am.py; reference R/AM.R:395-504 with
find_qtl, emma.REMLE / emma.MLE and calc_extBIC) against the golden results of the shipped demo data (SURVEY.md section 4,
tests/golden/demo.npz) and against the oracle's restated driver on a synthetic set: identical selected-QTL sequence,
extBIC trace to 1e-8 (EMMA's root search stops at uniroot's 1e-4 tolerance; the eigenvalues come from cuSOLVER here and
from LAPACK there)."""
import numpy as np
import pytest

from eagleeverything_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def am():
    from eagleeverything_b200 import am as m
    from eagleeverything_b200 import device
    device.init(0)
    return m


def test_demo_forward_search_golden(am, demo):
    z = demo["z"]
    g = am.FileGeno(demo["M"], demo["Mt"], (demo["n"], demo["L"]))
    r = am.AM(g, z["trait1"])
    assert r["selected"] == list(z["am1_selected"]) == [2207, 4503, 873]
    assert r["all_picked"] == list(z["am1_all_picked"])
    np.testing.assert_allclose(r["extBIC"], z["am1_extBIC"], rtol=1e-8)
    X0 = np.column_stack([np.ones(demo["n"]), z["pc1"], z["pc2"]])
    r2 = am.AM(g, z["trait2"], X0=X0)
    assert r2["selected"] == list(z["am2_selected"]) and r2["all_picked"] == list(z["am2_all_picked"])
    np.testing.assert_allclose(r2["extBIC"], z["am2_extBIC"], rtol=1e-8)
    assert r["seconds"]["total_s"] > 0 and r["iterations"] == len(r["extBIC"])


def test_resident_stores_equal_files_and_oracle_driver(am, synth_small):
    from eagleeverything_b200 import api
    from oracle import am_driver as oam
    from oracle import eagle_oracle as eo
    s = synth_small
    y, qtl = synth.phenotype(s["G"])
    ro = oam.AM(eo, s["geno"], y, maxit=6)
    rf = am.AM(am.FileGeno(s["M"], s["Mt"], (s["n"], s["L"])), y, maxit=6)
    M = api.GenotypeStore.from_host_ascii(synth.ascii_image(s["G"]), s["n"], s["L"])
    rr = am.AM(am.ResidentGeno(M, M.transpose()), y, maxit=6)
    assert rf["all_picked"] == rr["all_picked"] == ro["all_picked"] and rf["selected"] == rr["selected"] == ro["selected"]
    np.testing.assert_allclose(rf["extBIC"], ro["extBIC"], rtol=1e-8)
    np.testing.assert_allclose(rr["extBIC"], rf["extBIC"], rtol=1e-12)
    # maxit reached: every picked locus stays in the model (AM.R:477-481)
    r2 = am.AM(am.FileGeno(s["M"], s["Mt"], (s["n"], s["L"])), y, maxit=2)
    assert r2["selected"] == r2["all_picked"] == ro["all_picked"][:2]


def test_search_with_everything_resident_in_hbm(am, demo, synth_small):
    import torch
    from eagleeverything_b200 import device

    def stores(G):
        n, L = G.shape
        img = torch.from_numpy(np.concatenate([synth.ascii_image(G).reshape(-1), np.zeros(64, np.uint8)])).cuda()
        kb, err = device.decode_kb(img, L + 1, n, L)
        assert err[0].item() == 0
        return kb, device.transpose_kb(kb, n, L)

    z = demo["z"]
    kb, t = stores(demo["G"])
    r = am.AM_resident(kb, t, demo["n"], demo["L"], z["trait1"])
    assert r["selected"] == list(z["am1_selected"]) and r["all_picked"] == list(z["am1_all_picked"])
    np.testing.assert_allclose(r["extBIC"], z["am1_extBIC"], rtol=1e-8)
    X0 = np.column_stack([np.ones(demo["n"]), z["pc1"], z["pc2"]])
    r2 = am.AM_resident(kb, t, demo["n"], demo["L"], z["trait2"], X0=X0)
    assert r2["selected"] == list(z["am2_selected"]) and r2["all_picked"] == list(z["am2_all_picked"])
    np.testing.assert_allclose(r2["extBIC"], z["am2_extBIC"], rtol=1e-8)
    s = synth_small
    y, _ = synth.phenotype(s["G"])
    kb, t = stores(s["G"])
    rr = am.AM_resident(kb, t, s["n"], s["L"], y, maxit=6)
    rf = am.AM(am.FileGeno(s["M"], s["Mt"], (s["n"], s["L"])), y, maxit=6)
    assert rr["all_picked"] == rf["all_picked"] and rr["selected"] == rf["selected"]
    np.testing.assert_allclose(rr["extBIC"], rf["extBIC"], rtol=1e-8)


def test_repeated_measures_search(am):
    """BASELINE config 5's model at test size: records x individuals incidence matrix Z and fixed-effect covariates.
    am.AM_resident(Z = ...) -- EMMA's Z branches through the secular solve, the Z-aware scan -- against the dense
    restatement (oracle/am_driver.py::AM_Z: R/emma_eigen_R_w_Z.R, R/emma_REMLE.R:78-131, R/emma_MLE.R:57-117 and
    H = ve I + vg Z K Z'): same picks, extBIC to 1e-8, and the same with Z given as an index vector."""
    import torch
    from eagleeverything_b200 import device
    from oracle import am_driver as oam
    rng = np.random.default_rng(21)
    t, L, nrec = 150, 2000, 230
    G = synth.genotypes(t, L, seed=77)
    idx = np.concatenate([np.arange(t), rng.integers(0, t, nrec - t)])
    rng.shuffle(idx)
    Z = np.zeros((nrec, t))
    Z[np.arange(nrec), idx] = 1
    X0 = np.column_stack([np.ones(nrec), rng.standard_normal(nrec), rng.integers(0, 2, nrec).astype(float)])
    M = G.astype(np.float64) - 1.0
    y = 3.0 + 0.4 * X0[:, 1] + rng.standard_normal(nrec) + Z @ (1.2 * M[:, 300] - 0.9 * M[:, 1500])
    ro = oam.AM_Z(M, y, X0, Z, L, maxit=4)
    img = torch.from_numpy(np.concatenate([synth.ascii_image(G).reshape(-1), np.zeros(64, np.uint8)])).cuda()
    kb, err = device.decode_kb(img, L + 1, t, L)
    tT = device.transpose_kb(kb, t, L)
    rr = am.AM_resident(kb, tT, t, L, y, X0=X0, maxit=4, Z=Z)
    assert rr["all_picked"] == ro["all_picked"] and rr["selected"] == ro["selected"], (rr["all_picked"], ro["all_picked"])
    assert 301 in rr["all_picked"] and 1501 in rr["all_picked"]
    np.testing.assert_allclose(rr["extBIC"], ro["extBIC"], rtol=1e-8)
    r2 = am.AM_resident(kb, tT, t, L, y, X0=X0, maxit=4, Z=idx)
    assert r2["all_picked"] == rr["all_picked"] and r2["extBIC"] == rr["extBIC"]
    with pytest.raises(ValueError):
        am.AM_resident(kb, tT, t, L, y[:-1], X0=X0[:-1], maxit=2, Z=Z[:-1][:, : t - 1])
