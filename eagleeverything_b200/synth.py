"""Synthetic genotype / phenotype generators.  numpy twin of csrc/synth.cu:
the same counter-based splitmix64 hash, so CPU tests and the GPU generator agree byte for byte."""
from __future__ import annotations

import numpy as np

GENO_SEED = 20261018
PHENO_SEED = 7

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return x ^ (x >> np.uint64(31))


def marker_threshold(seed, j):
    with np.errstate(over="ignore"):
        z = _splitmix64(np.uint64(seed) ^ np.uint64(0xA5A5A5A5A5A5A5A5) ^ (j.astype(np.uint64) * np.uint64(0xD1342543DE82EF95)))
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    p = 0.05 + 0.45 * u
    return (p * 4294967296.0).astype(np.uint64).astype(np.uint32)


def genotypes(n, L, seed=GENO_SEED, col_offset=0, n_total=None, row_offset=0):
    """(n, L) uint8 genotypes in {0,1,2}: individual i (row), marker j (column)."""
    n_total = n if n_total is None else n_total
    j = (np.arange(L, dtype=np.uint64) + np.uint64(col_offset))[None, :]
    i = (np.arange(n, dtype=np.uint64) + np.uint64(row_offset))[:, None]
    thr = marker_threshold(seed, j)
    with np.errstate(over="ignore"):
        z = _splitmix64(np.uint64(seed) + (j * np.uint64(n_total) + i) * np.uint64(0x2545F4914F6CDD1D))
    lo = (z & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (z >> np.uint64(32)).astype(np.uint32)
    return ((lo < thr).astype(np.uint8) + (hi < thr).astype(np.uint8))


def ascii_image(G012):
    """Byte-exact M.ascii image (CreateASCIInospace.cpp:119-122): rows x (cols+1) uint8."""
    G = np.asarray(G012, dtype=np.uint8)
    img = np.empty((G.shape[0], G.shape[1] + 1), dtype=np.uint8)
    img[:, :-1] = G + ord("0")
    img[:, -1] = ord("\n")
    return img


def phenotype(G012, seed=PHENO_SEED, n_qtl=5):
    """y = mu + sum_q beta_q * m_{j_q} + e with evenly spaced QTL."""
    G = np.asarray(G012)
    n, L = G.shape
    rng = np.random.default_rng(seed)
    qtl = np.linspace(L // (2 * n_qtl), L - L // (2 * n_qtl) - 1, n_qtl).astype(np.int64)
    beta = np.array([1.0, 0.8, 0.6, 0.5, 0.4])[:n_qtl]
    y = 10.0 + rng.standard_normal(n)
    for b, j in zip(beta, qtl):
        y = y + b * (G[:, j].astype(np.float64) - 1.0)
    return y, qtl


def scan_inputs(n, seed=PHENO_SEED):
    """Synthetic S (sym. pos. def., plays K^-1/2), V (sym.), a_hat for pure-kernel runs."""
    rng = np.random.default_rng(seed + 1000)
    A = rng.standard_normal((n, n)) / np.sqrt(n)
    S = (A + A.T) * 0.5 + np.eye(n) * 2.0
    B = rng.standard_normal((n, n)) / np.sqrt(n)
    V = (B + B.T) * 0.5 + np.eye(n) * 1.5
    a = rng.standard_normal(n)
    return np.asfortranarray(S), np.asfortranarray(V), a
