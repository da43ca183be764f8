import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def unpack_demo():
    z = np.load(os.path.join(ROOT, "tests", "golden", "demo.npz"))
    n, L = int(z["n"]), int(z["L"])
    bits = np.unpackbits(z["geno_packed"])[: 2 * n * L].reshape(-1, 2)
    G = (bits[:, 0] * 2 + bits[:, 1]).astype(np.uint8).reshape(n, L)
    return z, G


@pytest.fixture(scope="session")
def demo(tmp_path_factory):
    """The reference's shipped demo data set as M.ascii / Mt.ascii files + oracle golden results."""
    from oracle import np_oracle as npo
    z, G = unpack_demo()
    d = tmp_path_factory.mktemp("demo")
    m, mt = str(d / "M.ascii"), str(d / "Mt.ascii")
    npo.write_ascii(m, G)
    npo.write_ascii(mt, G.T)
    return dict(z=z, G=G, n=G.shape[0], L=G.shape[1], M=m, Mt=mt,
                geno=dict(asciifileM=m, asciifileMt=mt, dim_of_ascii_M=G.shape))


@pytest.fixture(scope="session")
def synth_small(tmp_path_factory):
    """Seeded synthetic data set with ragged sizes (n, L not multiples of any tile)."""
    from eagleeverything_b200 import synth
    from oracle import np_oracle as npo
    n, L = 203, 3001
    G = synth.genotypes(n, L)
    d = tmp_path_factory.mktemp("synth")
    m, mt = str(d / "M.ascii"), str(d / "Mt.ascii")
    npo.write_ascii(m, G)
    npo.write_ascii(mt, G.T)
    return dict(G=G, n=n, L=L, M=m, Mt=mt, geno=dict(asciifileM=m, asciifileMt=mt, dim_of_ascii_M=(n, L)))
