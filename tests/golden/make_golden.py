"""Regenerates tests/golden/demo.npz.  Run in the build container, where /root/reference is mounted:

    python tests/golden/make_golden.py

Inputs : /root/reference/MyPackage/{genoDemo.dat,phenoDemo.dat}  (the reference's shipped demo data)
Outputs: the demo genotypes (2-bit packed), phenotype columns, and the results the CPU oracle
         (oracle/eagle_oracle.c + oracle/am_driver.py) produces on them: MM^T digest and corner,
         the forward-search trace, and the first-iteration scan inputs/outputs.
The reference itself ships no expected outputs; these pin the oracle against
regressions and give the GPU box a fixture that does not need /root/reference.
"""
import hashlib
import os
import sys
import tempfile

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import am_driver as am  # noqa: E402
from oracle import eagle_oracle as eo  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402

REF = "/root/reference/MyPackage"


def main():
    d = tempfile.mkdtemp()
    G = npo.create_ascii_nospace(f"{REF}/genoDemo.dat", f"{d}/M.ascii")  # AA=0, AB=1, BB=2
    npo.write_ascii(f"{d}/Mt.ascii", G.T)
    n, L = G.shape
    ph = pd.read_csv(f"{REF}/phenoDemo.dat", sep=r"\s+")
    MMt = eo.calculateMMt_rcpp(f"{d}/M.ascii", 8, 8, [eo.NA_REAL], (n, L))
    geno = dict(asciifileM=f"{d}/M.ascii", asciifileMt=f"{d}/Mt.ascii", dim_of_ascii_M=(n, L))
    r1 = am.AM(eo, geno, ph["trait1"].values, keep_trace=True)
    X2 = np.column_stack([np.ones(n), ph["pc1"].values, ph["pc2"].values])
    r2 = am.AM(eo, geno, ph["trait2"].values, X0=X2)
    t0 = r1["trace"][0]
    packed = np.packbits(np.unpackbits(G.reshape(-1, 1), axis=1)[:, 6:].reshape(-1))  # 2 bits per genotype
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "demo.npz"),
        geno_packed=packed, n=n, L=L,
        trait1=ph["trait1"].values, trait2=ph["trait2"].values, pc1=ph["pc1"].values, pc2=ph["pc2"].values,
        mmt_sha256=hashlib.sha256(MMt.astype("<i4").tobytes()).hexdigest(),
        mmt_corner=MMt[:4, :4], mmt_trace=np.trace(MMt), mmt_sum=MMt.sum(), mmt_max=MMt.max(), mmt_min=MMt.min(),
        am1_selected=np.array(r1["selected"]), am1_all_picked=np.array(r1["all_picked"]), am1_extBIC=np.array(r1["extBIC"]),
        am2_selected=np.array(r2["selected"]), am2_all_picked=np.array(r2["all_picked"]), am2_extBIC=np.array(r2["extBIC"]),
        it1_S=t0["S"], it1_V=t0["V"], it1_hat_a=t0["hat_a"], it1_a=t0["a"], it1_vara=t0["vara"],
    )
    print("wrote demo.npz", n, L, r1["selected"], r2["all_picked"])


if __name__ == "__main__":
    main()
