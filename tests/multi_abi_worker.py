"""Worker of tests/test_gpu_multi_abi.py (a fresh interpreter per case: the multi-GPU set is process-wide state):
the hot-path exports and the store API on an eg_init_multi set against the same calls on one GPU, bit for bit."""
import faulthandler
faulthandler.dump_traceback_later(int(__import__("os").environ.get("EAGLE_TEST_WATCHDOG_S", "240")), exit=True)
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eagleeverything_b200 import _lib, api, synth
from oracle import np_oracle as npo
ngpu, n, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
G = synth.genotypes(n, L)
S, V, a = synth.scan_inputs(n)
NA = api.NA_REAL
d = tempfile.mkdtemp()
m, mt = os.path.join(d, "M.ascii"), os.path.join(d, "Mt.ascii")
npo.write_ascii(m, G); npo.write_ascii(mt, G.T)
sel = [3.0, float(L - 2), float(L // 2)]
def run():
    out = {}
    out["K"] = api.calculateMMt_rcpp(m, 8, 1, [NA], (n, L))
    out["K2"] = api.calculateMMt_rcpp(m, 8, 1, sel, (n, L))          # cached store, zeroed columns
    r = api.calculate_a_and_vara_rcpp(mt, [NA], S, V, 8, (L, n), a)
    out["a"], out["vara"] = r["a"].copy(), r["vara"].copy()
    r = api.calculate_a_and_vara_rcpp(mt, sel, S, V, 8, (L, n), a)
    out["a2"], out["vara2"] = r["a"].copy(), r["vara"].copy()
    out["cols"] = np.stack([api.extract_geno_rcpp(m, 8, j, (n, L)) for j in (0, 127, 128, L // 2, L - 1)])
    out["ar"] = api.calculate_reduced_a_rcpp(mt, 1.7, V, a, 8, (n, L), [NA])
    img = synth.ascii_image(G)
    M = api.GenotypeStore.from_host_ascii(img, n, L)
    Mt = M.transpose()
    out["sK"] = M.mmt([5, L - 1])
    out["sa"], out["sv"] = Mt.a_and_vara(S, V, a, [7])
    out["scol"] = M.extract_col(L - 3)
    out["info"] = np.array([M.info()["rows"], M.info()["cols"], Mt.info()["rows"], Mt.info()["cols"]])
    M.free(); Mt.free()
    return out
_lib.check(_lib.require_gpu().eg_init(0))
one = run()
assert api.gpu_count() == 1
assert api.init_multi(ngpu) == ngpu
many = run()
for k in one:
    assert np.array_equal(one[k], many[k]), k
assert np.array_equal(many["cols"][3], G[:, L // 2].astype(np.int32) - 1)
# back to one GPU: the set is torn down and rebuilt
api.shutdown()
_lib.check(_lib.require_gpu().eg_init(0))
again = api.calculateMMt_rcpp(m, 8, 1, [NA], (n, L))
assert np.array_equal(again, one["K"])
print("MULTI_ABI_OK", ngpu)
