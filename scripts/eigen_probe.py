"""How long does the one library call of the search take, and what does it depend on?  cusolverDnDsyevd of order n
(eg_dev_eigen_sym), repeated in one process; run with different OMP_NUM_THREADS / after other GPU work."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eagleeverything_b200 import _lib, device
lib = device.init(0)
n = int(os.environ.get("PROBE_N", "10000"))
g = torch.Generator(device="cuda"); g.manual_seed(1)
A = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)
K = (A @ A.T) / n + 0.95 * torch.eye(n, dtype=torch.float64, device="cuda")
vals = torch.empty(n, dtype=torch.float64, device="cuda")
for rep in range(4):
    U = K.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(lib.eg_dev_eigen_sym(C.c_void_p(U.data_ptr()), n, C.c_void_p(vals.data_ptr()), None))
    torch.cuda.synchronize()
    print(f"OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')} dsyevd n={n} call {rep}: {time.perf_counter() - t0:.3f} s", flush=True)
