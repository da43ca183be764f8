"""Stage timing of the column-sharded pre-products under torchrun (N ranks)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from eagleeverything_b200 import device, _lib
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = device.init(local)
n = 10000
g = torch.Generator(device="cuda"); g.manual_seed(1)
S = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); S = (S + S.T) * (0.5 / n ** 0.5); S.diagonal().add_(2.0)
V = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); V = (V + V.T) * (0.5 / n ** 0.5); V.diagonal().add_(1.5)
a = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
Kpad = (n + 31) // 32 * 32
Wp = torch.empty(lib.eg_scan_wp_elems(n), dtype=torch.float64, device="cuda")
tmp = torch.empty(n * n, dtype=torch.float64, device="cuda")
vp = lambda t: C.c_void_p(t.data_ptr())
cuts = [min(n, int(round(n * ((1.0 + 3.0 * r / world) ** 0.5 - 1.0) / 32.0)) * 32) for r in range(world)] + [n]
for rep in range(3):
    ev = [torch.cuda.Event(True) for _ in range(5 + world)]
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ev[0].record()
    sym = C.c_int(0); _lib.check(lib.eg_dev_inputs_symmetric(vp(S), vp(V), n, C.byref(sym), None))
    ev[1].record()
    Wp.zero_()
    ev[2].record()
    _lib.check(lib.eg_dev_scan_prepare_cols(vp(S), vp(V), n, cuts[rank], cuts[rank + 1], 1, vp(tmp), vp(Wp), None))
    ev[3].record()
    for r in range(world):
        dist.broadcast(Wp[cuts[r] * Kpad:cuts[r + 1] * Kpad], src=r)
        ev[4 + r].record()
    _lib.check(lib.eg_dev_scan_fold(vp(S), vp(a), n, 1, vp(Wp), None))
    ev[4 + world].record()
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(4 + world)]
    print(f"rank {rank} rep {rep}: sym {t[0]:.2f} zero {t[1]:.2f} cols {t[2]:.2f} bcast {[round(x, 2) for x in t[3:3 + world]]} fold {t[3 + world]:.2f} total {ev[0].elapsed_time(ev[4 + world]):.2f}", flush=True)
dist.destroy_process_group()
