"""Where the time of device.scan_prepare_sharded goes at N ranks (config 3: n = 10,000): CUDA events around each part."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from eagleeverything_b200 import _lib, device
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
lib = device.init(int(os.environ["LOCAL_RANK"]))
n = int(os.environ.get("SW_N", 10000))
g = torch.Generator(device="cuda"); g.manual_seed(1)
S = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); S = (S + S.T) * (0.5 / n ** 0.5); S.diagonal().add_(2.0)
V = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g); V = (V + V.T) * (0.5 / n ** 0.5); V.diagonal().add_(1.5)
a = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
p = lambda t: C.c_void_p(t.data_ptr()); st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
Kpad = (n + 31) // 32 * 32
Wp = torch.empty(lib.eg_scan_wp_elems(n), dtype=torch.float64, device="cuda")
cuts = [min(n, int(round(n * ((1.0 + 3.0 * r / world) ** 0.5 - 1.0) / 32.0)) * 32) for r in range(world)] + [n]
c0, c1 = cuts[rank], cuts[rank + 1]
tmp = torch.empty(n * max(cuts[r + 1] - cuts[r] for r in range(world)), dtype=torch.float64, device="cuda")
for it in range(4):
    ev = [torch.cuda.Event(True) for _ in range(6)]
    dist.barrier(); torch.cuda.synchronize()
    ev[0].record()
    sym = C.c_int(0); _lib.check(lib.eg_dev_inputs_symmetric(p(S), p(V), n, C.byref(sym), st()))
    ev[1].record()
    Wp.zero_()
    ev[2].record()
    _lib.check(lib.eg_dev_scan_prepare_cols(p(S), p(V), n, c0, c1, int(sym.value), p(tmp), p(Wp), st()))
    ev[3].record()
    for r in range(world):
        if cuts[r + 1] > cuts[r]:
            dist.broadcast(Wp[cuts[r] * Kpad:cuts[r + 1] * Kpad], src=r)
    ev[4].record()
    _lib.check(lib.eg_dev_scan_fold(p(S), p(a), n, int(sym.value), p(Wp), st()))
    ev[5].record()
    torch.cuda.synchronize()
    names = ["symmetric?", "zero", "prepare_cols", "broadcasts", "fold"]
    print(f"[rank {rank}] it {it} cols [{c0},{c1}) " + "  ".join(f"{k} {ev[i].elapsed_time(ev[i + 1]):.2f}" for i, k in enumerate(names)), flush=True)
dist.destroy_process_group()
