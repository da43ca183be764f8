"""Decode (K-blocked and row-major destinations) and transpose at a given shape: GB/s of algorithmic traffic."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eagleeverything_b200 import device, synth
n = int(os.environ.get("SW_N", 10000)); L = int(os.environ.get("SW_L", 1000000))
device.init(0)
img = device.synth_ascii(n, L, synth.GENO_SEED)
def timed(fn, rep=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(rep): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rep
kb = torch.empty(((L + 127) // 128, n, 128), dtype=torch.int8, device="cuda")
err = torch.zeros(4, dtype=torch.int32, device="cuda")
bytes_dec = n * (L + 1) + n * L
for mode in ("1", "0"):
    os.environ["EAGLE_DECODE_KB_TILES"] = mode
    ms = timed(lambda: device.decode_kb(img, L + 1, n, L, out=kb, err=err))
    print(f"decode_kb tiles={mode}: {ms:.3f} ms  {bytes_dec / ms / 1e6:.0f} GB/s  err={err[0].item()}", flush=True)
    if mode == "1": ref = kb.clone()
    else: print("identical:", bool(torch.equal(ref, kb)))
del ref
rm = torch.empty((n, device.store_pitch(L)), dtype=torch.int8, device="cuda")
ms = timed(lambda: device.decode(img, L + 1, n, L, out=rm, err=err))
print(f"decode row-major: {ms:.3f} ms  {bytes_dec / ms / 1e6:.0f} GB/s", flush=True)
del rm
tt = torch.empty((L, device.store_pitch(n)), dtype=torch.int8, device="cuda")
ms = timed(lambda: device.transpose_kb(kb, n, L, out=tt))
print(f"transpose_kb: {ms:.3f} ms  {2.0 * n * L / ms / 1e6:.0f} GB/s", flush=True)
