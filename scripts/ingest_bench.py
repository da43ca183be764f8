"""Ingest kernels at a given shape: tokeniser (count + scan + emit) and ASCII encoder, in GB/s of
algorithmic traffic, with the CPU restatement (oracle, one core, file to file) on a bounded sample beside them."""
import ctypes as C
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from eagleeverything_b200 import _lib, device, synth

n = int(os.environ.get("SW_N", 10000)); L = int(os.environ.get("SW_L", 100000))
lib = device.init(0)
vp = lambda t: C.c_void_p(t.data_ptr())
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn, rep=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(rep): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rep


img = device.synth_ascii(n, L, synth.GENO_SEED)[: n * (L + 1)].view(n, L + 1)
text = torch.full((n, 2 * L), 32, dtype=torch.uint8, device="cuda")      # "g g g ... g\n": 2 bytes per genotype
text[:, 0::2] = img[:, :L]
text[:, -1] = 10
nbytes = text.numel()
buf = torch.zeros(nbytes + 64, dtype=torch.uint8, device="cuda"); buf[:nbytes] = text.view(-1); del text
nch = int(lib.eg_tokenise_chunks(nbytes))
counts = torch.empty(2 * nch, dtype=torch.int32, device="cuda")
prefix = torch.empty(2 * (nch + 1), dtype=torch.int64, device="cuda")
out = torch.empty(n * (L + 1) + 16, dtype=torch.uint8, device="cuda")
err = torch.empty(1, dtype=torch.int64, device="cuda")
scan = lambda: _lib.check(lib.eg_dev_tokenise_scan(vp(buf), nbytes, vp(counts), vp(prefix), st()))
def emit():
    err.fill_(-1)
    _lib.check(lib.eg_dev_tokenise_emit(vp(buf), nbytes, vp(prefix), L, b"0", b"1", b"2", b"NA", vp(out), n, vp(err), st()))
ms_scan, ms_emit = timed(scan), timed(emit)
assert err.item() == -1 and prefix[-2].item() == n * L and prefix[-1].item() == n
assert torch.equal(out[: n * (L + 1)].view(n, L + 1), img)
alg = nbytes + n * (L + 1)
print(f"tokeniser n={n} L={L}: text {nbytes / 1e9:.2f} GB -> {n * (L + 1) / 1e9:.2f} GB;  count+scan {ms_scan:.3f} ms, emit {ms_emit:.3f} ms, "
      f"total {ms_scan + ms_emit:.3f} ms = {alg / (ms_scan + ms_emit) / 1e6:.0f} GB/s algorithmic "
      f"({(2 * nbytes + n * (L + 1)) / (ms_scan + ms_emit) / 1e6:.0f} GB/s moved)", flush=True)
del buf, counts, prefix

# encoder: Mt store (L x n) -> Mt.ascii image
imgT = img[:, :L].t().contiguous()                                           # L x n characters
pitch = device.store_pitch(n)
store = torch.zeros((L, pitch), dtype=torch.int8, device="cuda")
store[:, :n] = (49 - imgT.to(torch.int16)).to(torch.int8)                     # 1 - code
enc = torch.empty(L * (n + 1) + 16, dtype=torch.uint8, device="cuda")
ms_enc = timed(lambda: _lib.check(lib.eg_dev_encode_ascii(vp(store), pitch, n, 0, L, vp(enc), st())))
want = torch.cat([imgT, torch.full((L, 1), 10, dtype=torch.uint8, device="cuda")], 1)
assert torch.equal(enc[: L * (n + 1)].view(L, n + 1), want)
print(f"encoder {L} x {n}: {ms_enc:.3f} ms = {2.0 * n * L / ms_enc / 1e6:.0f} GB/s", flush=True)

# CPU restatement on a bounded sample (file to file, one core)
from oracle import eagle_oracle as eo
rows = max(1, min(n, int(4e8 // (2 * L))))
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "s.txt")
    t = np.full((rows, 2 * L), 32, np.uint8); t[:, 0::2] = img[:rows, :L].cpu().numpy(); t[:, -1] = 10; t.tofile(p)
    t0 = time.perf_counter()
    ok, _ = eo.createM_ASCII_rcpp(p, os.path.join(d, "s.ascii"), "text", "0", "1", "2", 8.0, (rows, L), True, "NA")
    dt = time.perf_counter() - t0
    print(f"CPU restatement (1 core, {rows} rows = {rows * 2 * L / 1e6:.0f} MB of text, file to file): {dt:.2f} s = "
          f"{rows * (3 * L + 1) / dt / 1e9:.3f} GB/s algorithmic", flush=True)
