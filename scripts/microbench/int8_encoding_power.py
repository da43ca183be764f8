"""Does the int8 tensor-core throughput under the 1 kW cap depend on how the genotypes are ENCODED?
cuBLASLt int8 GEMM (torch._int_mm), sustained ~1.5 s per case, A = genotype-like operand with the same
class frequencies (55 % / 35 % / 10 %) in two encodings, B = full-range bytes (digit slices)."""
import subprocess, time, torch
N = 16384
g = torch.Generator(device="cuda"); g.manual_seed(0)
u = torch.rand((N, N), device="cuda", generator=g)
cls = (u > 0.55).to(torch.int8) + (u > 0.90).to(torch.int8)          # 0: 55 %, 1: 35 %, 2: 10 %
enc = {"{-1,0,1} (AA=-1=0xFF)": (cls - 1).to(torch.int8), "{0,1,2} (AA=0)": cls.clone(),
       "{1,0,-1} (AA=+1)": (1 - cls).to(torch.int8), "all zero": torch.zeros_like(cls)}
B = torch.randint(-128, 128, (N, N), dtype=torch.int8, device="cuda", generator=g)
def run(A):
    for _ in range(3): torch._int_mm(A, B)
    torch.cuda.synchronize(); time.sleep(1.0)
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(500): torch._int_mm(A, B)
    e1.record(); torch.cuda.synchronize()
    p.terminate(); out, _ = p.communicate()
    rows = [l.split(",") for l in out.strip().splitlines() if "," in l]; rows = rows[len(rows) // 3:]
    clk = sorted(float(r[0]) for r in rows); pw = sorted(float(r[1]) for r in rows)
    ms = e0.elapsed_time(e1) / 500
    return 2.0 * N ** 3 / ms / 1e9, clk[len(clk) // 2], pw[len(pw) // 2]
for name, A in enc.items():
    t, c, w = run(A)
    print(f"A = {name:26s}: {t:7.0f} TOP/s sustained, sm {c:.0f} MHz, {w:.0f} W", flush=True)
