"""Prints the headline fields of a bench.py JSON line (first line starting with '{' of the given file)."""
import json
import sys

for path in sys.argv[1:]:
    d = None
    for line in open(path):
        if line.startswith("{"):
            d = json.loads(line)
            break
    if d is None:
        print(path, "no JSON line")
        continue
    print(path)
    print("  value %.0f markers/s  step %.2f ms  N=%d" % (d["value"], d["ms_per_step"], d["n_gpus"]))
    print("  stages", {k: round(v, 2) for k, v in d["stage_ms"].items()})
    r = d["roofline"]
    print("  roofline", {k: r.get(k) for k in ("kernel", "achieved", "peak", "frac", "kernel_ms", "digits", "executed_ops", "algorithmic_ops")})
    e = d.get("e2e") or {}
    print("  e2e ms", e.get("ms_per_step"), "dropin", {k: v.get("ms_per_step") for k, v in (e.get("dropin_files") or {}).items() if isinstance(v, dict)})
    fs = d.get("forward_search") or {}
    print("  search", fs.get("seconds"), fs.get("all_picked_1based"), fs.get("extBIC"), fs.get("note"))
    print("  scan_kernels", fs.get("scan_kernels"))
    print("  checks", d.get("checks"))
    print("  cpu", (d.get("cpu_baseline") or {}).get("value"), "clocks", d.get("clocks"))
