"""Summarise an `ncu --set full` report: per kernel the metrics DESIGN.md / bench.py quote, as markdown on stdout,
and (with --traffic OUT.json WORKLOAD) the per-launch DRAM traffic table bench.py reads.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep "title" [--traffic profiles/traffic.json c3]
"""
import csv, io, json, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "launch__registers_per_thread", "smsp__inst_executed.sum"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    rep, title = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    print(f"# {title}\n")
    traffic = {}
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("eg::", "").split("<")[0]
        print(f"## {name}")
        vals = {}
        for w in WANT:
            if w in h:
                i = h.index(w)
                vals[w] = (r[i], units[i])
                print(f"{w} = {r[i]} {units[i]}")
        try:
            rd = float(vals["dram__bytes_read.sum"][0].replace(",", "")) * UNIT[vals["dram__bytes_read.sum"][1]]
            wr = float(vals["dram__bytes_write.sum"][0].replace(",", "")) * UNIT[vals["dram__bytes_write.sum"][1]]
            print(f"dram traffic (read+write) = {(rd + wr) / 1e9:.2f} GB")
            rec = traffic.setdefault(name, {"dram_bytes_per_launch": 0.0, "launches": 0})
            rec["dram_bytes_per_launch"] += rd + wr
            rec["launches"] += 1
            tp = vals.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
            if tp and float(tp[0]) > 0:
                rec["tensor_pipe_active_pct"] = float(tp[0])
        except Exception:  # noqa: BLE001
            pass
        print()
    if "--traffic" in sys.argv:
        i = sys.argv.index("--traffic")
        path, workload = sys.argv[i + 1], sys.argv[i + 2]
        try:
            cur = json.load(open(path))
        except Exception:  # noqa: BLE001
            cur = {}
        for k, v in traffic.items():
            v.update({"workload": workload, "n_gpus": 1, "source": sys.argv[i + 3] if len(sys.argv) > i + 3 else rep,
                      "note": "sum over the kernel's launches in one step" if v["launches"] > 1 else "one launch"})
            cur[k] = v
        json.dump(cur, open(path, "w"), indent=1)


main()
