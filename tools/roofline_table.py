"""digest of an `ncu --page raw --csv` export: per kernel launch duration, DRAM GB/s and % of the measured HBM peak,
tensor-pipe activity, L2 throughput -> markdown table (profiles/r01_roofline_summary_*.md)"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r01_ncu_raw_run22.csv")
peaks = {"hbm_gbs": 6486.1}
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peaks.update(json.load(open(p)))
rows = list(csv.reader(open(src)))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
f = lambda r, k: float(r[ix[k]]) if r[ix[k]] not in ("", "n/a") else float("nan")
agg = {}
for r in data:
    name = r[ix["Kernel Name"]].replace("void ", "").split("(")[0]
    grid = int(float(r[ix["launch__grid_size"]]))
    key = (name, grid)
    us = f(r, "gpu__time_duration.sum")
    mb = f(r, "dram__bytes_read.sum") + f(r, "dram__bytes_write.sum")
    a = agg.setdefault(key, dict(n=0, us=0.0, mb=0.0, tens=0.0, lts=0.0, regs=int(float(r[ix["launch__registers_per_thread"]]))))
    a["n"] += 1; a["us"] += us; a["mb"] += mb
    a["tens"] += f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    a["lts"] += f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
print(f"| kernel (grid) | launches | avg µs | DRAM MB/launch | DRAM GB/s | % of HBM peak ({peaks['hbm_gbs']:.0f} GB/s) | tensor pipe % | L2 % | regs |")
print("|---|---|---|---|---|---|---|---|---|")
for (name, grid), a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    n = a["n"]; us = a["us"] / n; mb = a["mb"] / n
    gbs = mb / us * 1e3 if us > 0 else 0.0
    print(f"| `{name}` ({grid}) | {n} | {us:.1f} | {mb:.1f} | {gbs:.0f} | {100 * gbs / peaks['hbm_gbs']:.1f} | {a['tens'] / n:.1f} | {a['lts'] / n:.1f} | {a['regs']} |")
