"""rank CUDA source lines of one kernel in an .ncu-rep by warp-stall samples.
usage: python tools/ncu_hot_lines.py REP KERNEL_REGEX [TOP] [LAUNCH_INDEX]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
       "--print-source", "sass,cuda"]
if len(sys.argv) > 4:
    cmd += ["--launch-skip", sys.argv[4], "--launch-count", "1"]   # noqa
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, hdr, out = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        try:
            s = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        st = {k[6:]: int(v) for k, v in zip(hdr, r) if k.startswith("stall_") and "Not" not in k and v not in ("", "0")}
        out.append((s, cur, r[0], r[1].strip()[:110], st))
tot = sum(o[0] for o in out)
print("total samples", tot)
agg = {}
for o in out:
    for k, v in o[4].items():
        agg[k] = agg.get(k, 0) + v
print("stall mix:", {k: round(100 * v / max(tot, 1), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for o in sorted(out, key=lambda x: -x[0])[:top]:
    print(f"{100 * o[0] / max(tot, 1):5.1f}% {o[1]}:{o[2]}  {o[3]}  {dict(sorted(o[4].items(), key=lambda kv: -kv[1])[:3])}")
