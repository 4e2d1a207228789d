#!/usr/bin/env python
"""Turn the raw captures of a `tools/gpu_full.sh` run (gpurun_out/) into the committed summaries
under profiles/: python tools/make_profiles.py r01"""
import csv, io, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT).stdout

# 1. launch list (gpu__time_duration per launch) + shares
rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if len(r) > 10 and r[0].isdigit()]
lines, tot = [], {}
for r in rows:
    name, ns = r[4].split("(")[0].replace("vaw::", "").replace("<unnamed>::", ""), float(r[-1])
    lines.append(f"{r[0]:>3s} {name:45s} grid {r[8]:18s} block {r[7]:12s} {ns / 1e3:10.1f} us")
    tot[name] = tot.get(name, 0.0) + ns
step = {k: v for k, v in tot.items() if "warp_nv12" in k or "build_pieces" in k}
with open(os.path.join(P, f"{tag}_launches.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 40  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline\n")
    f.write("(cold-cache, serialised launches: compare SHARES, not absolutes)\n\n" + "\n".join(lines) + "\n\nshare of the step:\n")
    s = sum(step.values())
    for k, v in step.items():
        f.write(f"  {k:45s} {100 * v / s:5.1f} %\n")

# 2. full-set metrics of the step's kernels
rep = os.path.join(G, "prof_tiled.ncu-rep")
open(os.path.join(P, f"{tag}_ncu_summary.txt"), "w").write(
    "ncu --set full --clock-control none --import-source on -k regex:warp_nv12|build_pieces -s 6 -c 2  (same command)\n\n"
    + run([sys.executable, "tools/ncu_summary.py", rep]))
src = run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:warp_nv12"])
open("/tmp/_src.csv", "w").write(src)
open(os.path.join(P, f"{tag}_hotspots.txt"), "w").write(
    "per-instruction samples of warp_nv12_tile_kernel (ncu --page source), aggregated by tools/ncu_hot.py\n\n"
    + run([sys.executable, "tools/ncu_hot.py", "/tmp/_src.csv", "25"]))

# 3. DRAM traffic per launch of the dominant kernel -> bench.py's roofline.traffic
raw = list(csv.reader(io.StringIO(run(["ncu", "-i", rep, "--page", "raw", "--csv"]))))
hdr, units = raw[0], raw[1]
traffic = {}
for r in raw[2:]:
    if "warp_nv12" in r[hdr.index("Kernel Name")]:
        def val(k):
            i = hdr.index(k)
            return float(r[i]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[i]]
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        traffic["C3_batch64"] = {"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr,
                                 "kernel": r[hdr.index("Kernel Name")].split("(")[0], "capture": f"profiles/{tag}_ncu_summary.txt"}
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)

# 4. bench lines and parity summary of the same run
for fn in ("bench.json", "bench_ref.json", "bench_C5.json", "bench_C2.json", "bench_C1.json"):
    if os.path.exists(os.path.join(G, fn)):
        open(os.path.join(P, f"{tag}_{fn}"), "w").write(open(os.path.join(G, fn)).read())
if os.path.exists(os.path.join(G, "parity.json")):
    d = json.load(open(os.path.join(G, "parity.json")))
    json.dump(d, open(os.path.join(P, f"{tag}_parity.json"), "w"), indent=1)
print(open(os.path.join(P, f"{tag}_launches.txt")).read()[-600:])
print(json.dumps(traffic))
