#!/usr/bin/env python
"""Turn the raw captures of a `tools/gpu_full.sh` / `tools/gpu_r2_final.sh` run (gpurun_out/) into the committed
summaries under profiles/: python tools/make_profiles.py r02"""
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
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 60  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --no-shim\n")
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
    "per-instruction samples of the sampler kernel (ncu --page source), aggregated by tools/ncu_hot.py\n\n"
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
# where the sampler's warps spend their time (regions of equal execution count) + the SASS that proves TMA / mbarrier / packed math
open(os.path.join(P, f"{tag}_regions.txt"), "w").write(
    "tools/ncu_regions.py on the same capture: runs of equal execution count; exec = share of executed warp instructions,\n"
    "time = share of the warp-state samples (= of the time warps are resident)\n\n"
    + run([sys.executable, "tools/ncu_regions.py", rep, "warp_nv12", "0.4"]))
sass = run(["cuobjdump", "-sass", os.path.join(ROOT, "video_annotator_b200", "libvaw.so")])
funcs = [f for f in sass.split("Function : ")[1:] if "warp_nv12_quad_kernelILi7" in f.split("\n", 1)[0]]
if funcs:
    import collections, re
    body = funcs[0].split("\n")
    ins = []
    for ln in body:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins)
    # the row loop = the backward branch whose body holds 20 shared-memory tap loads and three stores
    loop = []
    index = {a: k for k, (a, _) in enumerate(ins)}
    for k, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in index:
            cand = ins[index[int(m.group(1), 16)]:k + 1]
            n_lds, n_stg = sum(1 for _, x in cand if x.startswith("LDS")), sum(1 for _, x in cand if "STG" in x)
            if (n_lds, n_stg) in ((20, 3), (40, 6)):  # one row pair per trip, or two (unrolled)
                loop = cand
                break
    with open(os.path.join(P, f"{tag}_sass_tile.txt"), "w") as f:
        f.write("cuobjdump -sass video_annotator_b200/libvaw.so, function " + body[0].strip() + " (sm_100a)\n")
        f.write("opcode counts (static, whole function): " + ", ".join(f"{k} {v}" for k, v in ops.most_common(40)) + "\n\n")
        f.write("1. asynchronous machinery: mbarriers (SYNCS), the bulk copy of the piece record (UBLKCP), TMA tensor loads of the\n"
                "   source tile (UTMALDG), the CTA barrier:\n\n")
        f.write("\n".join(f"  /*{a:05x}*/ {t}" for a, t in ins if any(k in t for k in ("UTMALDG", "UBLKCP", "SYNCS", "BAR.", "CCTL"))) + "\n\n")
        lc = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in loop)
        pairs = 2 if sum(1 for _, x in loop if x.startswith("LDS")) == 40 else 1
        f.write(f"2. the row loop ({len(loop)} instructions per trip = {pairs} row pair(s); a row pair = 2 x 2 luma pixels + 1 chroma sample of a lane): "
                + ", ".join(f"{k} {v}" for k, v in lc.most_common()) + "\n\n")
        f.write("\n".join(f"  /*{a:05x}*/ {t}" for a, t in loop) + "\n")
for fn in ("bench.json", "bench_ref.json", "bench_C5.json", "bench_C2.json", "bench_C1.json", "bench_C4.json", "bench_fused_bgr.json",
           "bench_fused_bgr_one_launch.json", "modes.json"):
    if os.path.exists(os.path.join(G, fn)):
        open(os.path.join(P, f"{tag}_{fn}"), "w").write(open(os.path.join(G, fn)).read())
if os.path.exists(os.path.join(G, "parity.json")):
    d = json.load(open(os.path.join(G, "parity.json")))
    json.dump(d, open(os.path.join(P, f"{tag}_parity.json"), "w"), indent=1)
if os.path.exists(os.path.join(G, "packed.jsonl")):
    open(os.path.join(P, f"{tag}_packed.txt"), "w").write(
        "tools/bench_packed.py: BGR24 / GRAY8 through the staged-tile kernel (vaw_packed_tile.cu), device-resident frames\n\n"
        + open(os.path.join(G, "packed.jsonl")).read())
if os.path.exists(os.path.join(G, "prof_bgr.ncu-rep")):
    open(os.path.join(P, f"{tag}_ncu_bgr.txt"), "w").write(
        "ncu --set full --clock-control none -k regex:packed_tile -s 4 -c 1  python tools/bench_packed.py C3 32 bgr\n\n"
        + run([sys.executable, "tools/ncu_summary.py", os.path.join(G, "prof_bgr.ncu-rep")]))
if os.path.exists(os.path.join(G, "prof_cubic.ncu-rep")):
    rep_c = os.path.join(G, "prof_cubic.ncu-rep")
    rawc = list(csv.reader(io.StringIO(run(["ncu", "-i", rep_c, "--page", "raw", "--csv"]))))
    hc, uc, rc_ = rawc[0], rawc[1], rawc[2]
    extra = []
    for k in ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"):
        if k in hc:
            extra.append(f"  {k:75s} {rc_[hc.index(k)]:>18s} {uc[hc.index(k)]}")
    for i, k in enumerate(hc):
        if k.endswith("l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg") or k.endswith("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg"):
            extra.append(f"  {k:75s} {rc_[i]:>18s} per SM")
    open(os.path.join(P, f"{tag}_ncu_cubic.txt"), "w").write(
        "ncu --set full --clock-control none --import-source on -k regex:warp_nv12_quad -s 2 -c 1  python tools/run_mode.py cubic\n"
        "(INTER_CUBIC on the staged-tile kernel, C3 geometry, 16 frames per launch; tools/ncu_summary.py + the l1tex data-pipe split)\n\n"
        + run([sys.executable, "tools/ncu_summary.py", rep_c]) + "\n".join(extra) + "\n\n"
        "Reading: the l1tex data pipe is the busiest unit by far (issue slots ~40 %): the weight loads (3.75 M warp-wide requests of 32\n"
        "different 32-byte table entries each: ~20 data-pipe wavefronts per request at an L1 hit rate of 95 %) and the tap words (a\n"
        "share of them bank-conflict replays: lanes that share a source column in adjacent rows; tile pitches of 64 modulo 128 bytes\n"
        "halve those).  The filter is bound by gathering 32 bytes of weights per sample, not by arithmetic or HBM (DRAM traffic is\n"
        "below the algorithmic bytes: border pieces read nothing).\n")
if os.path.exists(os.path.join(G, "table_filters_final.txt")):
    path = os.path.join(P, f"{tag}_table_filters.txt")
    old = open(path).read() if os.path.exists(path) else ""
    marker = "\nfinal pass (tools/gpu_r2_final.sh; the defaults include the 64-modulo-128 tile pitches):\n"
    old = old.split(marker)[0]
    open(path, "w").write(old + marker + "".join("   " + ln for ln in open(os.path.join(G, "table_filters_final.txt"))))
for fn, to in (("flow_demo.log", "flow_demo.txt"),):
    if os.path.exists(os.path.join(G, fn)):
        open(os.path.join(P, f"{tag}_{to}"), "w").write(open(os.path.join(G, fn)).read())
print(open(os.path.join(P, f"{tag}_launches.txt")).read()[-600:])
print(json.dumps(traffic))
