#!/usr/bin/env python
"""Where a kernel's warps spend their time: python tools/ncu_regions.py rep kernel-regex
Splits the SASS into runs of equal execution count and prints, per run, instructions, executions, stall samples
(share of warp residency) and the top stall reasons."""
import csv, io, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
minshare = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_e = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_s = sum(int(r[ix["# Samples"]]) for r in data)
print("instructions", len(data), "executed", tot_e, "samples", tot_s)
def flush(lo, hi):
    e = sum(int(r[ix["Instructions Executed"]]) for r in data[lo:hi]); s = sum(int(r[ix["# Samples"]]) for r in data[lo:hi])
    if e / tot_e * 100 < minshare and s / tot_s * 100 < minshare: return
    c = collections.Counter()
    for r in data[lo:hi]:
        for st in stalls:
            v = int(r[ix[st]] or 0)
            if v: c[st[6:]] += v
    print(f"{lo:5d}-{hi-1:5d} n={hi-lo:4d} x{int(data[lo][ix['Instructions Executed']]):9d} exec {100*e/tot_e:5.1f}% time {100*s/tot_s:5.1f}%  {data[lo][ix['Source']].strip()[:40]:40s} {[(k, v) for k, v in c.most_common(4)]}")
start = 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(int(data[i][ix["Instructions Executed"]]) - int(data[start][ix["Instructions Executed"]])) > 0.03 * max(int(data[start][ix["Instructions Executed"]]), 1):
        flush(start, i); start = i
