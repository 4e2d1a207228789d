#!/usr/bin/env python
"""Hot spots from `ncu --page source --csv`: python tools/ncu_hot.py src.csv [topN]
Prints executed-instruction totals by region (address buckets), top stalled instructions, and opcode mix weighted by executions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][ix["Address"]], 16)
tot_exec = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
print("total warp-instr executed", tot_exec, "samples", tot_samp)
ops = collections.Counter()
for r in data:
    op = r[ix["Source"]].split()
    op = [t for t in op if not t.startswith("@")][0].split(".")[0]
    ops[op] += int(r[ix["Instructions Executed"]])
print("opcode mix (% of executed):", {k: round(100 * v / tot_exec, 1) for k, v in ops.most_common(22)})
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("top instructions by samples:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
    a = int(r[ix["Address"]], 16) - base
    st = {s[6:]: int(r[ix[s]]) for s in stalls if int(r[ix[s]]) > 0}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"  {a:#7x} {r[ix['Source']].strip()[:60]:60s} samp {r[ix['# Samples']]:>6s} exec {r[ix['Instructions Executed']]:>9s} {main}")
# executed instruction share by 0x800-byte address bucket
b = collections.Counter()
for r in data:
    a = int(r[ix["Address"]], 16) - base
    b[a // 0x1000] += int(r[ix["Instructions Executed"]])
print("executed share by 4KB code bucket:", {hex(k * 0x1000): round(100 * v / tot_exec, 1) for k, v in sorted(b.items()) if v / tot_exec > 0.01})
