#!/usr/bin/env python
"""Static instruction mix of a kernel's loops: python tools/sass_loop.py lib.so kernel-substring
Lists every backward branch (loop) with its instruction count by opcode class."""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print("==", name[:100], "total", len(ins))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                body = ins[addr_index[tgt]:i + 1]
                c = collections.Counter()
                for _, tt in body:
                    tt = re.sub(r"^@!?U?P\d+\s+", "", tt)
                    op = tt.split()[0].split(".")[0]
                    c[op] += 1
                print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instr:", dict(c.most_common()))
