#!/usr/bin/env python
"""Print SASS of a kernel between two addresses: sass_range.py lib kernel-substring lo hi"""
import re, subprocess, sys
lib, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", out)[1:]:
    if pat not in f.split("\n", 1)[0]:
        continue
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and lo <= int(m.group(1), 16) <= hi:
            print(m.group(1), m.group(2).strip())
