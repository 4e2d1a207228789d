#!/usr/bin/env python
"""Distribution of the shared-memory tile sizes the pieces of a workload need (GPU): python tools/tile_hist.py C3 [frame]"""
import sys, numpy as np
import video_annotator_b200 as V
from video_annotator_b200 import configs
for name in sys.argv[1:]:
    w = configs.workload(name)
    rots = w.rotations(64, first=0, total=64)
    ctx = V.WarpContext(w.input_camera, w.output_camera, out_size=w.out_size)
    for fi in (0, 32, 63):
        t = ctx.piece_tiles(rots[fi])
        t = t[t[:, 0] != 0] if False else t
        n = ((w.out_size[0] + 127) // 128) * ((w.out_size[1] + 31) // 32)
        t = t[:n]
        need = t[:, 1].astype(np.int64)
        s = need[(need > 0) & (need < 0x7fffffff)]
        pl = t[:, 2][(need > 0) & (need < 0x7fffffff)]
        nr = (t[:, 3] & 0xffff)[(need > 0) & (need < 0x7fffffff)]
        cr = (t[:, 3] >> 16)[(need > 0) & (need < 0x7fffffff)]
        print(name, "frame", fi, "pieces", n, "staged", len(s), "bytes pct 10/50/90/99/max", [int(np.percentile(s, q)) for q in (10, 50, 90, 99, 100)],
              "pl", dict(zip(*np.unique(pl, return_counts=True))), "luma rows 50/max", int(np.median(nr)), int(nr.max()), "chroma", int(np.median(cr)), int(cr.max()))
        print("   hist KB:", np.histogram(s / 1024, bins=[0, 8, 12, 16, 20, 24, 28, 32, 48, 64, 200])[0])
    ctx.close()
