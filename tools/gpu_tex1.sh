#!/bin/bash
# First look at variant TEX: differences against TILED and kernel times, at two register budgets.
mkdir -p gpurun_out
for defs in "VAW_TEX_CTAS=8" "VAW_TEX_CTAS=6"; do
  VAW_DEFINES="$defs" python -m video_annotator_b200._build --force > /dev/null 2>&1
  for w in 0 1; do
    echo "[$defs] white=$w" | tee -a gpurun_out/tex_probe.log
    timeout 300 python tools/tex_probe.py C3 64 $w 2>&1 | tail -3 | tee -a gpurun_out/tex_probe.log
  done
done
timeout 300 python tools/tex_probe.py C5 32 1 2>&1 | tail -3 | tee -a gpurun_out/tex_probe.log
timeout 300 python tools/tex_probe.py C1 32 1 2>&1 | tail -3 | tee -a gpurun_out/tex_probe.log
python -m video_annotator_b200._build --force > /dev/null 2>&1
