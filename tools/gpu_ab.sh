#!/bin/bash
# A/B of build-time options on the GPU box: tools/gpu_ab.sh "<defines A>" "<defines B>" ...
mkdir -p gpurun_out
for defs in "$@"; do
  VAW_DEFINES="$defs" python -m video_annotator_b200._build --force > /dev/null 2>&1
  python bench.py --no-cpu-baseline --no-e2e --steps 50 > gpurun_out/ab.json 2>> gpurun_out/ab.err
  python -c "
import json; d=json.load(open('gpurun_out/ab.json')); print('[$defs]', round(d['value']), round(d['roofline']['frac'],4), d['roofline']['launch_ms'], d['roofline']['other_kernels_ms'])" | tee -a gpurun_out/ab.log
done
python -m video_annotator_b200._build --force > /dev/null 2>&1
