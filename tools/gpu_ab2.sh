#!/bin/bash
# A/B of build-time options for one variant: tools/gpu_ab2.sh <variant> "<defines A>" "<defines B>" ...
mkdir -p gpurun_out
v=$1; shift
for defs in "$@"; do
  VAW_DEFINES="$defs" python -m video_annotator_b200._build --force > /dev/null 2>&1
  timeout 300 python -m pytest tests -m gpu -q -x -k "variants_produce or (pixels_bit_exact and C3)" > gpurun_out/pytest_ab.log 2>&1; echo "[$defs] pytest rc=$?" | tee -a gpurun_out/ab2.log
  timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 50 --variant $v > gpurun_out/ab2.json 2>> gpurun_out/ab2.err
  python -c "
import json; d=json.load(open('gpurun_out/ab2.json')); print('[$defs] variant $v', round(d['value']), round(d['roofline']['frac'],4), d['roofline']['launch_ms'], d['roofline']['other_kernels_ms'])" | tee -a gpurun_out/ab2.log
done
python -m video_annotator_b200._build --force > /dev/null 2>&1
tail -3 gpurun_out/ab2.err
