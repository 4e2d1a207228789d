#!/bin/bash
# Variant PIPE (ring pipeline) A/B over build options: tools/gpu_pipe2.sh "<defines A>" "<defines B>" ...
mkdir -p gpurun_out
for defs in "$@"; do
  VAW_DEFINES="$defs" python -m video_annotator_b200._build --force > /dev/null 2>&1
  timeout 600 python -m pytest tests -m gpu -q -x -k "variants_produce or (pixels_bit_exact and C3)" > gpurun_out/pytest_pipe.log 2>&1; echo "[$defs] pytest rc=$?" | tee -a gpurun_out/pipe.log
  tail -2 gpurun_out/pytest_pipe.log
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 50 --variant 4 > gpurun_out/pipe_v4.json 2>> gpurun_out/pipe.err
  python -c "
import json; d=json.load(open('gpurun_out/pipe_v4.json')); print('[$defs] variant 4', round(d['value']), round(d['roofline']['frac'],4), d['roofline']['launch_ms'], d['roofline']['other_kernels_ms'])" | tee -a gpurun_out/pipe.log
done
python -m video_annotator_b200._build --force > /dev/null 2>&1
tail -3 gpurun_out/pipe.err
