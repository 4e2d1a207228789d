#!/bin/bash
# C1 (1080p) with another rows-per-piece rule: tools/gpu_c1.sh "<defines>" ...
mkdir -p gpurun_out
for defs in "$@"; do
  VAW_DEFINES="$defs" python -m video_annotator_b200._build --force > /dev/null 2>&1
  timeout 600 python -m pytest tests -m gpu -q -x -k "C1 or windows or small_golden or classification" > gpurun_out/pytest_c1.log 2>&1; echo "[$defs] pytest rc=$?" | tee -a gpurun_out/c1.log; tail -3 gpurun_out/pytest_c1.log
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 50 --workload C1 --batch 32 > gpurun_out/c1.json 2>> gpurun_out/bench.err
  python -c "
import json; d=json.load(open('gpurun_out/c1.json')); print('[$defs] C1', round(d['value']), round(d['roofline']['frac'],4), d['roofline']['launch_ms']['avg'], d['roofline']['other_kernels_ms'], d['config']['pieces_128x32'], d['config']['variant_resolved'])" | tee -a gpurun_out/c1.log
done
python -m video_annotator_b200._build --force > /dev/null 2>&1
