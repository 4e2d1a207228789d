#!/bin/bash
# PIPE with other group counts on C3 / C5 / C2: tools/gpu_pipe_groups.sh "<defines>" ...
mkdir -p gpurun_out
for defs in "$@"; do
  VAW_DEFINES="$defs" python -m video_annotator_b200._build --force > /dev/null 2>&1
  for c in C3 C5 C2; do timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 50 --variant 4 --workload $c --batch $([ $c = C3 ] && echo 64 || echo 32) > gpurun_out/pg.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/pg.json')); print('[$defs] $c', round(d['value']), round(d['roofline']['frac'],4), d['roofline']['launch_ms']['avg'])" | tee -a gpurun_out/pipe_groups.log; done
done
python -m video_annotator_b200._build --force > /dev/null 2>&1
