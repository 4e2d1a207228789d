#!/bin/bash
# Variant PIPE (ring pipeline): parity tests that involve it, then a bench A/B against TILED.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "variant or pixels_bit_exact" > gpurun_out/pytest_pipe.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pipe.log
tail -5 gpurun_out/pytest_pipe.log
for v in 3 4; do
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 50 --variant $v > gpurun_out/pipe_v$v.json 2>> gpurun_out/pipe.err
  python -c "
import json; d=json.load(open('gpurun_out/pipe_v$v.json')); print('variant $v', round(d['value']), round(d['roofline']['frac'],4), d['roofline']['launch_ms'], d['roofline']['other_kernels_ms'])" | tee -a gpurun_out/pipe.log
done
for c in C5 C2; do timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 30 --variant 4 --workload $c --batch 32 > gpurun_out/pipe_$c.json 2>> gpurun_out/pipe.err; python -c "
import json; d=json.load(open('gpurun_out/pipe_$c.json')); print('$c variant 4', round(d['value']), round(d['roofline']['frac'],4), d['roofline']['launch_ms'])" | tee -a gpurun_out/pipe.log; done
tail -3 gpurun_out/pipe.err
