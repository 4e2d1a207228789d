#!/bin/bash
# GPU pass: parity tests, bench for variants (no ncu).
mkdir -p gpurun_out
rm -f gpurun_out/parity.json
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for v in 3 2; do
timeout 600 python bench.py --no-cpu-baseline --no-e2e --variant $v > gpurun_out/bench_v$v.json 2>> gpurun_out/bench.err
done
tail -25 gpurun_out/pytest_gpu.log; for v in 3 2; do python -c "
import json,sys; d=json.load(open('gpurun_out/bench_v$v.json')); print('variant $v', d['value'], d['roofline']['frac'], d['roofline']['launch_ms'])"; done; tail -3 gpurun_out/bench.err
