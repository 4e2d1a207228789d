#!/bin/bash
# Tests, then bench per variant.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for v in 4 3; do for c in C3 C5; do timeout 600 python bench.py --no-cpu-baseline --no-e2e --variant $v --workload $c --batch $([ $c = C3 ] && echo 64 || echo 32) > gpurun_out/bench_v${v}_$c.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v${v}_$c.json')); print('variant $v $c', round(d['value']), 'kernel frac', round(d['roofline']['frac'],4), 'step frac', round(d['roofline']['whole_step_frac'],4), d['roofline']['launch_ms']['avg'], d['roofline']['other_kernels_ms'])"; done; done
tail -3 gpurun_out/bench.err
