#!/bin/bash
# whole GPU suite, then TILED and PIPE on C3 / C5 / C2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for v in 3 4; do for c in C3 C5 C2; do timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 50 --variant $v --workload $c --batch $([ $c = C3 ] && echo 64 || echo 32) > gpurun_out/chk_v${v}_$c.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/chk_v${v}_$c.json')); print('variant $v $c', round(d['value']), round(d['roofline']['frac'],4), d['roofline']['launch_ms']['avg'], d['roofline']['other_kernels_ms'])"; done; done
