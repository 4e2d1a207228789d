#!/bin/bash
# A/B of sampler settings through environment switches: tools/gpu_ab.sh "ENV=1 ..." "ENV=2 ..." ; each on C3 (and C5/C2/C1 when AB_ALL=1)
mkdir -p gpurun_out
ab() { env $1 timeout 300 python bench.py --no-e2e --no-parity --no-cpu-baseline --no-shim ${@:2} 2>> gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$*', round(d['value']), 'frac', round(r['frac'],4), 'step', round(r['whole_step_frac'],4), 'sampler_ms', round(r['launch_ms']['avg'],4), 'builder', [round(v,4) for v in r['other_kernels_ms'].values()])" | tee -a gpurun_out/ab.log; }
for cfg in "$@"; do
  ab "$cfg"
  if [ -n "$AB_ALL" ]; then for c in C5 C2 C1; do ab "$cfg" --workload $c --batch 32; done; fi
done
if [ -n "$AB_TEST" ]; then env $AB_TEST timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_ab.log 2>&1; tail -3 gpurun_out/pytest_ab.log; fi
