#!/bin/bash
# ablation builds (tools/build_variants.py) timed on the default workload: tools/gpu_abl.sh name1 name2 ...
mkdir -p gpurun_out
for n in "$@"; do
  L=""; [ "$n" != base ] && L="VAW_LIBRARY=$PWD/build/variants/libvaw_$n.so"
  env $L timeout 300 python bench.py --no-e2e --no-parity --no-cpu-baseline --no-shim $ABL_ARGS 2>> gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$n', round(d['value']), 'frac', round(r['frac'],4), 'sampler_ms', round(r['launch_ms']['avg'],4))" | tee -a gpurun_out/abl.log
done
