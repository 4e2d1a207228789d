#!/bin/bash
# A/B of (library variant, environment) pairs on the bench line: tools/gpu_abv.sh "name ENV=1 ..." ...  (name = base or a build/variants suffix)
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg; n=$1; shift
  L=""; [ "$n" != base ] && L="VAW_LIBRARY=$PWD/build/variants/libvaw_$n.so"
  env $L "$@" timeout 300 python bench.py --no-e2e --no-parity --no-cpu-baseline --no-shim $ABL_ARGS 2>> gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$cfg', round(d['value']), 'frac', round(r['frac'],4), 'step', round(r['whole_step_frac'],4), 'sampler_ms', round(r['launch_ms']['avg'],4))" | tee -a gpurun_out/abv.log
done
