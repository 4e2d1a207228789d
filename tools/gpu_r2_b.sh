#!/bin/bash
# round 2, pass B: ncu capture of the quadrant kernel + register-cap variants
mkdir -p gpurun_out
ab() { lib=$1; shift; VAW_LIBRARY=$PWD/build/variants/libvaw_$lib.so timeout 300 python bench.py --no-e2e --no-parity --no-cpu-baseline --no-shim --no-split-builder "$@" 2>> gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$lib $*', round(d['value']), 'frac', round(r['frac'],4), 'step', round(r['whole_step_frac'],4), 'sampler_ms', round(r['launch_ms']['avg'],4), d['details']['pieces_128x32']['tile_cap'])" | tee -a gpurun_out/ab.log; }
for v in c8 c7 c6; do ab $v; ab $v --workload C5 --batch 32; ab $v --workload C2 --batch 32; done
CMD="python bench.py --steps 2 --warmup 2 --inner 1 --no-e2e --no-parity --no-cpu-baseline --no-shim --no-split-builder"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"warp_nv12_quad|build_pieces" -s 6 -c 2 -f -o gpurun_out/prof_quad $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
