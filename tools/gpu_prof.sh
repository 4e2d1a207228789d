#!/bin/bash
# ncu full capture (with source) of one sampler launch of the default bench workload: tools/gpu_prof.sh NAME [ENV=.. bench args]
mkdir -p gpurun_out
NAME=$1; shift
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --no-shim"
env $PROF_ENV timeout 300 $CMD "$@" > gpurun_out/plain_$NAME.log 2>&1 &&
env $PROF_ENV timeout 900 ncu --set full --clock-control none --import-source on -k regex:"warp_nv12" -s 3 -c 1 -f -o gpurun_out/prof_$NAME $CMD "$@" > gpurun_out/ncu_$NAME.log 2>&1
cut -c1-200 gpurun_out/plain_$NAME.log; tail -2 gpurun_out/ncu_$NAME.log
