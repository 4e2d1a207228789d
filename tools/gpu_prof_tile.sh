#!/bin/bash
# ncu full capture of one warp_nv12 launch of the default bench configuration.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline ${BENCH_ARGS}"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"warp_nv12" -s 3 -c 1 -f -o gpurun_out/prof_exp $CMD > gpurun_out/ncu_full.log 2>&1
cut -c1-200 gpurun_out/plain2.log
