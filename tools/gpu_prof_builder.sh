#!/bin/bash
# ncu full capture of one build_pieces launch of the default bench configuration.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"build_pieces" -s 3 -c 1 -f -o gpurun_out/prof_builder $CMD > gpurun_out/ncu_builder.log 2>&1
