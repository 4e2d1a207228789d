#!/bin/bash
# GPU pass: parity tests, TILED bench, ncu full capture of the TILED kernel + builder.
mkdir -p gpurun_out
rm -f gpurun_out/parity.json
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --variant 3"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"warp_nv12|build_pieces" -s 6 -c 2 -f -o gpurun_out/prof_tiled $CMD > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/plain2.log | cut -c1-300
