#!/bin/bash
# GPU pass: parity tests, bench for variants, ncu launch list + full capture of the top kernel.
mkdir -p gpurun_out
rm -f gpurun_out/parity.json
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --no-cpu-baseline --no-e2e --variant 1 > gpurun_out/bench_v1.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --no-cpu-baseline --no-e2e --workload C5 --batch 32 > gpurun_out/bench_c5.json 2>> gpurun_out/bench.err
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"warp_nv12|build_pieces" -s 6 -c 2 -f -o gpurun_out/prof_poly $CMD > gpurun_out/ncu_full.log 2>&1
tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json gpurun_out/bench_v1.json gpurun_out/bench_c5.json | cut -c1-1500; tail -3 gpurun_out/bench.err
