#!/bin/bash
# Full GPU pass: tests, smoke, bench (ours + reference), ncu launch list + full capture.
mkdir -p gpurun_out
rm -f gpurun_out/parity.json
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
for c in C1 C2 C5; do timeout 600 python bench.py --no-cpu-baseline --no-e2e --workload $c --batch 32 > gpurun_out/bench_$c.json 2>> gpurun_out/bench.err; done
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"warp_nv12|build_pieces" -s 6 -c 2 -f -o gpurun_out/prof_tiled $CMD > gpurun_out/ncu_full.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-2500 gpurun_out/bench.json; cut -c1-400 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench.err
for c in C1 C2 C5; do python -c "
import json; d=json.load(open('gpurun_out/bench_$c.json')); print('$c', round(d['value']), round(d['roofline']['frac'],4), d['roofline']['launch_ms']['avg'], d['roofline']['other_kernels_ms'])"; done
# the other NV12 variants on the same workloads (A/B lines for DESIGN.md), and ncu captures of their kernels
for v in 4 5; do for c in C3 C5 C2; do timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 50 --variant $v --workload $c --batch $([ $c = C3 ] && echo 64 || echo 32) > gpurun_out/bench_v${v}_$c.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v${v}_$c.json')); print('variant $v $c', round(d['value']), round(d['roofline']['frac'],4), d['roofline']['launch_ms']['avg'], d['roofline']['other_kernels_ms'])"; done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"warp_nv12_pipe" -s 3 -c 1 -f -o gpurun_out/prof_pipe $CMD --variant 4 > gpurun_out/ncu_pipe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"warp_nv12_tex" -s 3 -c 1 -f -o gpurun_out/prof_tex $CMD --variant 5 > gpurun_out/ncu_tex.log 2>&1
