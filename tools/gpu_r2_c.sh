#!/bin/bash
# round 2, pass C: GPU suite incl. fused NV12->BGR, shim bench, quad <6>/<8> instantiations
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
ab() { timeout 300 python bench.py --no-e2e --no-parity --no-cpu-baseline --no-shim "$@" 2>> gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$*', round(d['value']), 'frac', round(r['frac'],4), 'step', round(r['whole_step_frac'],4), 'sampler_ms', round(r['launch_ms']['avg'],4), 'builder', [round(v,4) for v in r['other_kernels_ms'].values()], d['details']['pieces_128x32']['tile_cap'])" | tee -a gpurun_out/ab.log; }
ab
ab --no-split-builder
ab --workload C5 --batch 32
ab --workload C2 --batch 32
ab --workload C1 --batch 32
ab --fused-bgr
ab --fused-bgr --workload C1 --batch 32
./video_annotator_b200/host/vaw_demo --bench 1500 16 | tee gpurun_out/shim_bench.json
./video_annotator_b200/host/vaw_demo --bench 1500 32 | tee -a gpurun_out/shim_bench.json
tail -3 gpurun_out/bench.err
