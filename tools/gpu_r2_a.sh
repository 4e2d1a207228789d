#!/bin/bash
# round 2, pass A: the GPU suite after the oracle/_ref + split-builder + quadrant-kernel + bench changes; A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/env.log 2>&1
nproc >> gpurun_out/env.log; lscpu | grep -E "Model name|Socket|NUMA" >> gpurun_out/env.log
ab() { timeout 300 python bench.py --no-e2e --no-parity --no-cpu-baseline --no-shim "$@" 2>> gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$*', round(d['value']), 'frac', round(r['frac'],4), 'step', round(r['whole_step_frac'],4), 'sampler_ms', round(r['launch_ms']['avg'],4), 'builder', [round(v,4) for v in r['other_kernels_ms'].values()], d['details']['pieces_128x32'])" | tee -a gpurun_out/ab.log; }
ab --tile-kernel 2
ab --tile-kernel 1
ab --tile-kernel 2 --no-split-builder
ab --tile-kernel 1 --no-split-builder
ab --tile-kernel 2 --variant 4
for c in C5 C2 C1; do ab --workload $c --batch 32 --tile-kernel 2 --variant 3; ab --workload $c --batch 32 --tile-kernel 1 --variant 3; ab --workload $c --batch 32 --variant 4; done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2>> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload C4 --no-cpu-baseline --no-shim > gpurun_out/bench_C4.json 2>> gpurun_out/bench.err; echo "C4 rc=$?"
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
python - <<'PY'
import json
for f in ("bench","bench_C4","bench_ref"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        r=d.get("roofline",{})
        print(f, round(d["value"]), "ms/step", round(d["ms_per_step"],3), "inner", d.get("launches_per_step"), "frac", round(r.get("frac",0),4), "step_frac", round(r.get("whole_step_frac",0),4), r.get("launch_ms"), r.get("other_kernels_ms"))
        if d.get("e2e"): print("  e2e", round(d["e2e"]["value"]), {k:v for k,v in d["e2e"].items() if k in ("gbs_per_direction_per_gpu","copy_peak","roofline_frac")})
        if d.get("parity"): print("  parity", d["parity"]["coord_max_err_px"], d["parity"]["same_map"]["max_lsb"], d["parity"]["vs_reference_path"]["gt1"], d["parity"]["coordinate_oracle"])
        if d.get("cpu_baseline"): print("  cpu", d["cpu_baseline"])
        print("  details", d.get("details"))
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/bench.err
