#!/bin/bash
# round 2, final pass: GPU suite, smoke, bench (ours + reference), the other workloads, every mode, launch list and a
# full ncu capture of the step's kernels.  tools/make_profiles.py r02 turns gpurun_out/ into profiles/r02_*.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/env.log 2>&1
nproc >> gpurun_out/env.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
for c in C1 C2 C5; do timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-shim --workload $c --batch 32 > gpurun_out/bench_$c.json 2>> gpurun_out/bench.err; done
timeout 600 python bench.py --workload C4 --no-cpu-baseline --no-shim > gpurun_out/bench_C4.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --fused-bgr --no-cpu-baseline --no-e2e --no-shim > gpurun_out/bench_fused_bgr.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --fused-bgr --variant 2 --no-cpu-baseline --no-e2e --no-shim > gpurun_out/bench_fused_bgr_one_launch.json 2>> gpurun_out/bench.err
(PYTHONPATH=. timeout 300 python tools/bench_packed.py C3 32 bgr gray; PYTHONPATH=. timeout 300 python tools/bench_packed.py C1 32 bgr gray; PYTHONPATH=. timeout 300 python tools/bench_packed.py C5 16 bgr gray) > gpurun_out/packed.jsonl 2>> gpurun_out/bench.err
PYTHONPATH=. timeout 600 python tools/bench_modes.py > gpurun_out/modes.json 2>> gpurun_out/bench.err
PYTHONPATH=. timeout 300 python tools/bench_table_filters.py > gpurun_out/table_filters_final.txt 2>> gpurun_out/bench.err
./video_annotator_b200/host/vaw_demo --flow 40 3840 2160 0.5 > gpurun_out/flow_demo.log 2>&1; tail -1 gpurun_out/flow_demo.log
for b in 8 16 24 30; do ./video_annotator_b200/host/vaw_demo --bench 1500 $b 2>&1 | grep "^{" ; done > gpurun_out/shim_batch.log; cat gpurun_out/shim_batch.log
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --no-shim"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"warp_nv12|build_pieces" -s 6 -c 2 -f -o gpurun_out/prof_tiled $CMD > gpurun_out/ncu_full.log 2>&1
PYTHONPATH=. timeout 600 ncu --set full --clock-control none --import-source on -k regex:"packed_tile" -s 4 -c 1 -f -o gpurun_out/prof_bgr python tools/bench_packed.py C3 32 bgr > gpurun_out/ncu_bgr.log 2>&1
timeout 60 python tools/run_mode.py cubic > gpurun_out/run_cubic.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_nv12_quad -s 2 -c 1 -f -o gpurun_out/prof_cubic python tools/run_mode.py cubic > gpurun_out/ncu_cubic.log 2>&1
python - <<'PY'
import json
for f in ("bench", "bench_ref", "bench_C1", "bench_C2", "bench_C5", "bench_C4", "bench_fused_bgr"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json")); r = d.get("roofline", {})
        print(f, round(d["value"]), "frac", round(r.get("frac", 0), 4), "step", round(r.get("whole_step_frac", 0), 4), r.get("launch_ms", {}).get("avg"), "e2e", (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/bench.err
