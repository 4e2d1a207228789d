#!/bin/bash
# round 2, multi-GPU pass: usage  gpurun --gpus N -- 'bash tools/gpu_r2_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/topo_$N.txt
numactl -H >> gpurun_out/topo_$N.txt 2>&1
# the in-library scheduler on distinct devices (GPU test) -- output kept
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "clip_scheduler" > gpurun_out/pytest_clip_${N}gpu.log 2>&1; tail -3 gpurun_out/pytest_clip_${N}gpu.log
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  if [ $n -eq 1 ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"; fi
  timeout 600 $LAUNCH bench.py --gpus $n --no-cpu-baseline --no-parity --no-shim > gpurun_out/scale_C3_$n.json 2>> gpurun_out/scale.err
  timeout 600 $LAUNCH bench.py --gpus $n --workload C4 --no-cpu-baseline --no-parity --no-shim > gpurun_out/scale_C4_$n.json 2>> gpurun_out/scale.err
  timeout 600 $LAUNCH bench.py --gpus $n --no-numa --no-cpu-baseline --no-parity --no-shim --steps 5 > gpurun_out/scale_C3_nonuma_$n.json 2>> gpurun_out/scale.err
  timeout 600 python tools/bench_clip.py --gpus $n --check > gpurun_out/clip_$n.json 2>> gpurun_out/scale.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/scale_C*_*.json"))+sorted(glob.glob("gpurun_out/clip_*.json")):
    try:
        d=json.load(open(f))
        e=d.get("e2e") or {}
        print(f.split("/")[-1], "value", round(d["value"]), "e2e", round(e.get("value",0)), "e2e_roofline", e.get("roofline_frac"), "copy_peak", e.get("copy_peak_gbs"), (d.get("details") or {}).get("numa"), d.get("device_numa_nodes"), d.get("first_shard_equals_single_context"), d.get("last_shard_equals_single_context"))
    except Exception as ex: print(f, "ERR", ex)
PY
tail -5 gpurun_out/scale.err
