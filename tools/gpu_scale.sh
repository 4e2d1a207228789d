#!/bin/bash
# Scaling run on one box: N = 1, 2, 4, 8 ranks, one per GPU (the launch line the driver uses).
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  [ $n -gt $NG ] && break
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/scale_$n.json 2>> gpurun_out/scale.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/scale_$n.json 2>> gpurun_out/scale.err
  fi
  python -c "
import json; d=json.load(open('gpurun_out/scale_$n.json')); print('N=$n', round(d['value']), 'frames/s  step', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value']), d['clocks'])"
done
tail -2 gpurun_out/scale.err
