#!/bin/bash
# End-to-end (host-buffer) pipeline shape at N ranks: usage  gpurun --gpus N -- 'bash tools/gpu_e2e_sweep.sh N'
# Each line: stages, chunk MiB, e2e frames/s, GB/s per direction per GPU, the plain-copy peak measured in the same run, their ratio.
N=${1:-8}
mkdir -p gpurun_out
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
[ "$N" = 1 ] && LAUNCH=python
for cfg in ${SWEEP:-"4 32" "8 32" "4 128" "8 128"}; do
  set -- $cfg
  timeout 300 $LAUNCH bench.py --gpus $N --no-cpu-baseline --no-parity --no-shim --steps 5 --host-stages $1 --host-chunk-mb $2 2>> gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('stages $1 chunk_mb $2 n $N e2e', round(e['value']), 'gbs/dir/gpu', round(e['gbs_per_direction_per_gpu'],2), 'copy_peak', round(e['copy_peak_gbs'],2), 'frac', round(e['roofline_frac'],3))" | tee -a gpurun_out/e2e_sweep_$N.log
done
tail -3 gpurun_out/sweep.err
