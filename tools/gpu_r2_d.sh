#!/bin/bash
# round 2, pass D: full GPU suite (incl. certificate sweep, fused BGR, shim batching) + e2e pipeline shape at N = 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/parity.json"))
print(json.dumps(d.get("certificate_sweep"), indent=1))
PY
for st in 4 6 8; do for mb in 32 64 128; do
timeout 300 python bench.py --no-parity --no-cpu-baseline --no-shim --steps 3 --host-stages $st --host-chunk-mb $mb 2>> gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('stages $st chunk $mb MB: e2e', round(e['value']), 'GB/s/dir', round(e['gbs_per_direction_per_gpu'],2), 'peak', round(e['copy_peak_gbs'],2), 'frac', round(e['roofline_frac'],3))" | tee -a gpurun_out/e2e_shape.log
done; done
