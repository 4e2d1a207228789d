#!/bin/bash
# Instrumented build: every staged tap address range-checked, whole GPU suite, then restore the normal build.
mkdir -p gpurun_out
VAW_DEFINES="VAW_BOUNDS_CHECK=1" python -m video_annotator_b200._build --force > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_boundscheck.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_boundscheck.log
tail -4 gpurun_out/pytest_boundscheck.log
python -c "
import json; print('oob_taps', json.load(open('gpurun_out/parity.json')).get('oob_taps'))"
python -m video_annotator_b200._build --force > /dev/null 2>&1
