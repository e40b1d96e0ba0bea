#!/bin/bash
# kernel C bring-up: parity tests, then timing sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pdps.py -m gpu -x -q -k "tblock or error_behaviour" 2>&1 | tail -15 | tee gpurun_out/tblock_tests.log
timeout 600 python tools/time_pdps.py 240 2>&1 | tee gpurun_out/tblock_time64.log
BPLTV_PREC=32 timeout 600 python tools/time_pdps.py 240 2>&1 | tee gpurun_out/tblock_time32.log
