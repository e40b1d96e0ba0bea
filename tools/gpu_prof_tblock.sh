#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-2}
export BPLTV_TBLOCK_T=$T
python tools/profile_case.py tblock 12 > gpurun_out/plain_tblock.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pdps_tblock -s 2 -c 2 -f -o gpurun_out/prof_tblock_T$T python tools/profile_case.py tblock 12 > gpurun_out/ncu_tblock.log 2>&1
tail -n 3 gpurun_out/plain_tblock.log gpurun_out/ncu_tblock.log
