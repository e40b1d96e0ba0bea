set -x
mkdir -p gpurun_out
L=gpurun_out/srr_check.log
( timeout 900 python -m pytest tests/test_gpu_sumregs.py -x -q 2>&1 | tail -25
  timeout 600 python tools/time_sumregs_pdps128.py 2>&1 | tail -20 ) > $L 2>&1
tail -60 $L
