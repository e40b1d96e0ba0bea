set -x
mkdir -p gpurun_out
L=gpurun_out/wide_solve.log
( timeout 900 python -m pytest tests/test_gpu_sumregs.py tests/test_gpu_gradient.py -x -q 2>&1 | tail -5
  timeout 900 python tools/time_sumregs_grad.py 2>&1 | tail -20 ) > $L 2>&1
tail -60 $L
