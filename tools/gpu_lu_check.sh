set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_sumregs.py tests/test_gpu_gradient.py -x -q 2>&1 | tail -4
  timeout 300 python tools/_probe2.py 2>&1 | tail -12 ) > gpurun_out/probe2.log 2>&1
tail -30 gpurun_out/probe2.log
