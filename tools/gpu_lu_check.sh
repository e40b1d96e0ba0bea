set -x
mkdir -p gpurun_out
L=gpurun_out/tv_lu.log
( BPLTV_GRAD_REG_LU=1 timeout 900 python -m pytest tests/test_gpu_gradient.py tests/test_trbox.py -x -q 2>&1 | tail -5
  timeout 900 python tools/time_tv_grad_reg.py 2>&1 | tail -20 ) > $L 2>&1
tail -60 $L
