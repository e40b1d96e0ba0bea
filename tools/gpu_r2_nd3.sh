set -x
mkdir -p gpurun_out
timeout 900 python tools/dbg_nd_256.py 128 5000 2>&1 | tail -40
timeout 900 python tools/time_grad_nd.py 2>&1 | tee gpurun_out/time_grad_nd.txt
( time timeout 1500 python -m pytest tests/test_gpu_gradient.py -m gpu -x -q 2>&1 | tail -15 ) 2>&1
