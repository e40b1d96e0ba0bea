set -x
mkdir -p gpurun_out
timeout 900 python tools/time_grad_nd.py 2>&1 | tee gpurun_out/time_grad_nd.txt
( time timeout 1500 python -m pytest tests/test_gpu_gradient.py -m gpu -x -q -k "binary128 or fp32 or nested or reports or 256" 2>&1 | tail -15 ) 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:nd_ --csv --log-file gpurun_out/nd_launches_148x128.csv python tools/prof_grad_nd.py 128 148 1000 > gpurun_out/ncu_nd.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:nd_ --csv --log-file gpurun_out/nd_launches_1x128.csv python tools/prof_grad_nd.py 128 1 5000 >> gpurun_out/ncu_nd.log 2>&1
tail -3 gpurun_out/ncu_nd.log
