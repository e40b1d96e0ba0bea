set -x
mkdir -p gpurun_out
timeout 900 python tools/dbg_nd_256.py 128 5000 2>&1 | tail -30
timeout 300 python tools/prof_grad_nd.py 128 148 1000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:nd_ --csv --log-file gpurun_out/nd_launches_148x128.csv python tools/prof_grad_nd.py 128 148 1000 > gpurun_out/ncu_nd.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:nd_ --csv --log-file gpurun_out/nd_launches_1x128.csv python tools/prof_grad_nd.py 128 1 5000 >> gpurun_out/ncu_nd.log 2>&1
tail -5 gpurun_out/ncu_nd.log
