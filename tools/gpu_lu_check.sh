set -x
mkdir -p gpurun_out
L=gpurun_out/lu_cluster.log
( timeout 900 python -m pytest tests/test_gpu_sumregs.py -x -q 2>&1 | tail -5
  timeout 900 python tools/time_lu_cluster.py 2>&1 | tail -30 ) > $L 2>&1
tail -60 $L
