# Evidence run for the node-space band LU and the resident sum-of-regularisers solve (1 GPU, ≈ 4 min):
# parity tests of both gradient paths, then the timing tables quoted in DESIGN.md / BASELINE.md §6.
set -x
mkdir -p gpurun_out
L=gpurun_out/lu_evidence.log
( timeout 900 python -m pytest tests/test_gpu_sumregs.py tests/test_gpu_gradient.py -x -q 2>&1 | tail -5
  timeout 600 python tools/time_sumregs_pdps128.py 2>&1 | tail -4
  timeout 600 python tools/time_sumregs_grad.py 2>&1 | tail -10
  timeout 600 python tools/time_lu_cluster.py 2>&1 | tail -20
  timeout 600 python tools/time_tv_grad_reg.py 2>&1 | tail -10
  timeout 600 python tools/time_c5_reg.py 128 2>&1 | tail -3 ) > $L 2>&1
tail -80 $L
