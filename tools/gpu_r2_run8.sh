# Round-2 run 8 (1 GPU): full parity suite after the stats / tolerance changes of the banded solvers and the depth rule;
# timings at config 5's per-GPU shape
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/gputests.txt 2>&1 ) 2>&1 | tail -4
tail -30 gpurun_out/gputests.txt
timeout 600 python tools/time_grad_nd.py 128x256 2>&1 | tee gpurun_out/time_grad_nd_c5.txt
