# Round-2 run 9 (1 GPU): kernel B with the st.async / mbarrier halo exchange against the cluster-barrier version
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pdps.py tests/test_gpu_sweep.py -m gpu -q 2>&1 | tail -5
timeout 600 python tools/time_resident.py 2>&1 | tee gpurun_out/time_resident_async.txt
BPLTV_RESIDENT_ASYNC=0 timeout 600 python tools/time_resident.py 2>&1 | tee gpurun_out/time_resident_sync.txt
