set -x
timeout 600 python tools/time_resident.py
timeout 900 python -m pytest tests/test_gpu_pdps.py -m gpu -x -q 2>&1 | tail -5
