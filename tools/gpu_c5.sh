python -m pytest tests/test_trbox.py -m gpu -x -q 2>&1 | tail -3
python tools/config5.py 148 5000
python tools/config5.py 296 5000
