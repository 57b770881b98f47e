timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s -k "small" 2>&1 | tail -12
