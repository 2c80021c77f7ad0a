timeout 1200 python -m pytest tests/test_gpu_eas.py -x -q -m gpu 2>&1 | tail -15
