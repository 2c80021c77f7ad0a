timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "h9" 2>&1 | grep -E "AssertionError|assert |passed|failed" | head -8
