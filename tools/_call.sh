timeout 900 python -m pytest tests/test_gpu_results.py tests/test_gpu_configs.py tests/test_cpp_wrapper.py -x -q -m gpu 2>&1 | tail -8
