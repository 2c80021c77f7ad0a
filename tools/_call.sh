timeout 600 python -m pytest tests/test_cpp_wrapper.py -x -q 2>&1 | tail -12
