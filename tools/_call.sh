timeout 900 python -m pytest tests/test_gpu_trust_region.py tests/test_cpp_wrapper.py -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 50 --warmup 5 --no-c5 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['trust_region_inner'], d['newton_step'])"
