timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
