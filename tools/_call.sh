python bench.py --steps 200 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 200 gpurun_out/bench_ref.json
python tools/config_times.py C1 C3 C4 C4b C4dg 2>/dev/null > gpurun_out/config_times_r2.jsonl; wc -l gpurun_out/config_times_r2.jsonl
python __graft_entry__.py --smoke 2>&1 | tail -2
