python tools/config_times.py C1 C3 C4 C4b C4dg 2>/dev/null | tee gpurun_out/config_times_r2.jsonl
