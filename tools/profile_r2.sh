#!/bin/bash
# Round-2 profile captures (run on the GPU box through gpurun; outputs land in gpurun_out/, summaries are made from them
# by tools/summarise_profiles.py and committed under profiles/).  Never a bench number: everything here runs under ncu.
set -x
PART=${1:-all}   # gpurun brings back at most 64 MiB: run "a" and "b" in separate calls
NCU="ncu --clock-control none"
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-c5 --no-c4"
[ "$PART" = b ] || $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_r2.csv $B --no-newton > gpurun_out/launches_r2.log 2>&1
[ "$PART" = b ] || $NCU --set full --import-source on -k regex:'elem_h8_mma|gather_pull' -s 6 -c 2 -o gpurun_out/prof_r2_c2 $B --no-newton > /dev/null 2>&1
[ "$PART" = b ] || IKB_ELEM=fma $NCU --set full --import-source on -k regex:'elem_q1' -s 3 -c 1 -o gpurun_out/prof_r2_c2_fma $B --no-newton > /dev/null 2>&1
[ "$PART" = a ] || $NCU --set full -k regex:'spmv_node_dot' -s 40 -c 1 -o gpurun_out/prof_r2_spmv $B > /dev/null 2>&1
[ "$PART" = a ] || C4_N=48 $NCU --set full -k regex:'elem_eas' -s 2 -c 1 -o gpurun_out/prof_r2_eas python tools/config_times.py C4 > /dev/null 2>&1
[ "$PART" = a ] || C3_N=24 $NCU --set full -k regex:'elem_q2' -s 2 -c 1 -o gpurun_out/prof_r2_q2 python tools/config_times.py C3 > /dev/null 2>&1
[ "$PART" = a ] || C4_N=48 $NCU --set full -k regex:'elem_easdg' -s 2 -c 1 -o gpurun_out/prof_r2_easdg python tools/config_times.py C4dg > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
