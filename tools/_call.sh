NCU="ncu --clock-control none"
B="python bench.py --steps 5 --warmup 3 --no-cpu --no-c5 --no-mirror"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_r2.csv $B --no-newton > gpurun_out/launches_r2.log 2>&1
C3_N=24 $NCU --set full -k regex:'elem_q2' -s 2 -c 1 -f -o gpurun_out/prof_r2_q2 python tools/config_times.py C3 > /dev/null 2>&1
C4_N=48 $NCU --set full -k regex:'elem_easdg' -s 2 -c 1 -f -o gpurun_out/prof_r2_easdg python tools/config_times.py C4dg > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
