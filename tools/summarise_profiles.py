"""Turns the raw ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1f"
go = os.path.join(ROOT, "gpurun_out")
pr = os.path.join(ROOT, "profiles")

rows = [r for r in csv.reader(open(os.path.join(go, f"launches_{tag}.csv"))) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try:
        agg[r[ki][:80]].append(float(r[vi].replace(",", "")))
    except ValueError:
        pass
tot = sum(sum(v) for v in agg.values())
lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-newton",
         "# per-launch times are cold-cache and serialised: compare SHARES (setup kernels run once, step kernels 8x+)",
         "kernel,launches,mean_us,share_pct"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"\"{k}\",{len(v)},{sum(v) / len(v) / 1e3:.1f},{sum(v) / tot * 100:.1f}")
open(os.path.join(pr, "r1_launch_list_summary.csv"), "w").write("\n".join(lines) + "\n")

out = subprocess.run(["ncu", "-i", os.path.join(go, f"prof_{tag}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts", "smsp__thread_inst_executed_per_inst_executed", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__throughput.avg.pct", "registers_per_thread", "sm__warps_active.avg.pct", "pipe_fp64", "occupancy_limit",
        "l1tex__t_sectors_pipe_lsu_mem_global_op", "l1tex__t_requests_pipe_lsu_mem_global_op",
        "bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "l1tex__throughput.avg.pct",
        "lts__throughput.avg.pct", "issue_stalled_long_scoreboard_per", "issue_stalled_short_scoreboard_per", "issue_stalled_wait_per",
        "issue_stalled_math_pipe", "issue_stalled_barrier_per", "issue_stalled_mio", "lts__t_sectors_srcunit_tex_op"]
keep = [i for i, h in enumerate(hdr) if any(k in h for k in keys)]
traffic = {"source": f"profiles/r1_ncu_full_elem_gather_v5.csv (ncu --set full, C2 workload, one launch each): dram__bytes_read.sum + dram__bytes_write.sum"}
with open(os.path.join(pr, "r1_ncu_full_elem_gather_v5.csv"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:'elem_q1|gather_pull' -s 6 -c 2 python bench.py --steps 3 --warmup 3 --no-cpu --no-newton (C2 workload)\n")
    for r in rows[2:]:
        for i in keep:
            f.write(f"{hdr[i]},{rows[1][i]},{r[i]}\n")
        f.write("\n")
        name = "gather_pull_kernel" if "gather" in r[hdr.index("Kernel Name")] else "elem_q1_kernel"
        rd, wr = float(r[hdr.index("dram__bytes_read.sum")]), float(r[hdr.index("dram__bytes_write.sum")])
        unit = rows[1][hdr.index("dram__bytes_read.sum")]
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1e6)
        traffic[name] = int((rd + wr) * scale)
        print(name, r[hdr.index("gpu__time_duration.sum")], "us  dram R/W", rd, wr, unit)
json.dump(traffic, open(os.path.join(pr, "r1_traffic.json"), "w"), indent=1)
for f in (f"bench_{tag}_n1.json",):
    src = os.path.join(go, f)
    if os.path.exists(src):
        open(os.path.join(pr, "r1_bench_n1.json"), "w").write(open(src).read())
print(open(os.path.join(pr, "r1_launch_list_summary.csv")).read()[:900])
