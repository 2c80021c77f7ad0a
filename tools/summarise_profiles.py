"""Turns the raw ncu outputs brought back in gpurun_out/ (tools/profile_r2.sh) into the small text summaries committed
under profiles/: the launch list with per-kernel shares, one CSV of the judged counters per `ncu --set full` capture, and
the DRAM bytes per launch of the two step kernels (read by bench.py for roofline.traffic)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
go = os.path.join(ROOT, "gpurun_out")
pr = os.path.join(ROOT, "profiles")

KEYS = ["Kernel Name", "gpu__time_duration.sum", "l1tex__data_pipe_lsu_wavefronts", "smsp__thread_inst_executed_per_inst_executed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct", "sm__throughput.avg.pct",
        "registers_per_thread", "sm__warps_active.avg.pct", "pipe_fp64", "pipe_tensor", "occupancy_limit",
        "l1tex__t_sectors_pipe_lsu_mem_global_op", "l1tex__t_requests_pipe_lsu_mem_global_op",
        "bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
        "l1tex__throughput.avg.pct", "lts__throughput.avg.pct", "lts__t_sector_hit_rate", "issue_stalled_long_scoreboard_per",
        "issue_stalled_short_scoreboard_per", "issue_stalled_wait_per", "issue_stalled_math_pipe", "issue_stalled_barrier_per",
        "issue_stalled_mio", "issue_stalled_lg_throttle", "lts__t_sectors_srcunit_tex_op", "sass__inst_executed_shared",
        "launch__grid_size", "launch__block_size"]


def launch_list():
    src = os.path.join(go, f"launches_{tag}.csv")
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            agg[r[ki][:90]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-c5 --no-newton",
             "# per-launch times are cold-cache and serialised: compare SHARES (setup kernels run once, step kernels 8x+)",
             "kernel,launches,mean_us,share_pct"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"\"{k}\",{len(v)},{sum(v) / len(v) / 1e3:.1f},{sum(v) / tot * 100:.1f}")
    open(os.path.join(pr, f"{tag}_launch_list_summary.csv"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:14]))


def full_capture(rep, out_name, command, traffic=None):
    src = os.path.join(go, rep)
    if not os.path.exists(src):
        print("missing", rep)
        return
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    keep = [i for i, h in enumerate(hdr) if any(k in h for k in KEYS)]
    with open(os.path.join(pr, out_name), "w") as f:
        f.write(f"# {command}\n")
        for r in rows[2:]:
            for i in keep:
                f.write(f"{hdr[i]},{rows[1][i]},{r[i]}\n")
            f.write("\n")
            name = r[hdr.index("Kernel Name")]
            rd, wr = float(r[hdr.index("dram__bytes_read.sum")]), float(r[hdr.index("dram__bytes_write.sum")])
            unit = rows[1][hdr.index("dram__bytes_read.sum")]
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1e6)
            short = name.split("<")[0].replace("void ", "").replace("ikb::", "")
            print(f"{out_name}: {short} {r[hdr.index('gpu__time_duration.sum')]} {rows[1][hdr.index('gpu__time_duration.sum')]}"
                  f"  dram R/W {rd} {wr} {unit}")
            if traffic is not None:
                traffic[short] = int((rd + wr) * scale)


launch_list()
traffic = {"source": f"profiles/{tag}_ncu_full_c2_elem_gather.csv (ncu --set full, C2 workload, one launch each): "
                     "dram__bytes_read.sum + dram__bytes_write.sum"}
B = "python bench.py --steps 5 --warmup 3 --no-cpu --no-c5"
full_capture(f"prof_{tag}_c2.ncu-rep", f"{tag}_ncu_full_c2_elem_gather.csv",
             f"ncu --set full --clock-control none --import-source on -k regex:'elem_h8_mma|gather_pull' -s 6 -c 2 {B} --no-newton (C2)",
             traffic)
full_capture(f"prof_{tag}_c2_fma.ncu-rep", f"{tag}_ncu_full_c2_elem_fma.csv",
             f"IKB_ELEM=fma ncu --set full --clock-control none --import-source on -k regex:'elem_q1' -s 3 -c 1 {B} --no-newton "
             "(C2, the FMA formulation of the same element kernel: the DMMA-vs-FMA comparison)")
full_capture(f"prof_{tag}_spmv.ncu-rep", f"{tag}_ncu_full_spmv.csv",
             f"ncu --set full --clock-control none --import-source on -k regex:'spmv_node_dot' -s 40 -c 1 {B} (C2, inside the PCG)")
full_capture(f"prof_{tag}_eas.ncu-rep", f"{tag}_ncu_full_eas.csv",
             "C4_N=48 ncu --set full --clock-control none --import-source on -k regex:'elem_eas' -s 2 -c 1 python tools/config_times.py C4 "
             "(Hex8 + E21 NeoHooke nu=0.499, 48^3)")
full_capture(f"prof_{tag}_q2.ncu-rep", f"{tag}_ncu_full_q2.csv",
             "C3_N=24 ncu --set full --clock-control none --import-source on -k regex:'elem_q2' -s 2 -c 1 python tools/config_times.py C3 "
             "(Hex27 SVK, 24^3)")
full_capture(f"prof_{tag}_easdg.ncu-rep", f"{tag}_ncu_full_easdg.csv",
             "C4_N=48 ncu --set full --clock-control none -k regex:'elem_easdg' -s 2 -c 1 python tools/config_times.py C4dg "
             "(Hex8 + H9 EAS::DisplacementGradient NeoHooke nu=0.499, 48^3)")
if len(traffic) > 1:
    json.dump(traffic, open(os.path.join(pr, f"{tag}_traffic.json"), "w"), indent=1)
