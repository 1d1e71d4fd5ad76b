"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (CSV):
   python tools/ncu_summary.py launches <launches.csv> <out.csv> "<command>"
   python tools/ncu_summary.py full <prof.ncu-rep> <out.csv> "<command>"
"""
import collections
import csv
import re
import subprocess
import sys


def launches(src, out, cmd):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            t = float(row.get("Metric Value", "0").replace(",", ""))
        except ValueError:
            continue
        unit = row.get("Metric Unit", "")
        t *= {"us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", row.get("Kernel Name", "")).replace("void ", "")
        tot[name][0] += 1
        tot[name][1] += t
    allt = sum(v[1] for v in tot.values())
    with open(out, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none); cold-cache, serialised: compare SHARES\n")
        f.write(f"# command: {cmd}\n")
        f.write("total_ms,share_pct,launches,mean_us,kernel\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:24]:
            f.write(f"{v[1] / 1e6:.3f},{100 * v[1] / allt:.2f},{v[0]},{v[1] / v[0] / 1e3:.1f},{k}\n")


def full(rep, out, cmd):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
            "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
            "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
            "smsp__average_warp_latency_issue_stalled_not_selected.ratio", "smsp__average_warp_latency_issue_stalled_membar.ratio",
            "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_sleeping.ratio",
            "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio", "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
            "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
            "sm__cycles_active.avg", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on (one B200)\n")
        f.write(f"# command: {cmd}\n")
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([rows[1][i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])
