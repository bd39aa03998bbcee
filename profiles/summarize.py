"""Summarise gpurun_out ncu artefacts into small tracked text files under profiles/.
usage: python profiles/summarize.py <tag>   (reads gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep)"""
import csv
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1]
out = open(f"profiles/{tag}_summary.txt", "w")


def emit(*a):
    print(*a)
    print(*a, file=out)


try:
    rows = list(csv.reader(l for l in open(f"gpurun_out/launches_{tag}.csv") if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    d = defaultdict(list)
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        d[r[ki][:70]].append(v / (1e6 if r[ui] in ("ns", "nsecond") else 1e3 if r[ui] in ("us", "usecond") else 1))
    tot = sum(sum(v) for v in d.values())
    emit(f"# ncu launch list ({tag}): gpu__time_duration.sum per kernel, --clock-control none (cold-cache, serialised)")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        emit(f"{k:70s} n={len(v):4d} total_ms={sum(v):10.3f} mean_ms={sum(v) / len(v):9.4f} share={100 * sum(v) / tot:5.1f}%")
except FileNotFoundError:
    emit("no launch list")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
try:
    raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    emit(f"\n# ncu --set full ({tag})")
    for r in rows[2:]:
        emit("kernel:", r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                emit(f"  {w} = {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
except Exception as e:  # noqa
    emit("no full capture", e)
