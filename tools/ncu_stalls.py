"""Stall-reason shares (pc sampling), SM-active fraction and the SM -> L2 path of every kernel in an ncu report:
python tools/ncu_stalls.py gpurun_out/prof_<tag>.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def num(r, h):
    try:
        return float(r[col[h]].replace(",", ""))
    except (KeyError, ValueError):
        return float("nan")


seen = set()
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0]
    if name in seen:
        continue
    seen.add(name)
    st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), num(r, h)) for h in hdr
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    tot = sum(v for _, v in st if v == v) or 1.0
    top = ", ".join(f"{k} {100 * v / tot:.0f} %" for k, v in sorted(st, key=lambda kv: -kv[1])[:6])
    print(f"kernel: {name}")
    print(f"  stall samples: {top}")
    print(f"  sm__cycles_active.avg / sm__cycles_elapsed.avg = {num(r, 'sm__cycles_active.avg') / num(r, 'sm__cycles_elapsed.avg'):.2f}")
    for h in ("l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed",
              "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
              "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
              "l1tex__m_l1tex2xbar_write_sectors_mem_lg_op_st.sum", "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum"):
        print(f"  {h} = {num(r, h):.4g}")
