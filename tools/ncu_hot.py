"""Hot SASS instructions of one kernel from an ncu report (source page): python tools/ncu_hot.py rep kernel_regex [n]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(raw.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
for b in [b for b in blocks if pat in b['name']][:1]:
    h = b["hdr"]
    si, ii, src = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    ti = h.index("Thread Instructions Executed")
    stalls = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[si]) for r in b["rows"])
    tin = sum(int(r[ii]) for r in b["rows"])
    print(b["name"], "samples", tot, "warp-instr", tin, "avg threads", sum(int(r[ti]) for r in b["rows"]) / max(tin, 1))
    agg = {}
    for r in b["rows"]:
        for i in stalls:
            agg[h[i]] = agg.get(h[i], 0) + int(r[i])
    print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for idx, r in sorted(enumerate(b["rows"]), key=lambda kv: -int(kv[1][si]))[:n]:
        top = sorted(((int(r[i]), h[i]) for i in stalls), reverse=True)[:2]
        print(f"{idx:5d} {int(r[si]):7d} {100 * int(r[si]) / tot:5.1f}%  exec={int(r[ii]):9d} thr={int(r[ti]) / max(int(r[ii]), 1):5.1f} {r[src].strip()[:70]:70s} {top}")
