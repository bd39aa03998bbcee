"""Dump the SASS of the hot loop (instructions executed >= frac * max) with per-instruction samples."""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(raw.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None: cur["hdr"] = row
    elif cur is not None and row: cur["rows"].append(row)
b = [b for b in blocks if pat in b["name"]][0]
h = b["hdr"]; si, ii, src = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
mx = max(int(r[ii]) for r in b["rows"])
tot = sum(int(r[ii]) for r in b["rows"])
ops = {}
n = 0
for idx, r in enumerate(b["rows"]):
    e = int(r[ii])
    if e >= frac * mx:
        n += 1
        op = r[src].strip().split()[0]
        if op.startswith("@"): op = r[src].strip().split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
        if len(sys.argv) > 4: print(f"{idx:5d} {e:9d} {int(r[si]):6d} {r[src].strip()}")
print(b["name"], "loop instrs", n, "total exec", tot, "max exec", mx)
print(sorted(ops.items(), key=lambda kv: -kv[1]))
