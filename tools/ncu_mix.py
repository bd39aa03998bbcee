"""Dynamic instruction mix of one kernel from an ncu report: python tools/ncu_mix.py rep kernel_substr [--dump]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(raw.splitlines()):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None: cur["hdr"] = row
    elif cur is not None and row: cur["rows"].append(row)
b = [b for b in blocks if pat in b["name"]][0]
h = b["hdr"]; ii, src = h.index("Instructions Executed"), h.index("Source")
tot = sum(int(r[ii]) for r in b["rows"])
ops = {}
for idx, r in enumerate(b["rows"]):
    e = int(r[ii])
    t = r[src].strip().split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + e
    if len(sys.argv) > 3 and e > 0: print(f"{idx:5d} {e:9d} {r[src].strip()}")
print(b["name"], "total warp-instr", tot)
for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:40]:
    print(f"  {k:10s} {v:12d} {100*v/tot:5.1f}%")
