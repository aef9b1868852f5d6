"""Aggregate an `ncu --page source --csv` dump: executed warp-instructions per opcode and per source line.
usage: ncu_opcode_hist.py <csv> <kernel-index (0-based, in file order)> [view=sass]"""
import csv, sys, collections, re
path, kidx = sys.argv[1], int(sys.argv[2])
blocks, cur = [], None
for row in csv.reader(open(path)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; blocks.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = row; continue
    cur["rows"].append(row)
b = blocks[kidx]
h = b["hdr"]; si, ei = h.index("Source"), h.index("Instructions Executed")
tot = 0; ops = collections.Counter()
for r in b["rows"]:
    try: n = int(r[ei])
    except: continue
    src = r[si].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src[:12]
    ops[op.split(".")[0] if len(sys.argv) < 4 else op] += n; tot += n
print(b["name"][:90]); print("total warp-instr", tot)
for op, n in ops.most_common(40): print(f"  {op:14s} {n:10d} {100*n/tot:5.1f}%")
