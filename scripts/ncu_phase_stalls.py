"""Per-phase (between BAR.SYNCs) stall-reason breakdown of one kernel from `ncu --page source --csv --print-source sass`.
usage: ncu_phase_stalls.py <csv> <kernel-name-substring> [top]"""
import csv, collections, sys
f, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 16
blocks = []; cur = None
for row in csv.reader(open(f)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; blocks.append(cur); continue
    if cur is None: continue
    if cur["hdr"] is None: cur["hdr"] = row; continue
    cur["rows"].append(row)
if sub == "list":
    for i, b in enumerate(blocks): print(i, b["name"][:110], len(b["rows"]))
    sys.exit(0)
cands = [b for b in blocks if sub in b["name"]]
b = cands[int(sys.argv[4])] if len(sys.argv) > 4 else cands[0]
print("=======", b["name"][:100])
h = b["hdr"]; si = h.index("Source"); sa = h.index("Warp Stall Sampling (All Samples)"); ie = h.index("Instructions Executed")
reasons = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
ridx = {c: h.index(c) for c in reasons}
phase = 0; acc = collections.defaultdict(collections.Counter)
for r in b["rows"]:
    try: s = int(r[sa]); e = int(r[ie])
    except ValueError: continue
    acc[phase]["samples"] += s; acc[phase]["instr"] += e; acc[phase]["static"] += 1
    for c in reasons:
        try: acc[phase][c] += int(r[ridx[c]])
        except ValueError: pass
    if "BAR.SYNC" in r[si]: phase += 1
tot = sum(v["samples"] for v in acc.values()); toti = sum(v["instr"] for v in acc.values())
print("total samples", tot, "warp-instr", toti)
for p, v in acc.items():
    rs = sorted(((v[c], c.replace("stall_", "")) for c in reasons), reverse=True)[:6]
    print(f"phase {p}: samples {100*v['samples']/tot:5.1f}%  instr {100*v['instr']/toti:5.1f}% ({v['instr']})  static {v['static']}  | " +
          ", ".join(f"{n}:{100*x/max(v['samples'],1):.0f}%" for x, n in rs))
rows = []
for r in b["rows"]:
    try: rows.append((int(r[sa]), int(r[ie]), r[si].strip()[:70]))
    except ValueError: pass
for s, e, src in sorted(rows, reverse=True)[:top]: print(f"{s:6d} {100*s/tot:5.1f}% exec {e:8d}  {src}")
