"""Per-kernel share of the summed device time from an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file ...).
usage: launch_shares.py <launches.csv>"""
import csv, sys, collections, re
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]; ki, gi, vi, ui = h.index("Kernel Name"), h.index("Grid Size"), h.index("Metric Value"), h.index("Metric Unit")
acc = collections.OrderedDict(); n = collections.Counter()
for r in rows[1:]:
    if "meao::" not in r[ki]: continue
    name = re.sub(r"void meao::<unnamed>::|\(.*", "", r[ki]) + " " + r[gi]
    t = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3 if float(r[vi].replace(",", "")) > 1000 else 1.0)
    acc[name] = acc.get(name, 0.0) + t; n[name] += 1
tot = sum(acc.values())
print(f"{'kernel grid':58s} launches  mean us   share")
for k, v in acc.items(): print(f"{k:58s} {n[k]:6d}  {v / n[k]:8.2f}  {100 * v / tot:5.1f} %")
print("frames", max(n.values()), "sum per frame us", round(tot / max(n.values()), 1))
