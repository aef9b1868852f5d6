#!/bin/bash
# device time of every frame kernel under ncu with caches left as the pipeline leaves them (--cache-control none) vs flushed (default)
mkdir -p gpurun_out
for cc in none all; do
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --cache-control $cc -s 18 -c 18 --csv --log-file gpurun_out/r2j_ncu_cache_${cc}.csv python scripts/profile_frames.py 3840 2160 4 > /dev/null 2>&1
done
python - <<'PY'
import csv, re, collections
for cc in ("none", "all"):
    rows = [r for r in csv.reader(l for l in open(f"gpurun_out/r2j_ncu_cache_{cc}.csv") if l.startswith('"'))]
    h = rows[0]; ki, gi, mi, vi = h.index("Kernel Name"), h.index("Grid Size"), h.index("Metric Name"), h.index("Metric Value")
    acc = collections.OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum": continue
        name = re.sub(r"void meao::<unnamed>::|\(.*", "", r[ki]) + " " + r[gi]
        acc.setdefault(name, []).append(float(r[vi].replace(",", "")) / 1e3)
    print(f"--cache-control {cc}: " + "; ".join(f"{k}: {sum(v)/len(v):.2f} us" for k, v in acc.items()), "| sum", round(sum(sum(v)/len(v) for v in acc.values()), 1))
PY
