#!/bin/bash
# cuobjdump -sass mnemonic counts per kernel of the shipped libmeao.so -> profiles/r2_sass_summary.txt (no GPU needed)
OUT=${1:-profiles/r2_sass_summary.txt}
cuobjdump -sass miniengineao_b200/libmeao.so > /tmp/meao_all.sass
python - "$OUT" <<'PY'
import re, sys, subprocess, collections
out = sys.argv[1]
cur = None; per = collections.OrderedDict()
for ln in open("/tmp/meao_all.sass"):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"meao::\(anonymous namespace\)::|void |\(.*", "", cur); per[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if cur and m: per[cur][m.group(1)] += 1; per[cur]["_total"] += 1
cols = ["_total", "UTMALDG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "MUFU", "VIMNMX3", "F2I", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "LDL", "STL", "ACQBULK", "MEMBAR", "ATOMG", "RED"]
with open(out, "w") as f:
    f.write("cuobjdump -sass of miniengineao_b200/libmeao.so (sm_100a), static instruction counts per kernel\n")
    f.write(f"{'kernel':62s}" + "".join(f"{c.replace('_total','total'):>8s}" for c in cols) + "\n")
    for k, c in per.items():
        f.write(f"{k[:62]:62s}" + "".join(f"{c.get(x, 0):8d}" for x in cols) + "\n")
    tot = collections.Counter()
    for c in per.values(): tot.update(c)
    f.write(f"{'ALL KERNELS':62s}" + "".join(f"{tot.get(x, 0):8d}" for x in cols) + "\n")
    f.write("\nUTMALDG = TMA (cp.async.bulk.tensor), SYNCS = mbarrier, FFMA2/FADD2/FMUL2 = packed f32x2, MUFU = rcp.approx, VIMNMX3 = 3-input integer max\n"
            "(the grouped range tests), SHFL = warp shuffle (none: neighbouring blur taps live in one thread's registers), LDL/STL = local memory\n"
            "(only in the out-of-line IEEE slow paths).  ld.acquire.sys / st.release.sys of band_exchange_kernel show as LDG.E.STRONG.SYS / STG.E.STRONG.SYS.\n")
print(open(out).read())
PY
