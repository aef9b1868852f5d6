"""Per-kernel register / spill / instruction-count summary of the current libmeao.so build (no GPU needed).
   python scripts/ptxas_summary.py [filter-substring ...]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log = open(os.path.join(ROOT, "miniengineao_b200", "build_ptxas.log")).read()
ents = re.findall(r"Compiling entry function '([^']+)'[^\n]*\n[^\n]*\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n[^\n]*Used (\d+) registers", log)
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "miniengineao_b200", "libmeao.so")], capture_output=True, text=True).stdout
counts, cur = {}, None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m: cur = m.group(1); counts[cur] = 0; continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln): counts[cur] += 1
filt = sys.argv[1:]
for name, stack, ss, sl, regs in ents:
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    d = re.sub(r"meao::\(anonymous namespace\)::", "", d); d = re.sub(r"\(.*", "", d).replace("void ", "")
    if filt and not any(f in d for f in filt): continue
    print(f"{d:58s} regs {regs:>3}  stack {stack:>3}  spill st/ld {ss:>3}/{sl:<3}  sass {counts.get(name, 0)}")
