"""Per CUDA source line executed warp-instructions from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu_line_hist.py <csv> <function-substring> <file-substring> [top]"""
import csv, sys, collections
path, fsub, filesub = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
cur_file = cur_fn = None; hdr = None
acc = collections.Counter(); src = {}
n_fn = 0
for row in csv.reader(open(path)):
    if not row: continue
    if row[0] == "File Path": cur_file = row[1]; hdr = None; continue
    if row[0] == "Function Name":
        if cur_fn != row[1]: n_fn += 1
        cur_fn = row[1]; hdr = None; continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None or fsub not in (cur_fn or "") or filesub not in (cur_file or ""): continue
    if row[0] != "":     # a CUDA source line aggregate row
        try: acc[int(row[0])] += int(row[hdr.index("Instructions Executed")])
        except ValueError: pass
        src[int(row[0])] = row[1]
tot = sum(acc.values())
print("total", tot)
for ln, n in sorted(acc.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{ln:5d} {n:10d} {100*n/tot:5.1f}%  {src.get(ln,'')[:110]}")
