"""Turns an .ncu-rep (--set full) into the small CSV/JSON summaries committed under profiles/.
usage: summarize_ncu.py <report.ncu-rep> <out_prefix>"""
import csv, json, subprocess, sys, io
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']
idx = [hdr.index(w) for w in want if w in hdr]
with open(out + "_summary.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in rows[2:]:
        w.writerow([r[i] for i in idx])
def num(r, name):
    v = float(r[hdr.index(name)].replace(",", "")); u = units[hdr.index(name)]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
traffic = []
for r in rows[2:]:
    traffic.append({"kernel": r[hdr.index('Kernel Name')], "grid": r[hdr.index('launch__grid_size')],
                    "dram_bytes": num(r, 'dram__bytes_read.sum') + num(r, 'dram__bytes_write.sum'),
                    "time_us": float(r[hdr.index('gpu__time_duration.sum')]),
                    "warp_inst": float(r[hdr.index('smsp__inst_executed.sum')]),
                    "issue_active_pct": float(r[hdr.index('smsp__issue_active.avg.pct_of_peak_sustained_active')])})
json.dump(traffic, open(out + "_traffic.json", "w"), indent=1)
print(open(out + "_summary.csv").read()[:300])
