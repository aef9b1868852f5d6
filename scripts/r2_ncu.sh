#!/bin/bash
# ncu session (one GPU): full-set capture of one serial 4K frame (9 kernels, second frame), raw + SASS source pages exported as CSV,
# and the launch list of a short bench run.  usage: scripts/r2_ncu.sh <tag>
TAG=${1:-r2c}
mkdir -p gpurun_out
T=gpurun_out/${TAG}
ncu --set full --clock-control none --import-source on -s 9 -c 9 -f -o ${T}_full python scripts/profile_frames.py 3840 2160 2 > ${T}_ncu.log 2>&1
tail -3 ${T}_ncu.log
ncu -i ${T}_full.ncu-rep --page raw --csv > ${T}_raw.csv 2>/dev/null
ncu -i ${T}_full.ncu-rep --page source --csv --print-source sass > ${T}_source_sass.csv 2>/dev/null
gzip -f ${T}_source_sass.csv
python scripts/summarize_ncu.py ${T}_full.ncu-rep ${T} > /dev/null 2>&1
ls -la ${T}_full.ncu-rep
sz=$(stat -c %s ${T}_full.ncu-rep); if [ $sz -gt 40000000 ]; then rm ${T}_full.ncu-rep; fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file ${T}_launches.csv python bench.py --steps 2 --warmup 3 --quick --no-cpu > ${T}_launch_bench.log 2>&1
python - <<PY
import csv, io
rows = list(csv.reader(open("${T}_summary.csv")))
h = rows[0]
for r in rows[2:]:
    d = dict(zip(h, r))
    print(d['Kernel Name'][:40], d['launch__grid_size'], 'us', d['gpu__time_duration.sum'], 'regs', d['launch__registers_per_thread'], 'inst', d['smsp__inst_executed.sum'],
          'issue%', d['smsp__issue_active.avg.pct_of_peak_sustained_active'], 'fma%', d.get('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
          'xu%', d.get('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'), 'conflicts', d.get('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'))
PY
