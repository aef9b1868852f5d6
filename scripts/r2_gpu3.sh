#!/bin/bash
# session 3 (one GPU): ncu captures of the current build + render tile variants in serial / throughput mode
bash scripts/r2_ncu.sh r2c
for t in 1 2; do
  MEAO_REN_TILE=$t python bench.py --steps 300 --warmup 20 --quick --no-cpu > gpurun_out/r2c_bench_tile$t.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_tile$t.json')); k=d['kernels']
print('TILE=$t', d['value'], round(d['ms_per_step']*1e3,2), 'serial', round(d['serial_frames']['ms_per_frame']*1e3,2), {n:round(v['ms']*1e3,1) for n,v in k.items()})"
done
