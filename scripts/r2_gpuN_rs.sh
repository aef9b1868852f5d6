#!/bin/bash
# usage: scripts/r2_gpuN_rs.sh N "rs list" [full-rs]  -- band-stream sweep of the native row-banded 8K, optionally the full bench with --band-streams full-rs
N=$1; mkdir -p gpurun_out; T=gpurun_out/r2n${N}
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N "$@"; }
for rs in $2; do
  run --steps 200 --warmup 20 --only-8k --band-streams $rs > ${T}_8k_native_rs${rs}.json 2> ${T}_8k_native_rs${rs}.err
  echo "native rs=$rs: $(grep "{" ${T}_8k_native_rs${rs}.json | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('us/step', round(d['ms_per_step']*1e3,1), 'speedup', d['speedup_vs_1gpu'], '1gpu us', round(d['one_gpu_ms_per_frame']*1e3,1), 'ok', d.get('bands_match_oracle'), 'err', d.get('exchange_error'))" 2>&1 | tail -1)"
done
if [ -n "$3" ]; then
  run --steps 200 --warmup 20 --band-streams $3 > ${T}_bench.json 2> ${T}_bench.err
  grep "{" ${T}_bench.json | tail -1 > ${T}_bench_line.json
  python -c "
import json; d=json.load(open('${T}_bench_line.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['d16_ingest']['value'])
print(json.dumps(d['configs'])[:1800])"
fi
