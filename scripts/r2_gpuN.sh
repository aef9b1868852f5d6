#!/bin/bash
# Round-2 multi-GPU session: usage scripts/r2_gpuN.sh N [full].  Row-banded 8K with the native exchange (and the NCCL p2p form
# for comparison), band-stream sweep; with "full" also the complete bench line at N GPUs.
N=${1:-2}
mkdir -p gpurun_out
T=gpurun_out/r2n${N}
nvidia-smi topo -m > ${T}_topo.txt 2>&1
run() { timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N "$@"; }
for rs in ${RS_LIST:-3 4 2}; do
  run --steps 200 --warmup 20 --only-8k --band-streams $rs > ${T}_8k_native_rs${rs}.json 2> ${T}_8k_native_rs${rs}.err
  echo "native rs=$rs: $(grep "{" ${T}_8k_native_rs${rs}.json | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d.get('exchange'), 'us/step', round(d['ms_per_step']*1e3,1), 'speedup', d['speedup_vs_1gpu'], '1gpu us', round(d['one_gpu_ms_per_frame']*1e3,1), 'ok', d.get('bands_match_oracle'), 'err', d.get('exchange_error'))" 2>&1 | tail -1)"
  tail -2 ${T}_8k_native_rs${rs}.err
done
run --steps 200 --warmup 20 --only-8k --band-streams 3 --band-mode p2p > ${T}_8k_p2p_rs3.json 2> ${T}_8k_p2p_rs3.err
echo "p2p rs=3: $(grep "{" ${T}_8k_p2p_rs3.json | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d.get('exchange'), 'us/step', round(d['ms_per_step']*1e3,1), 'speedup', d['speedup_vs_1gpu'], 'ok', d.get('bands_match_oracle'))" 2>&1 | tail -1)"
run --steps 200 --warmup 20 --only-8k --band-streams 2 --band-mode p2p > ${T}_8k_p2p_rs2.json 2> ${T}_8k_p2p_rs2.err
echo "p2p rs=2: $(grep "{" ${T}_8k_p2p_rs2.json | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d.get('exchange'), 'us/step', round(d['ms_per_step']*1e3,1), 'speedup', d['speedup_vs_1gpu'], 'ok', d.get('bands_match_oracle'))" 2>&1 | tail -1)"
if [ "$2" = "full" ]; then
  run --steps 200 --warmup 20 > ${T}_bench.json 2> ${T}_bench.err
  echo "FULL rc=$? $(tail -2 ${T}_bench.err)"
  grep "{" ${T}_bench.json | tail -1 > ${T}_bench_line.json
  python -c "
import json; d=json.load(open('${T}_bench_line.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['d16_ingest']['value'], 'numa', d['config']['numa'])
print(json.dumps(d['configs'])[:1800])"
fi
