#!/bin/bash
# usage: scripts/tune.sh "<defs1>" "<defs2>" ...   (run on the GPU box; rebuilds libmeao.so per variant and runs bench)
for defs in "$@"; do
  MEAO_NVCC_DEFS="$defs" python miniengineao_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
  for S in ${STREAMS:-3}; do
    python bench.py --steps 300 --warmup 20 --no-cpu --streams $S > /tmp/b.json 2>/tmp/b.err || { tail -3 /tmp/b.err; continue; }
    python - "$defs" $S <<'PY'
import json,sys
d=json.load(open('/tmp/b.json'))
k=d['kernels']
print(f"[{sys.argv[1]}] streams={sys.argv[2]} value={d['value']} ms={d['ms_per_step']} prep={k['prepare_depth']['ms']} ren1={k['render_ao L1']['ms']} ups21={k['blur_upsample L2->L1']['ms']} ups10={k['blur_upsample L1->L0']['ms']}")
PY
  done
done
