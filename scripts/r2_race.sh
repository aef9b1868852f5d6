#!/bin/bash
# racecheck of one frame with interior (TMA) tiles, full report
cat > /tmp/race.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from miniengineao_b200 import AmbientOcclusion, Camera, synth
from oracle.oracle import Oracle
W, H = int(sys.argv[1]), int(sys.argv[2])
depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=3))
ao = AmbientOcclusion(Camera(W, H), device=0, use_graph=False); ao.intensity = 1.1
got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
print("frame matches oracle:", np.array_equal(got, Oracle(W, H, intensity=1.1, threads=8).run(depth)))
PY
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/race.py 330 170 > gpurun_out/r2i_racecheck_330.txt 2>&1; tail -25 gpurun_out/r2i_racecheck_330.txt
MEAO_UPS_PERSIST_MIN_WAVES=0.01 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/race.py 700 420 > gpurun_out/r2i_racecheck_700_persist.txt 2>&1; tail -25 gpurun_out/r2i_racecheck_700_persist.txt
