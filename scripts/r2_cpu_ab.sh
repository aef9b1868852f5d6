#!/bin/bash
# CPU-side A/B on the GPU box's host: the oracle's three threading modes at 4K, all host threads; plus the C bands client test.
mkdir -p gpurun_out
for m in 0 1 2; do gcc -O2 -fno-tree-vectorize -ffp-contract=off -fPIC -std=c11 -DMEAO_ORACLE_POOL=$m -shared -o /tmp/orc_pool$m.so oracle/meao_oracle.c -lm -lpthread; done
python - <<'PY' | tee gpurun_out/r2g_oracle_threading.txt
import ctypes as C, os, sys, time, statistics
sys.path.insert(0, os.getcwd())
import numpy as np
from miniengineao_b200 import synth
from oracle import oracle as O
W, H = 3840, 2160
depth = synth.lin01_to_raw(synth.corridor(W, H))
cores = os.cpu_count()
print("host threads", cores)
for mode, name in ((0, "round 1: create + join per stage, row split"), (1, "pool, fixed (row x column) unit ranges"), (2, "pool, atomic cursor")):
    O._libs.pop("fma", None)
    real = O.os.path.join
    O.os.path.join = lambda *a, _m=mode: (f"/tmp/orc_pool{_m}.so" if a[-1] == "libmeao_oracle.so" else real(*a))
    try:
        for threads in (cores, cores // 2):
            o = O.Oracle(W, H, threads=threads, intensity=1.1)
            o.run(depth); o.run(depth)
            ts = []
            for _ in range(8):
                t = time.perf_counter(); o.run(depth); ts.append(time.perf_counter() - t)
            print(f"mode {mode} ({name}), {threads} threads: median {W * H / statistics.median(ts) / 1e6:.1f} Mpx/s, best {W * H / min(ts) / 1e6:.1f}")
    finally:
        O.os.path.join = real
PY
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "pure_c_client" 2>&1 | tail -3
