#!/bin/bash
# A/B of build switches on one GPU.  usage: scripts/r2_ab.sh <tag> "<defs 1>" "<defs 2>" ...   ("" = default build as shipped)
TAG=$1; shift
mkdir -p gpurun_out
T=gpurun_out/${TAG}
QUICK='full_pipe_bit_exact_random or parameter_sweep or sky_pixels or linear_depth_ingest or native_depth_formats or row_bands_equal_whole_frame or proven_range or golden or variants_full_pipe'
for defs in "$@"; do
  tag=$(echo "x$defs" | tr -c 'A-Za-z0-9=' '_')
  MEAO_NVCC_DEFS="$defs" python miniengineao_b200/build.py --force > /dev/null 2>&1 || { echo "[$defs] build failed"; continue; }
  timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "$QUICK" > ${T}_quick_${tag}.log 2>&1
  echo "[$defs] quick parity: $(tail -1 ${T}_quick_${tag}.log)"
  for rep in 1 2; do
    timeout 300 python bench.py --steps 500 --warmup 20 --quick --no-cpu > ${T}_bench_${tag}_${rep}.json 2> ${T}_bench_${tag}_${rep}.err
  done
done
python miniengineao_b200/build.py --force > /dev/null 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    k = d.get("kernels") or {}
    print(f.split("_bench_")[1], "value", d["value"], "us/frame", round(d["ms_per_step"] * 1e3, 2), "serial", round(d["serial_frames"]["ms_per_frame"] * 1e3, 2),
          {n.replace("blur_upsample ", "ups").replace("render_ao ", "ren").replace("prepare_depth", "prep"): round(v["ms"] * 1e3, 1) for n, v in k.items()})
PY
