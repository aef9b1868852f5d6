#!/bin/bash
# Run on the GPU box.  usage: scripts/pick_best.sh "<defs A>" "<defs B>" ...
# The FIRST configuration is assumed to be the libmeao.so that travelled with the snapshot (no rebuild).  Every other
# configuration is rebuilt on the box.  Each one: quick parity subset (skipped for an empty defs string = the validated
# default), then bench (300 steps, 5 streams).  Finally the fastest parity-clean configuration is rebuilt (if it is not the
# current build) and the FULL gpu test-suite runs on it.  Results: gpurun_out/pick_best.txt
mkdir -p gpurun_out
OUT=gpurun_out/pick_best.txt
: > $OUT
QUICK='full_pipe_bit_exact_random or parameter_sweep or sky_pixels or linear_depth_ingest or native_depth_formats or row_bands_equal_whole_frame'
best=""; bestv=0; cur=""; first=1
for defs in "$@"; do
  if [ $first -eq 0 ]; then
    MEAO_NVCC_DEFS="$defs" python miniengineao_b200/build.py --force > /dev/null 2>&1 || { echo "[$defs] build failed" >> $OUT; continue; }
  fi
  first=0; cur="$defs"
  if [ -n "$defs" ]; then
    if ! python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "$QUICK" > /tmp/quick.log 2>&1; then
      echo "[$defs] QUICK PARITY FAILED: $(tail -3 /tmp/quick.log | tr '\n' ' ')" >> $OUT; continue
    fi
  fi
  python bench.py --steps 1000 --warmup 20 --no-cpu --no-rowtile > /tmp/b.json 2>/tmp/b.err || { echo "[$defs] bench failed: $(tail -2 /tmp/b.err | tr '\n' ' ')" >> $OUT; continue; }
  v=$(python -c "import json; d=json.load(open('/tmp/b.json')); k=d['kernels']; print(d['value']); print('prep', k['prepare_depth']['ms'], 'ren1', k['render_ao L1']['ms'], 'ups21', k['blur_upsample L2->L1']['ms'], 'ups10', k['blur_upsample L1->L0']['ms'], 'serial', d['serial_frames']['ms_per_frame'])")
  val=$(echo "$v" | head -1)
  echo "[$defs] value=$val $(echo "$v" | tail -1)" >> $OUT
  cp /tmp/b.json "gpurun_out/pick_best_bench_$(echo "$defs" | tr -c 'A-Za-z0-9=' '_').json"
  if python -c "import sys; sys.exit(0 if float('$val') > float('$bestv') else 1)"; then best="$defs"; bestv=$val; fi
done
echo "BEST [$best] $bestv" >> $OUT
if [ "$best" != "$cur" ]; then
  MEAO_NVCC_DEFS="$best" python miniengineao_b200/build.py --force > /dev/null 2>&1 || echo "rebuild of best failed" >> $OUT
fi
if [ -z "$SKIP_FULL" ]; then
  python -m pytest tests -m gpu -q --maxfail=5 --tb=short > gpurun_out/pick_best_tests.log 2>&1
  echo "FULL TESTS on [$best]: $(tail -1 gpurun_out/pick_best_tests.log)" >> $OUT
fi
cat $OUT
