#!/bin/bash
# compute-sanitizer over one small frame + one banded frame (run on the GPU box); summary lines only
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool"
  compute-sanitizer --tool $tool --print-limit 5 python scripts/debug_parity.py 330 170 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|final mismatches|Error|hazard" | head -8
done
# the undispatched shader variants (wide render, exhaustive sampling, premin upsample) + all 17 debug views
for tool in memcheck racecheck; do
  echo "== $tool (variants)"
  MEAO_HQ_MASK=15 MEAO_EXH=1 MEAO_VIEWS=1 compute-sanitizer --tool $tool --print-limit 5 python scripts/debug_parity.py 330 170 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|final mismatches|debug view mismatches|Error|hazard" | head -8
done
