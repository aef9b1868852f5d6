#!/bin/bash
# Round-2 GPU session 1 (one GPU): full GPU test-suite, full bench, and the env-switch A/B runs (PDL level, render tile rule).
mkdir -p gpurun_out
T=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > ${T}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --tb=short > ${T}_gpu_tests.log 2>&1
echo "TESTS: $(tail -1 ${T}_gpu_tests.log)"
grep -E "FAILED|Error|error" ${T}_gpu_tests.log | head -20
timeout 600 python bench.py --steps 200 --warmup 20 > ${T}_bench_4k.json 2> ${T}_bench_4k.err
echo "BENCH rc=$? $(tail -3 ${T}_bench_4k.err)"
for cfg in "MEAO_PDL=0" "MEAO_REN_TILE=0" "MEAO_PDL=0 MEAO_REN_TILE=0" "MEAO_PDL=1"; do
  tag=$(echo "$cfg" | tr -c 'A-Za-z0-9=' '_')
  env $cfg timeout 300 python bench.py --steps 200 --warmup 20 --quick --no-cpu > ${T}_bench_${tag}.json 2> ${T}_bench_${tag}.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2a_bench_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    k = d.get("kernels") or {}
    print(f.split("r2a_bench_")[1], "value", d["value"], "us/frame", round(d["ms_per_step"] * 1e3, 2), "serial", round(d["serial_frames"]["ms_per_frame"] * 1e3, 2),
          "pdl", d["config"].get("pdl_level"), "e2e", d["e2e"]["value"], d["e2e"]["d16_ingest"]["value"],
          {n.replace("blur_upsample ", "ups").replace("render_ao ", "ren").replace("prepare_depth", "prep"): round(v["ms"] * 1e3, 1) for n, v in k.items()})
    if "configs" in d and d["configs"]:
        print("   configs:", json.dumps(d["configs"])[:1500])
PY
