#!/bin/bash
# Final round-2 session (one GPU): full GPU suite, smoke, ncu captures of the shipped build, full bench, reference arm, sanitizer.
mkdir -p gpurun_out
T=gpurun_out/r2f
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --tb=short > ${T}_gpu_tests.log 2>&1
echo "FULL TESTS: $(tail -1 ${T}_gpu_tests.log)"; grep -E "^FAILED|^ERROR" ${T}_gpu_tests.log | head
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > ${T}_smoke.log 2>&1; tail -1 ${T}_smoke.log
bash scripts/r2_ncu.sh r2f
python scripts/launch_shares.py ${T}_launches.csv > ${T}_launch_shares.txt; cat ${T}_launch_shares.txt
gunzip -k -f ${T}_source_sass.csv.gz
for k in "blur_upsample_kernel<(bool)0, (bool)1>" "render_ao_kernel<(int)0, (bool)0, (int)32>" "prepare_depth_kernel"; do python scripts/ncu_phase_stalls.py ${T}_source_sass.csv "$k" 10; done > ${T}_phase_stalls.txt 2>&1
rm -f ${T}_source_sass.csv
timeout 600 python bench.py --steps 200 --warmup 20 > ${T}_bench_4k.json 2> ${T}_bench_4k.err; echo "BENCH rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > ${T}_bench_4k_driverflags.json 2> ${T}_bench_4k_driverflags.err; echo "BENCH(driver flags) rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > ${T}_bench_reference.json 2> ${T}_bench_reference.err
python - <<'PY'
import json
for f in ("gpurun_out/r2f_bench_4k.json", "gpurun_out/r2f_bench_4k_driverflags.json", "gpurun_out/r2f_bench_reference.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("r2f_")[1], "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline"))
    if d.get("roofline"): print("  roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "pipe", d["roofline"]["pipe_frac"], "serial", d["serial_frames"])
    if d.get("configs"): print("  configs", json.dumps(d["configs"])[:1500])
PY
# compute-sanitizer: memcheck + racecheck on a small frame (every kernel, TMA and border paths) and memcheck on two connected bands
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from miniengineao_b200 import AmbientOcclusion, Camera, synth, rowtile
from oracle.oracle import Oracle
W, H = 330, 170
depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=3)); depth[40:60, 100:180] = 0.0
ao = AmbientOcclusion(Camera(W, H), device=0); ao.intensity = 1.1
got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
print("frame matches oracle:", np.array_equal(got, Oracle(W, H, intensity=1.1).run(depth)))
ss = AmbientOcclusion(Camera(W, H), device=0); ss.intensity = 1.1; ss.singleScale = True
print("single-scale matches:", np.array_equal(ss.render(torch.from_numpy(depth).cuda()).cpu().numpy(), Oracle(W, H, intensity=1.1, single_scale=True).run(depth)))
if len(sys.argv) > 1:
    W, H = 640, 736
    depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=4)); ref = Oracle(W, H, intensity=1.1, threads=8).run(depth)
    cuts = rowtile.partition(H, 2); ctx = []
    for i in range(2):
        a = AmbientOcclusion(Camera(W, H), device=0); a.intensity = 1.1
        a.set_row_band(cuts[i], cuts[i + 1], *rowtile.neighbours(cuts, i)); ctx.append(a)
    hs = [a.band_export() for a in ctx]; ctx[0].band_connect(1, hs[1]); ctx[1].band_connect(0, hs[0])
    st = [torch.cuda.Stream() for _ in range(2)]
    outs = [torch.zeros((cuts[i + 1] - cuts[i], W), dtype=torch.uint8, device="cuda") for i in range(2)]
    ds = [torch.from_numpy(depth[cuts[i]:cuts[i + 1]]).cuda() for i in range(2)]
    for rep in range(2):
        for i in range(2): ctx[i].band_step(ds[i], outs[i], stream=st[i])
        torch.cuda.synchronize()
    print("bands match oracle:", np.array_equal(np.concatenate([o.cpu().numpy() for o in outs]), ref), [a.band_status() for a in ctx])
PY
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool (330x170 frame + single-scale)" >> ${T}_compute_sanitizer.txt
  timeout 600 compute-sanitizer --tool $tool python /tmp/san.py 2>&1 | grep -E "matches|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -12 >> ${T}_compute_sanitizer.txt
done
echo "(the native exchange is not run under the sanitizer: it serialises kernels, and two neighbouring bands on ONE GPU need their exchange kernels resident together -- the run ends in the kernel's own time-out, error 1, as designed)" >> ${T}_compute_sanitizer.txt
cat ${T}_compute_sanitizer.txt
