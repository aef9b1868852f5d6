"""Times the undispatched shader variants at 4K on cuda:0: per-kernel event times of one serial frame
(meao_profile_frame) and the 3-stream graph-replay throughput, for the reference configuration and for
SAMPLE_EXHAUSTIVELY / high-quality masks.  Writes one JSON object to the path given as argv[1] (default stdout)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from miniengineao_b200 import AmbientOcclusion, Camera, synth  # noqa: E402

W, H = 3840, 2160
CONFIGS = {"reference": dict(), "exhaustive": dict(exh=True), "hq_level4": dict(mask=8), "hq_level3_4": dict(mask=12),
           "hq_all": dict(mask=15), "hq_all_exhaustive": dict(mask=15, exh=True)}
S, NBUF, K = 3, 8, 300


def main():
    base = torch.from_numpy(synth.lin01_to_raw(synth.corridor(W, H))).cuda()
    depths = [torch.roll(base, shifts=37 * i, dims=1).contiguous() for i in range(NBUF)]
    outs = [torch.empty((H, W), dtype=torch.uint8, device="cuda") for _ in range(NBUF)]
    streams = [torch.cuda.Stream() for _ in range(S)]
    res = {"workload": f"{W}x{H} corridor, intensity 1.1", "streams": S, "frames": K, "configs": {}}
    for name, cfg in CONFIGS.items():
        aos = []
        for _ in range(S):
            a = AmbientOcclusion(Camera(W, H), device=0)
            a.intensity = 1.1
            a.highQualityMask, a.sampleExhaustively = cfg.get("mask", 0), cfg.get("exh", False)
            aos.append(a)
        prof = None
        for _ in range(3):
            prof = aos[0].profile_frame(depths[0], outs[0])
        for i in range(NBUF * S):
            aos[i % S].render(depths[i % NBUF], outs[i % NBUF], stream=streams[i % S])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main_s = torch.cuda.current_stream()
        e0.record(main_s)
        for st in streams:
            st.wait_event(e0)
        for i in range(K):
            aos[i % S].render(depths[i % NBUF], outs[i % NBUF], stream=streams[i % S])
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            main_s.wait_event(ev)
        e1.record(main_s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        res["configs"][name] = {"us_per_frame": 1e3 * ms / K, "mpix_per_s": W * H * K / (ms * 1e-3) / 1e6,
                                "kernels_per_frame": aos[0].kernels_per_frame,
                                "serial_kernel_us": {n: round(1e3 * t, 2) for n, t in prof}}
        for a in aos:
            a.close()
    s = json.dumps(res, indent=1)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(s)
    print(s)


if __name__ == "__main__":
    main()
