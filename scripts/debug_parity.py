"""Per-buffer mismatch report of the CUDA path vs the oracle (debugging aid, run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from miniengineao_b200 import AmbientOcclusion, Camera, synth
from oracle.oracle import Oracle

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 256)
graph = os.environ.get("GRAPH", "0") == "1"
depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=1))
mask, exh = int(os.environ.get("MEAO_HQ_MASK", "0")), os.environ.get("MEAO_EXH", "0") == "1"
ao = AmbientOcclusion(Camera(W, H), device=0, use_graph=graph)
ao.highQualityMask, ao.sampleExhaustively = mask, exh
orc = Oracle(W, H, threads=8, high_quality_mask=mask, sample_exhaustively=exh)
ref = orc.run(depth)
try:
    got = ao.render(torch.from_numpy(depth).cuda())
    torch.cuda.synchronize()
    got = got.cpu().numpy()
except Exception as e:
    print("render failed:", e); sys.exit(1)
print("final mismatches", int((got != ref).sum()), "of", got.size)
for bid in list(range(1, 18)) + [17 + k for k in range(1, 5) if (mask >> (k - 1)) & 1]:
    g, r = ao.debug_buffer(bid), orc.buffer(bid)
    if g.dtype == np.uint8:
        r = orc.codes(bid); bad = g != r
    elif g.dtype == np.float16:
        bad = g.view(np.uint16) != r.astype(np.float16).view(np.uint16)
    else:
        bad = g.view(np.uint32) != r.view(np.uint32)
    n = int(bad.sum())
    msg = ""
    if n:
        idx = np.argwhere(bad)
        msg = f" first {idx[0].tolist()} got {g[tuple(idx[0])]} ref {r[tuple(idx[0])]}; rows {idx[:,-2].min()}..{idx[:,-2].max()} cols {idx[:,-1].min()}..{idx[:,-1].max()}"
    print(f"  {bid:2d} {ao.DEBUG_NAMES[bid]:18s} mismatches {n:8d} / {g.size}{msg}")
if os.environ.get("MEAO_VIEWS", "0") == "1":
    bad = 0
    for bid in range(1, 18):
        v = ao.debug_view(bid)
        ao.synchronize()
        bad += int((v.cpu().numpy() != orc.debug_view(bid)).sum())
    print("debug view mismatches", bad)
