"""Runs a few 4K frames (stream launches, no graph) -- the target command for ncu captures.
usage: profile_frames.py [W H [frames]]   env: MEAO_HQ_MASK (0..15), MEAO_EXH (0/1), MEAO_DEBUG_VIEW (buffer id, 0 = none)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from miniengineao_b200 import AmbientOcclusion, Camera, synth

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ao = AmbientOcclusion(Camera(W, H), device=0, use_graph=False)
ao.intensity = 1.1
ao.highQualityMask = int(os.environ.get("MEAO_HQ_MASK", "0"))
ao.sampleExhaustively = os.environ.get("MEAO_EXH", "0") == "1"
d = torch.from_numpy(synth.lin01_to_raw(synth.corridor(W, H))).cuda()
o = torch.empty((H, W), dtype=torch.uint8, device="cuda")
for _ in range(n):
    ao.render(d, o)
view = int(os.environ.get("MEAO_DEBUG_VIEW", "0"))
if view:
    v = ao.debug_view(view)
torch.cuda.synchronize()
print("frames", n, "launches", ao.launch_count, "checksum", int(o.sum().item()))
