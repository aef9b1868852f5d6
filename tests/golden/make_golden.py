"""Generates tests/golden/*.npz from the CPU oracle (oracle/meao_oracle.c).

The reference (HLSL compute + Unity C#) cannot be executed in this image and ships no golden vectors of
its own, so these fixtures are SELF-GENERATED: they pin the oracle against regressions
and let the GPU tests compare against committed bytes without running the oracle.  Re-run:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from miniengineao_b200 import synth  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

CASES = {
    # name: (W, H, lin01 generator, oracle kwargs)
    "flat_sphere_96": (96, 96, lambda: synth.flat_sphere(96, 96), dict()),
    "random_83x61_sponza_intensity": (83, 61, lambda: synth.random_depth(83, 61, seed=42), dict(intensity=1.1)),
    "random_130x70_params": (130, 70, lambda: synth.random_depth(130, 70, seed=7),
                             dict(intensity=1.2, thickness_modifier=2.0, blur_tolerance=-3.0, upsample_tolerance=-5.0, noise_filter_tolerance=-1.0)),
    "corridor_160x90": (160, 90, lambda: synth.corridor(160, 90), dict(intensity=1.1)),
    # the shader variants the reference ships but never dispatches: extra key "variants" =
    # [single_pass_stereo, sample_exhaustively, high_quality_mask], extra buffers 18..21 where selected
    "variant_exhaustive_101x67": (101, 67, lambda: synth.random_depth(101, 67, seed=5), dict(intensity=1.1, sample_exhaustively=True)),
    "variant_hq15_exhaustive_118x90": (118, 90, lambda: synth.random_depth(118, 90, seed=6),
                                       dict(intensity=1.2, thickness_modifier=2.0, high_quality_mask=15, sample_exhaustively=True)),
    "variant_hq10_corridor_144x80": (144, 80, lambda: synth.corridor(144, 80), dict(high_quality_mask=0b1010)),
    # BASELINE.json configs[0]: 256 x 256 flat + sphere, SINGLE-SCALE plan (Downsample -> Render level 1 -> final-style Upsample);
    # "variants" then carries a 4th entry = single_scale, and only the buffers that plan writes are stored
    "single_scale_flat_sphere_256": (256, 256, lambda: synth.flat_sphere(256, 256), dict(intensity=1.1, single_scale=True)),
}


def main():
    only_new = "--only-new" in sys.argv       # keep the committed bytes of existing fixtures untouched
    for name, (W, H, gen, kw) in CASES.items():
        if only_new and os.path.exists(os.path.join(HERE, name + ".npz")):
            continue
        lin = gen()
        depth = synth.lin01_to_raw(lin)
        o = Oracle(W, H, **kw)
        ao = o.run(depth)
        data = {"depth": depth, "ao": ao, "params": np.array([kw.get("noise_filter_tolerance", 0.0), kw.get("blur_tolerance", -4.6),
                                                              kw.get("upsample_tolerance", -12.0), kw.get("thickness_modifier", 1.0),
                                                              kw.get("intensity", 1.0)], np.float32)}
        mask = kw.get("high_quality_mask", 0)
        single = bool(kw.get("single_scale", False))
        if single:
            data["variants"] = np.array([0, int(kw.get("sample_exhaustively", False)), mask, 1], np.int32)
        elif mask or kw.get("sample_exhaustively"):
            data["variants"] = np.array([0, int(kw.get("sample_exhaustively", False)), mask], np.int32)
        bids = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 17] if single else list(range(1, 18)) + [17 + k for k in range(1, 5) if (mask >> (k - 1)) & 1]
        for bid in bids:
            b = o.buffer(bid)
            if bid >= 10:
                data[f"buf{bid}"] = o.codes(bid)
            elif bid == 1 or 6 <= bid <= 9:
                data[f"buf{bid}"] = b.astype(np.float16)
            else:
                data[f"buf{bid}"] = b.copy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, W, H, "ao mean", ao.mean())


if __name__ == "__main__":
    main()
