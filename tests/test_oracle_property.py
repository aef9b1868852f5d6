"""Property-based pins of the oracle (CPU): for random small frame sizes and parameters the literal thread-group
restatement and the independent global formulation must agree bit for bit on all 17 buffers (no-FMA builds)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from miniengineao_b200 import synth
from oracle import direct_formulation as DF
from oracle.oracle import Oracle


@settings(max_examples=25, deadline=None)
@given(W=st.integers(1, 70), H=st.integers(1, 70), seed=st.integers(0, 10_000),
       intensity=st.floats(0.0, 2.0), thickness=st.floats(1.0, 10.0), blur=st.floats(-8.0, -1.0),
       ups=st.floats(-12.0, -1.0), noise=st.floats(-8.0, 0.0), reversed_z=st.booleans())
def test_two_restatements_agree(W, H, seed, intensity, thickness, blur, ups, noise, reversed_z):
    kw = dict(intensity=intensity, thickness_modifier=thickness, blur_tolerance=blur, upsample_tolerance=ups,
              noise_filter_tolerance=noise, reversed_z=reversed_z)
    o = Oracle(W, H, variant="nofma", **kw)
    depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=seed), reversed_z=reversed_z)
    ao = o.run(depth)
    rc = {k: o.render_constants(k) for k in range(1, 5)}
    uc = {k: o.upsample_constants(k) for k in range(1, 5)}
    r = DF.run(depth, rc, uc, o.zbuffer_params(), reversed_z=reversed_z, return_all=True)
    for k in range(1, 5):
        assert np.array_equal(r["occ"][k], o.codes(9 + k)), f"Occlusion{k}"
    for k in range(1, 4):
        assert np.array_equal(r["comb"][k], o.codes(13 + k)), f"Combined{k}"
    assert np.array_equal(r["comb"][0], ao)


@settings(max_examples=15, deadline=None)
@given(W=st.integers(1, 60), H=st.integers(1, 60), seed=st.integers(0, 10_000), mask=st.integers(0, 15),
       exhaustive=st.booleans(), stereo=st.booleans(), intensity=st.floats(0.0, 2.0), thickness=st.floats(1.0, 10.0))
def test_two_restatements_agree_on_the_shader_variants(W, H, seed, mask, exhaustive, stereo, intensity, thickness):
    """Render.compute kernel `main` (WIDE_SAMPLING), SAMPLE_EXHAUSTIVELY, Upsample.compute main_premin* and the stereo
    thickness: thread-group restatement == global formulation on every AO buffer incl. HighQuality1..4."""
    kw = dict(intensity=intensity, thickness_modifier=thickness, high_quality_mask=mask, sample_exhaustively=exhaustive,
              single_pass_stereo=stereo)
    o = Oracle(W, H, variant="nofma", **kw)
    depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=seed))
    ao = o.run(depth)
    rc = {k: o.render_constants(k) for k in range(1, 5)}
    rcw = {k: o.render_constants(k, wide=True) for k in range(1, 5)}
    uc = {k: o.upsample_constants(k) for k in range(1, 5)}
    r = DF.run(depth, rc, uc, o.zbuffer_params(), return_all=True, exhaustive=exhaustive, high_quality_mask=mask, render_consts_wide=rcw)
    for k in range(1, 5):
        assert np.array_equal(r["occ"][k], o.codes(9 + k)), f"Occlusion{k}"
        if (mask >> (k - 1)) & 1:
            assert np.array_equal(r["hq"][k], o.codes(17 + k)), f"HighQuality{k}"
    for k in range(1, 4):
        assert np.array_equal(r["comb"][k], o.codes(13 + k)), f"Combined{k}"
    assert np.array_equal(r["comb"][0], ao)
