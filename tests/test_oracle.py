"""CPU tests of the oracle itself (no GPU): storage conversions, CPU-side constants against the values
derived from AmbientOcclusion.cs in SURVEY.md 8(a), the analytic known answers (P5), agreement with an
independent second restatement, thread-count invariance, and the committed golden fixtures.

PARITY UNPINNED: the reference has no golden vectors; these are the pins that exist."""
import glob
import os

import numpy as np
import pytest

from miniengineao_b200 import synth
from oracle import direct_formulation as DF
from oracle.oracle import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- storage conversions -------------------------------------------------------------------------
def test_f16_round_matches_numpy_rtne():
    o = Oracle(8, 8)
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        rng.uniform(-70000, 70000, 2000), rng.uniform(-1, 1, 2000), 10.0 ** rng.uniform(-9, 5, 2000),
        np.array([0.0, -0.0, 65504, 65519.99, 65520, 65536, 1e5, 2.0 ** -24, 2.0 ** -25, 2.0 ** -25 * 1.0001, 2.0 ** -14, 6.1e-5,
                  1.0009765625, 1.00048828125, 1.000488281251, 1.00146484375, np.inf, -np.inf]),
    ]).astype(np.float32)
    with np.errstate(over="ignore"):
        ref = vals.astype(np.float16)
    got = np.array([o.f16_bits(float(v)) for v in vals], np.uint16)
    assert np.array_equal(got, ref.view(np.uint16))
    assert np.isnan(o.f16_round(float("nan")))


def test_unorm8_store_convention():
    o = Oracle(8, 8)
    assert o.unorm8_code(float("nan")) == 0                      # NaN -> 0
    assert o.unorm8_code(-3.0) == 0 and o.unorm8_code(7.0) == 255
    for k in range(256):
        assert o.unorm8_code(float(np.float32(k) * np.float32(1.0 / 255.0))) == k      # load(store(k)) round-trips
    assert o.unorm8_code(0.5) == 128                             # 127.5 + 0.5 truncates to 128


# ---- CPU-side constants (AmbientOcclusion.cs:561-593, 660-734, 750-771) -------------------------------
def test_sample_thickness_and_weights_match_survey_table():
    o = Oracle(3840, 2160)
    t = o.sample_thickness()
    ref_t = [0.9797959, 0.9165151, 0.8, 0.6, 0.9591663, 0.8944272, 0.7745966, 0.5656854, 0.8246211, 0.6928203, 0.4472135, 0.5291502]
    assert np.allclose(t, ref_t, atol=2e-7)
    rc = o.render_constants(1)
    ref_w = [0, .1461031, 0, .0956469, .1529021, 0, .2469592, 0, .1314541, 0, .1425820, .0843526]
    assert np.allclose(rc["sample_weight"], ref_w, atol=1e-7)
    assert abs(float(rc["sample_weight"].astype(np.float64).sum()) - 1.0) < 2e-7          # P5(iii)
    assert rc["reject_fadeoff"] == np.float32(-1.0) and rc["intensity"] == np.float32(1.0)
    # InvThicknessTable[i] = (1/TM)/thickness, TM = 2*tanHalfFovH*10/source.width  (AO.cs:678-688)
    tm = np.float32(2) * np.float32(o.camera.tan_half_fov_h) * np.float32(10) / np.float32(480)
    assert np.allclose(rc["inv_thickness"], (np.float32(1) / tm) / t, rtol=1e-6)


def test_upsample_constants_defaults():
    o = Oracle(3840, 2160)
    u1, u4 = o.upsample_constants(1), o.upsample_constants(4)
    assert u1["step_size"] == np.float32(1.0) and u4["step_size"] == np.float32(8.0)     # 1920 / lowRes.width
    assert u1["upsample_tolerance"] == np.float32(1e-12) and u1["noise_filter_strength"] == np.float32(1.0)
    assert abs(float(u1["blur_tolerance"]) - 0.99994981) < 1e-7
    assert abs(float(u4["blur_tolerance"]) - 0.99959821) < 1e-7


def test_zbuffer_params():
    o = Oracle(64, 64, near=0.3, far=100.0)
    zb = o.zbuffer_params()
    assert zb[0] == np.float32(np.float32(100.0) / np.float32(0.3)) - np.float32(1) and zb[1] == 1
    o2 = Oracle(64, 64, near=0.3, far=100.0, reversed_z=False)
    zb2 = o2.zbuffer_params()
    fpn = np.float32(100.0) / np.float32(0.3)
    assert zb2[0] == np.float32(1) - fpn and zb2[1] == fpn


def test_level_dims_ceil():
    o = Oracle(1920, 1080)
    assert [o.level_dims(l) for l in range(7)] == [(1920, 1080), (960, 540), (480, 270), (240, 135), (120, 68), (60, 34), (30, 17)]


# ---- analytic known answers ----------------------------------------------------------
@pytest.mark.parametrize("W,H", [(256, 256), (192, 128), (320, 192)])
def test_constant_depth_is_all_255(W, H):
    """P5(i).  Exact when no atlas has padding texels (sizes that are multiples of 64); with ragged sizes a
    zero-depth padding tap is rejected only if invThickness - 0.5 >= 1, which holds (after unorm8 rounding)
    at 1080p and above (checked on the GPU at 1080p/4K/8K) -- small ragged frames such as 480x270 legitimately deviate."""
    o = Oracle(W, H)
    ao = o.run(synth.lin01_to_raw(np.full((H, W), 0.1, np.float32)))
    assert int((ao != 255).sum()) == 0
    for bid in range(10, 17):
        assert int((o.codes(bid) != 255).sum()) == 0


def test_intensity_zero_is_all_255():
    o = Oracle(200, 120, intensity=0.0)
    assert int((o.run(synth.lin01_to_raw(synth.random_depth(200, 120, 3))) != 255).sum()) == 0


def test_point_sampled_mips_and_deinterleave():
    """LowDepth<k>(i,j) = lin(2^k i, 2^k j) (DS1:64-66, DS2:35); TiledDepth<k> slice (x&3 | (y&3)<<2)."""
    W, H = 200, 120
    o = Oracle(W, H)
    depth = synth.lin01_to_raw(synth.random_depth(W, H, 4))
    o.downsample(depth)
    zb = o.zbuffer_params()
    lin = (np.float32(1) / (depth * zb[0] + zb[1])).astype(np.float32)     # may differ from the fused mad by 1 ulp
    for k in range(1, 5):
        low = o.buffer(1 + k)
        assert low.shape == (o.level_dims(k)[1], o.level_dims(k)[0])
        assert np.allclose(low, lin[::1 << k, ::1 << k], rtol=3e-7)
        t = o.buffer(5 + k)
        assert t.shape == (16, o.level_dims(k + 2)[1], o.level_dims(k + 2)[0])
        for s in (0, 5, 15):
            sub = low[(s >> 2)::4, (s & 3)::4]
            with np.errstate(over="ignore"):
                assert np.array_equal(t[s, :sub.shape[0], :sub.shape[1]], sub.astype(np.float16).astype(np.float32))
    # upsample identity P5(iv): with an exact low/high depth match the result collapses onto hiAO * blurredLowAO
    # (checked implicitly by the bit-exact comparisons below)


def test_padding_texels_of_the_atlases():
    """DS1-written padding = Linearize(0) (reversed-Z: 1e5 -> +inf in f16), DS2-written padding = 0."""
    W, H = 1920 // 8, 1080 // 8 + 3     # 240 x 138: levels 3..6 have ragged bottoms
    o = Oracle(W, H)
    o.downsample(synth.lin01_to_raw(synth.random_depth(W, H, 5)))
    lh1 = o.level_dims(1)[1]
    t1 = o.buffer(6)
    sh = t1.shape[1]
    for s in range(16):
        sy = s >> 2
        valid_rows = len(range(sy, lh1, 4))
        if valid_rows < sh:
            assert np.all(np.isinf(t1[s, valid_rows:, :]))
    t3 = o.buffer(8)
    lh3 = o.level_dims(3)[1]
    for s in range(16):
        valid_rows = len(range(s >> 2, lh3, 4))
        if valid_rows < t3.shape[1]:
            assert np.all(t3[s, valid_rows:, :] == 0)


# ---- independent second restatement -----------------------------------------------------------------------
@pytest.mark.parametrize("W,H,kw", [
    (256, 256, {}), (250, 131, {}), (97, 203, dict(intensity=1.3, thickness_modifier=2.5, blur_tolerance=-3.0, upsample_tolerance=-6.0, noise_filter_tolerance=-2.0)),
    (200, 150, dict(reversed_z=False)), (64, 48, dict(intensity=2.0)), (5, 3, {}), (1, 1, {}),
])
def test_direct_global_formulation_equals_thread_group_oracle(W, H, kw):
    """oracle/direct_formulation.py (global per-pixel formulas, natural layout, virtual atlas, virtual-coordinate
    blur -- the structure the CUDA kernels use) == the literal thread-group restatement, bit for bit (no-FMA builds)."""
    o = Oracle(W, H, variant="nofma", **kw)
    rz = kw.get("reversed_z", True)
    lin = synth.flat_sphere(W, H) if (W, H) == (256, 256) else synth.random_depth(W, H, seed=W + H)
    depth = synth.lin01_to_raw(lin, reversed_z=rz)
    ao = o.run(depth)
    rc = {k: o.render_constants(k) for k in range(1, 5)}
    uc = {k: o.upsample_constants(k) for k in range(1, 5)}
    r = DF.run(depth, rc, uc, o.zbuffer_params(), reversed_z=rz, return_all=True)
    with np.errstate(over="ignore"):
        assert np.array_equal(r["linear"], o.buffer(1))
    dims = DF.level_dims(W, H)
    for k in range(1, 5):
        assert np.array_equal(r["low"][k], o.buffer(1 + k)), f"LowDepth{k}"
        tv = DF.tiled_view(r["low"][k], dims[k + 2][0], dims[k + 2][1], r["pad12"] if k <= 2 else np.float32(0))
        assert np.array_equal(tv, o.buffer(5 + k)), f"TiledDepth{k}"
        assert np.array_equal(r["occ"][k], o.codes(9 + k)), f"Occlusion{k}"
    for k in range(1, 4):
        assert np.array_equal(r["comb"][k], o.codes(13 + k)), f"Combined{k}"
    assert np.array_equal(r["comb"][0], ao)


def test_fma_and_nofma_conventions_differ_rarely():
    """The mad-contraction convention is unpinned (D3D11 'mad' may or may not fuse): quantify its effect."""
    W, H = 480, 270
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    a = Oracle(W, H, intensity=1.1).run(depth)
    b = Oracle(W, H, intensity=1.1, variant="nofma").run(depth)
    diff = np.abs(a.astype(int) - b.astype(int))
    assert diff.max() <= 1 and (diff != 0).mean() < 2e-3


def test_float32_storage_switch_shows_quantisation_share():
    """quantize_storage=0: same math without f16/unorm8 rounding of intermediates."""
    W, H = 256, 256
    depth = synth.lin01_to_raw(synth.flat_sphere(W, H))
    q = Oracle(W, H).run(depth).astype(np.float64) / 255
    f = Oracle(W, H, quantize_storage=False)
    f.run(depth)
    err = np.abs(f.buffer(17).astype(np.float64) - q)
    assert err.max() < 0.08 and err.mean() < 0.01     # storage rounding dominates any plausible float-op discrepancy


def test_thread_count_invariance():
    W, H = 320, 180
    depth = synth.lin01_to_raw(synth.random_depth(W, H, 9))
    a = Oracle(W, H, threads=1).run(depth)
    for t in (2, 3, 8):
        assert np.array_equal(Oracle(W, H, threads=t).run(depth), a)


# ---- committed golden fixtures ----------------------------------------------------------------------------
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))))
def test_oracle_reproduces_golden_fixture(path):
    g = np.load(path)
    H, W = g["depth"].shape
    p = g["params"]
    v = np.zeros(4, np.int32)
    if "variants" in g.files:
        v[:len(g["variants"])] = g["variants"]
    o = Oracle(W, H, noise_filter_tolerance=float(p[0]), blur_tolerance=float(p[1]), upsample_tolerance=float(p[2]),
               thickness_modifier=float(p[3]), intensity=float(p[4]),
               single_pass_stereo=bool(v[0]), sample_exhaustively=bool(v[1]), high_quality_mask=int(v[2]), single_scale=bool(v[3]))
    ao = o.run(g["depth"])
    assert np.array_equal(ao, g["ao"])
    for bid in [int(k[3:]) for k in g.files if k.startswith("buf")]:
        ref = g[f"buf{bid}"]
        if ref.dtype == np.uint8:
            assert np.array_equal(o.codes(bid), ref), bid
        elif ref.dtype == np.float16:
            with np.errstate(over="ignore"):
                assert np.array_equal(o.buffer(bid).astype(np.float16).view(np.uint16), ref.view(np.uint16)), bid
        else:
            assert np.array_equal(o.buffer(bid), ref), bid


def test_composite_oracle_identities():
    """ao = 255 leaves every target unchanged (x * 1); ao = 0 zeroes the scaled channels; untouched channels survive."""
    from oracle import oracle as O
    rng = np.random.default_rng(1)
    c8 = rng.integers(0, 256, size=(7, 9, 4), dtype=np.uint8)
    c16 = rng.uniform(0, 8, size=(7, 9, 4)).astype(np.float16)
    one, zero = np.full((7, 9), 255, np.uint8), np.zeros((7, 9), np.uint8)
    assert np.array_equal(O.composite_framebuffer(one, c8), c8) and np.array_equal(O.composite_framebuffer(one, c16), c16)
    assert not O.composite_framebuffer(zero, c8).any() and not O.composite_framebuffer(zero, c16).any()
    g0, g3 = O.composite_gbuffer(zero, c8, c16)
    assert np.array_equal(g0[..., :3], c8[..., :3]) and not g0[..., 3].any()
    assert np.array_equal(g3[..., 3], c16[..., 3]) and not g3[..., :3].any()
    g0, g3 = O.composite_gbuffer(one, c8, c8)
    assert np.array_equal(g0, c8) and np.array_equal(g3, c8)


# ---- variants the reference ships but never dispatches ---------------------------------
def _run_both(W, H, seed, **kw):
    o = Oracle(W, H, variant="nofma", **kw)
    depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=seed))
    ao = o.run(depth)
    rc = {k: o.render_constants(k) for k in range(1, 5)}
    rcw = {k: o.render_constants(k, wide=True) for k in range(1, 5)}
    uc = {k: o.upsample_constants(k) for k in range(1, 5)}
    r = DF.run(depth, rc, uc, o.zbuffer_params(), return_all=True, exhaustive=kw.get("sample_exhaustively", False),
               high_quality_mask=kw.get("high_quality_mask", 0), render_consts_wide=rcw)
    return o, ao, r


@pytest.mark.parametrize("W,H,kw", [
    (250, 131, dict(sample_exhaustively=True)),
    (97, 203, dict(high_quality_mask=0b1111)),
    (130, 70, dict(high_quality_mask=0b0101, sample_exhaustively=True, intensity=1.2)),
    (64, 48, dict(high_quality_mask=0b1000, single_pass_stereo=True)),
    (33, 17, dict(high_quality_mask=0b0011, thickness_modifier=3.0)),
    (5, 3, dict(high_quality_mask=0b1111, sample_exhaustively=True)),
])
def test_variants_direct_formulation_equals_thread_group_oracle(W, H, kw):
    """Render.compute kernel `main` (WIDE_SAMPLING, 32x32 LDS tile, 16x16 threads), SAMPLE_EXHAUSTIVELY and the
    Upsample.compute main_premin* kernels: global per-pixel restatement == literal thread-group restatement."""
    o, ao, r = _run_both(W, H, W * 3 + H, **kw)
    for k in range(1, 5):
        assert np.array_equal(r["occ"][k], o.codes(9 + k)), f"Occlusion{k}"
        if (kw.get("high_quality_mask", 0) >> (k - 1)) & 1:
            assert np.array_equal(r["hq"][k], o.codes(17 + k)), f"HighQuality{k}"
    for k in range(1, 4):
        assert np.array_equal(r["comb"][k], o.codes(13 + k)), f"Combined{k}"
    assert np.array_equal(r["comb"][0], ao)


def test_exhaustive_weights_and_stereo_thickness():
    """AO.cs:696-724 without the zeroing of :709-715: all twelve weights, still normalised; AO.cs:679-680 double the
    thickness (halve InvThicknessTable) for a non-tiled source and again for single-pass stereo."""
    W, H = 3840, 2160
    base, exh = Oracle(W, H), Oracle(W, H, sample_exhaustively=True)
    w = exh.render_constants(1)["sample_weight"]
    t = exh.sample_thickness()
    mult = np.array([4, 4, 4, 4, 4, 8, 8, 8, 4, 8, 8, 4], np.float32)
    assert np.all(w > 0) and abs(float(w.astype(np.float64).sum()) - 1.0) < 3e-7
    assert np.allclose(w, mult * t / np.float32((mult * t).astype(np.float64).sum()), rtol=1e-6)
    assert np.array_equal(exh.render_constants(1)["inv_thickness"], base.render_constants(1)["inv_thickness"])
    st = Oracle(W, H, single_pass_stereo=True)
    assert np.array_equal(st.render_constants(2)["inv_thickness"] * np.float32(2), base.render_constants(2)["inv_thickness"])
    # wide: source = LowDepth<k> (width lw[k] = 4 * lw[k+2]) and x2 for !isTiled => TM is half the tiled one
    a, b = base.render_constants(3, wide=True), base.render_constants(3)
    assert np.allclose(a["inv_thickness"], b["inv_thickness"] * np.float32(2), rtol=1e-6)
    assert a["inv_slice_dim"][0] == np.float32(1.0 / 480) and b["inv_slice_dim"][0] == np.float32(1.0 / 120)


@pytest.mark.parametrize("W,H", [(256, 256), (192, 128)])
def test_variants_constant_depth_is_all_255(W, H):
    """P5(i) carries over: constant depth => every pair = 1 in either sample set and either kernel; min(255, 255) = 255."""
    o = Oracle(W, H, sample_exhaustively=True, high_quality_mask=15)
    ao = o.run(synth.lin01_to_raw(np.full((H, W), 0.1, np.float32)))
    assert int((ao != 255).sum()) == 0
    for bid in list(range(10, 17)) + [18, 19, 20, 21]:
        assert int((o.codes(bid) != 255).sum()) == 0, bid


def test_premin_never_brightens():
    """LoResAO1 = min(LoResAO1, LoResAO2) feeds a blur + upsample that are monotone in the low-res AO
    (non-negative weights), so enabling the high-quality pass can only darken the combined buffers."""
    W, H = 200, 120
    depth = synth.lin01_to_raw(synth.random_depth(W, H, 12))
    a, b = Oracle(W, H), Oracle(W, H, high_quality_mask=15)
    a.run(depth), b.run(depth)
    for bid in (14, 15, 16, 17):
        assert np.all(b.codes(bid).astype(int) <= a.codes(bid).astype(int) + 1), bid     # +1: unorm8 rounding of a ratio of sums
    assert (b.codes(17).astype(int) < a.codes(17).astype(int)).mean() > 0.05


@pytest.mark.parametrize("W,H", [(256, 256), (130, 70), (83, 61)])
def test_debug_views_two_restatements(W, H):
    """PushDebugBlitCommands (AO.cs:787-820) + Blit.shader pass 4: C loop == vectorised numpy; structural checks."""
    o = Oracle(W, H, intensity=1.1)
    o.run(synth.lin01_to_raw(synth.random_depth(W, H, 2)))
    for bid in range(1, 18):
        assert np.array_equal(o.debug_view(bid), DF.debug_view(o.buffer(bid), W, H)), bid
    assert np.array_equal(o.debug_view(17), o.codes(17))                         # _debug == 17 shows _result itself
    v = o.debug_view(10)                                                          # half-res R8 source: every texel shown 2 x 2
    if W % 2 == 0 and H % 2 == 0:
        assert np.array_equal(v[::2, ::2], o.codes(10)) and np.array_equal(v[1::2, 1::2], o.codes(10))
    if (W, H) == (256, 256):                                                      # de-tile: 4 x 4 mosaic of 64 x 64 cells showing 32 x 32 slices 2 x 2
        t = o.buffer(6)
        m = o.debug_view(6)
        for s in (0, 7, 15):
            cell = m[64 * (s >> 2): 64 * (s >> 2) + 64, 64 * (s & 3): 64 * (s & 3) + 64]
            assert np.array_equal(cell[::2, ::2], DF._unorm8(t[s]))


# ---- round 2: the single-scale plan (BASELINE.json configs[0]) and the unpinned-convention switches -------------------------
@pytest.mark.parametrize("W,H", [(256, 256), (97, 203), (5, 3)])
def test_single_scale_two_restatements_agree(W, H):
    """Downsample -> Render level 1 -> final-style Upsample with LoResAO1 = Occlusion1: thread-group C restatement == global
    numpy formulation, bit for bit (no-FMA builds), and the plan really ignores the coarser levels."""
    lin = synth.flat_sphere(W, H) if (W, H) == (256, 256) else synth.random_depth(W, H, seed=W + H)
    depth = synth.lin01_to_raw(lin)
    o = Oracle(W, H, variant="nofma", intensity=1.1, single_scale=True)
    ao = o.run(depth)
    rc = {k: o.render_constants(k) for k in range(1, 5)}
    uc = {k: o.upsample_constants(k) for k in range(1, 5)}
    r = DF.run(depth, rc, uc, o.zbuffer_params(), return_all=True, single_scale=True)
    assert np.array_equal(r["occ"][1], o.codes(10))
    assert np.array_equal(r["comb"][0], ao)
    # stage-wise: the same as running the three stages by hand on a fresh oracle
    m = Oracle(W, H, variant="nofma", intensity=1.1, single_scale=True)
    m.downsample(depth); m.render(1); m.upsample(1)
    assert np.array_equal(m.ao_u8(), ao)
    full = Oracle(W, H, variant="nofma", intensity=1.1).run(depth)
    if W >= 64:
        assert not np.array_equal(full, ao)          # the multi-scale result is darker: coarser levels multiply in
        assert ao.astype(int).sum() >= full.astype(int).sum()


def test_single_scale_known_answers():
    """P5 identities hold for the single-scale plan too: constant depth -> 255 (no padded level involved: 256 = 4 * 64), intensity 0 -> 255."""
    W = H = 256
    flat = synth.lin01_to_raw(np.full((H, W), 0.25, np.float32))
    assert (Oracle(W, H, single_scale=True).run(flat) == 255).all()
    d = synth.lin01_to_raw(synth.flat_sphere(W, H))
    assert (Oracle(W, H, single_scale=True, intensity=0.0).run(d) == 255).all()
    a = Oracle(W, H, single_scale=True, intensity=1.1).run(d)
    assert a.min() < 200 and (a == 255).mean() > 0.3     # the sphere's contact shadow is there, the open plane is unoccluded


def test_unpinned_convention_switches_flip_rates():
    """how often each LEGAL alternative to a convention the reference does not pin changes an output
    code.  The mad and division freedoms move <= 1 code on a fraction of a percent; the f16 store rounding is the one that
    matters (RTZ vs RTNE moves up to several codes on ~10 % of the pixels) -- a D3D11 capture would have to settle that one."""
    W, H = 640, 360
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    ref = Oracle(W, H, threads=4, intensity=1.1).run(depth).astype(int)
    rates = {}
    for v in ("nofma", "divmulrcp", "divrtz", "f16rtz"):
        d = np.abs(Oracle(W, H, threads=4, intensity=1.1, variant=v).run(depth).astype(int) - ref)
        rates[v] = ((d != 0).mean(), int(d.max()))
    assert rates["nofma"][1] <= 1 and rates["nofma"][0] < 2e-3
    assert rates["divmulrcp"][1] <= 1 and rates["divmulrcp"][0] < 2e-3
    assert rates["divrtz"][1] <= 1 and rates["divrtz"][0] < 1e-2
    assert 0.01 < rates["f16rtz"][0] < 0.3 and rates["f16rtz"][1] <= 12


def test_f16_rtz_store_convention():
    rtz = Oracle(8, 8, variant="f16rtz")
    rne = Oracle(8, 8)
    assert rne.f16_bits(1e5) == 0x7c00 and rtz.f16_bits(1e5) == 0x7bff           # overflow: inf vs largest finite
    assert rtz.f16_bits(float("inf")) == 0x7c00
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-70000, 70000, 2000), rng.uniform(-1e-4, 1e-4, 2000), rng.uniform(-1e-7, 1e-7, 500)]).astype(np.float32)
    for v in x:
        b = rtz.f16_bits(float(v))
        back = np.array([b], np.uint16).view(np.float16)[0].astype(np.float32)
        assert abs(back) <= abs(v)                                               # truncation never grows the magnitude
        n = rne.f16_bits(float(v))
        assert abs(int(b & 0x7fff) - int(n & 0x7fff)) <= 1 or abs(v) > 65504     # and is at most one code below nearest
