"""GPU parity tests: the CUDA path, called through the C ABI (via the ctypes host), against the
CPU oracle on identical inputs.  Bar: BIT-EXACT on every buffer (f16 bits, f32 bits, unorm8
codes) -- north_star's tolerance is 1e-3 absolute on the AO value, and one unorm8 code is 3.9e-3,
so any differing code would already be out of tolerance (TOL_CODES = 0).

The reference publishes no golden vectors (parity unpinned, see oracle/meao_oracle.h); the oracle
is pinned by analytic identities and a second independent restatement in tests/test_oracle.py.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_CODES = 0   # allowed |delta| in unorm8 codes; 1e-3 absolute < 1/255


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("GPU tests need a GPU")
    return torch


def _mk(W, H, **params):
    """CUDA host + oracle with the same parameters.  Variant keys: sample_exhaustively,
    high_quality_mask, single_pass_stereo (W is then the DOUBLE-WIDE eye pair; the camera is W/2 wide)."""
    from miniengineao_b200 import AmbientOcclusion, Camera
    from oracle.oracle import Oracle
    cam_kw = {}
    if "reversed_z" in params:
        cam_kw["usesReversedZBuffer"] = params["reversed_z"]
    stereo = bool(params.get("single_pass_stereo", False))
    cam = Camera(W // 2 if stereo else W, H, stereoEnabled=stereo, **cam_kw)
    ao = AmbientOcclusion(cam, device=0, use_graph=params.pop("use_graph", True))
    ao.sampleExhaustively = bool(params.get("sample_exhaustively", False))
    ao.highQualityMask = int(params.get("high_quality_mask", 0))
    okw = {k: params[k] for k in ("sample_exhaustively", "high_quality_mask", "single_pass_stereo") if k in params}
    if stereo:
        ao.OnPreRender()                    # one draw for both eyes (AO.cs:352-355, 392-401)
        okw["tan_half_fov_h_"] = 1.0 / cam.projection00
    for py, cs in (("noise_filter_tolerance", "noiseFilterTolerance"), ("blur_tolerance", "blurTolerance"),
                   ("upsample_tolerance", "upsampleTolerance"), ("thickness_modifier", "thicknessModifier"),
                   ("intensity", "intensity")):
        if py in params:
            setattr(ao, cs, params[py])
            okw[py] = params[py]
    if "reversed_z" in params:
        okw["reversed_z"] = params["reversed_z"]
    orc = Oracle(W, H, threads=8, **okw)
    return ao, orc


def _compare_all(ao, orc, tag, extra=()):
    """Every debug buffer 1..17 (+ extension ids in `extra`) of the CUDA path against the oracle, bit for bit."""
    bad = []
    for bid in list(range(1, 18)) + list(extra):
        got = ao.debug_buffer(bid)
        ref = orc.buffer(bid)
        if got.dtype == np.uint8:
            refc = orc.codes(bid)
            n = int((np.abs(got.astype(np.int16) - refc.astype(np.int16)) > TOL_CODES).sum())
        elif got.dtype == np.float16:
            n = int((got.view(np.uint16) != ref.astype(np.float16).view(np.uint16)).sum())
        else:
            n = int((got.view(np.uint32) != ref.view(np.uint32)).sum())
        if n:
            bad.append((bid, ao.DEBUG_NAMES[bid], n, got.size))
    assert not bad, f"{tag}: mismatching buffers (id, name, #diff, size): {bad}"


SIZES = [(256, 256), (64, 64), (16, 16), (8, 8), (1, 1), (3, 5), (130, 70), (250, 131), (321, 203), (640, 360), (1000, 37), (37, 1000)]


@pytest.mark.parametrize("W,H", SIZES)
def test_full_pipe_bit_exact_random(torch_cuda, W, H):
    from miniengineao_b200 import synth
    torch = torch_cuda
    ao, orc = _mk(W, H, intensity=1.1)
    depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=W * 7 + H))
    ref = orc.run(depth)
    got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
    assert int((got != ref).sum()) == 0
    _compare_all(ao, orc, f"{W}x{H}")


def test_flat_sphere_256(torch_cuda):
    """BASELINE.json configs[0]: 256x256 synthetic flat+sphere depth."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    ao, orc = _mk(256, 256)
    depth = synth.lin01_to_raw(synth.flat_sphere(256, 256))
    ref = orc.run(depth)
    got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
    assert np.array_equal(got, ref)
    _compare_all(ao, orc, "flat+sphere")


@pytest.mark.parametrize("params", [
    dict(intensity=0.0), dict(intensity=2.0), dict(thickness_modifier=10.0), dict(thickness_modifier=2.5, intensity=1.3),
    dict(blur_tolerance=-1.0), dict(blur_tolerance=-8.0), dict(upsample_tolerance=-1.0), dict(upsample_tolerance=-6.0, noise_filter_tolerance=-8.0),
    dict(noise_filter_tolerance=-3.0, blur_tolerance=-3.0, upsample_tolerance=-4.0, thickness_modifier=4.0, intensity=0.7),
    dict(reversed_z=False),
])
def test_parameter_sweep(torch_cuda, params):
    from miniengineao_b200 import synth
    torch = torch_cuda
    W, H = 322, 190
    ao, orc = _mk(W, H, **params)
    depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=11), reversed_z=params.get("reversed_z", True))
    ref = orc.run(depth)
    got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
    assert np.array_equal(got, ref), params
    _compare_all(ao, orc, str(params))


def test_corridor_1080p(torch_cuda):
    """BASELINE.json configs[1]: 1920x1080 Sponza-like depth, full multi-scale pipe."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    W, H = 1920, 1080
    ao, orc = _mk(W, H, intensity=1.1)
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    ref = orc.run(depth)
    got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
    assert int((got != ref).sum()) == 0
    _compare_all(ao, orc, "1080p corridor")


def test_sky_pixels_and_nan_semantics(torch_cuda):
    """Raw depth 0 (reversed-Z sky) -> 1e5 -> +inf in f16 -> inf*0 = NaN inside the sampler; the HLSL
    saturate/min/max NaN rules must hold on both sides."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    W, H = 200, 120
    ao, orc = _mk(W, H)
    lin = synth.random_depth(W, H, seed=5)
    depth = synth.lin01_to_raw(lin)
    depth[20:60, 30:90] = 0.0           # a sky window
    depth[::17, ::13] = 0.0             # isolated sky pixels
    ref = orc.run(depth)
    got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
    assert np.array_equal(got, ref)
    _compare_all(ao, orc, "sky")


def test_linear_depth_ingest(torch_cuda):
    from miniengineao_b200 import synth
    from oracle.oracle import Oracle
    from miniengineao_b200 import AmbientOcclusion, Camera
    torch = torch_cuda
    W, H = 250, 131
    lin = synth.random_depth(W, H, seed=9)
    ao = AmbientOcclusion(Camera(W, H), device=0)
    orc = Oracle(W, H, depth_is_linear=True)
    ref = orc.run(lin)
    got = ao.render(torch.from_numpy(lin).cuda(), linear=True).cpu().numpy()
    assert np.array_equal(got, ref)
    _compare_all(ao, orc, "linear ingest")


def test_stagewise_with_injected_inputs(torch_cuda):
    """Each stage alone, fed the ORACLE's inputs (meao_set_buffer), like the reference's debug views."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    W, H = 330, 170
    ao, orc = _mk(W, H, intensity=1.2)
    depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=21))
    orc.run(depth)
    d = torch.from_numpy(depth).cuda()
    ao.stage_downsample(d)
    for bid in (1, 2, 3, 4, 5, 6, 7, 8, 9):
        got, ref = ao.debug_buffer(bid), orc.buffer(bid)
        if got.dtype == np.float16:
            assert np.array_equal(got.view(np.uint16), ref.astype(np.float16).view(np.uint16)), bid
        else:
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), bid
    for k in range(1, 5):
        ao.set_debug_buffer(1 + k, orc.buffer(1 + k))
        ao.stage_render(k)
        assert np.array_equal(ao.debug_buffer(9 + k), orc.codes(9 + k)), f"render level {k}"
    for lo in range(4, 0, -1):
        hi = lo - 1
        ao.set_debug_buffer(1 + lo, orc.buffer(1 + lo))
        ao.set_debug_buffer(13 if lo == 4 else 13 + lo, orc.codes(13 if lo == 4 else 13 + lo))
        if hi > 0:
            ao.set_debug_buffer(1 + hi, orc.buffer(1 + hi))
            ao.set_debug_buffer(9 + hi, orc.codes(9 + hi))
        else:
            ao.set_debug_buffer(1, orc.buffer(1).astype(np.float16))
        ao.stage_upsample(lo)
        out_id = 17 if hi == 0 else 13 + hi
        assert np.array_equal(ao.debug_buffer(out_id), orc.codes(out_id)), f"upsample {lo}->{hi}"


# ---- size-independent properties at BASELINE.json's full sizes ---------------------------------------

@pytest.mark.parametrize("W,H", [(1920, 1080), (3840, 2160), (7680, 4320)])
def test_constant_depth_gives_255_everywhere(torch_cuda, W, H):
    """constant depth => every pair = 1, sum of weights = 1 => code 255 at all levels."""
    torch = torch_cuda
    ao, _ = _mk(16, 16)
    ao.camera.pixelWidth, ao.camera.pixelHeight = W, H
    d = torch.full((H, W), 0.02, dtype=torch.float32, device="cuda")
    out = ao.render(d)
    assert int((out != 255).sum().item()) == 0
    for bid in range(10, 17):
        assert int((ao.debug_buffer(bid) != 255).sum()) == 0, bid


@pytest.mark.parametrize("W,H", [(3840, 2160)])
def test_intensity_zero_gives_255(torch_cuda, W, H):
    """intensity 0 => lerp(1, ao, 0) = 1 for any depth."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    ao, _ = _mk(W, H, intensity=0.0)
    d = torch.from_numpy(synth.lin01_to_raw(synth.corridor(W, H))).cuda()
    out = ao.render(d)
    assert int((out != 255).sum().item()) == 0


def test_4k_matches_oracle_on_crops_and_checksum(torch_cuda):
    """4K (the metric's config): full oracle comparison (the C oracle does 4K in a few seconds with 8 threads)."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    W, H = 3840, 2160
    ao, orc = _mk(W, H, intensity=1.1)
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    ref = orc.run(depth)
    got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
    assert int((got != ref).sum()) == 0
    assert int(got.astype(np.uint64).sum()) == int(ref.astype(np.uint64).sum())


def test_graph_replay_equals_stream_launch_and_is_deterministic(torch_cuda):
    from miniengineao_b200 import synth
    torch = torch_cuda
    W, H = 1920, 1080
    depth = torch.from_numpy(synth.lin01_to_raw(synth.corridor(W, H))).cuda()
    a1, _ = _mk(W, H)
    a2, _ = _mk(W, H, use_graph=False)
    o1 = a1.render(depth).clone()
    o1b = a1.render(depth).clone()       # second call replays the captured graph
    o2 = a2.render(depth)
    torch.cuda.synchronize()
    assert torch.equal(o1, o1b) and torch.equal(o1, o2)
    assert a1.launch_count == 18 and a2.launch_count == 9


def test_replan_on_parameter_and_size_change(torch_cuda):
    """LateUpdate semantics (AO.cs:329-350): rebuild only when a property or the size changed."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    ao, orc = _mk(128, 96)
    d = torch.from_numpy(synth.lin01_to_raw(synth.random_depth(128, 96, 3))).cuda()
    ao.render(d)
    n = ao.rebuild_count
    ao.render(d)
    assert ao.rebuild_count == n
    ao.intensity = 0.5
    out = ao.render(d).cpu().numpy()
    assert ao.rebuild_count == n + 1
    orc.params.intensity = 0.5
    assert np.array_equal(out, orc.run(d.cpu().numpy()))
    ao.camera.pixelWidth, ao.camera.pixelHeight = 96, 128
    d2 = synth.lin01_to_raw(synth.random_depth(96, 128, 4))
    out2 = ao.render(torch.from_numpy(d2).cuda()).cpu().numpy()
    assert ao.rebuild_count == n + 2
    from oracle.oracle import Oracle
    assert np.array_equal(out2, Oracle(96, 128, intensity=0.5).run(d2))


def test_host_buffer_path_and_event_hook(torch_cuda):
    import ctypes as C
    from miniengineao_b200 import synth, _native as N
    torch = torch_cuda
    W, H = 320, 200
    ao, orc = _mk(W, H)
    depth = synth.lin01_to_raw(synth.random_depth(W, H, 8))
    ref = orc.run(depth)
    assert np.array_equal(ao.render_host(depth), ref)
    # command-buffer hook: IssuePluginEvent-style callback
    d = torch.from_numpy(depth).cuda()
    out = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
    lib = N.lib()
    st = torch.cuda.Stream()                                   # ABI 3: the event renders on the stream it was bound with
    N.check(ao._ctx, lib.meao_bind_event(ao._ctx, 42, d.data_ptr(), 0, out.data_ptr(), st.cuda_stream))
    fn = lib.meao_get_render_event_func()
    fn(42)
    st.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)
    out.zero_()
    N.check(ao._ctx, lib.meao_bind_event(ao._ctx, 43, d.data_ptr(), 0, out.data_ptr(), None))     # NULL = legacy default stream
    fn(43)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("W,H,bands", [(1280, 720, 2), (1920, 1080, 2), (3840, 2160, 4)])
def test_row_bands_equal_whole_frame(torch_cuda, W, H, bands):
    """Tile-vs-whole equality: each band context computes its rows from its own depth rows
    plus the halo rows of LowDepth1..4 packed by its neighbours; the union must be bit-identical to the
    single-context frame.  All bands live on one GPU here; the exchange is a device copy."""
    from miniengineao_b200 import AmbientOcclusion, Camera, synth
    torch = torch_cuda
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    whole = AmbientOcclusion(Camera(W, H), device=0)
    ref = whole.render(torch.from_numpy(depth).cuda()).cpu().numpy()
    blocks = (H + 15) // 16
    cuts = [min(H, 16 * ((blocks * i) // bands)) for i in range(bands)] + [H]
    ctxs = []
    for i in range(bands):
        a = AmbientOcclusion(Camera(W, H), device=0)
        a.set_row_band(cuts[i], cuts[i + 1], cuts[i - 1] if i > 0 else -1, cuts[i + 2] if i + 2 <= bands else -1)
        ctxs.append(a)
    dts = [torch.from_numpy(depth[cuts[i]:cuts[i + 1]]).cuda() for i in range(bands)]
    for a, d in zip(ctxs, dts):
        a.band_prepare(d)
    torch.cuda.synchronize()
    for i in range(bands):
        for side, j in ((0, i - 1), (1, i + 1)):
            if j < 0 or j >= bands:
                continue
            nbytes = ctxs[i].halo_bytes(side)
            assert nbytes == ctxs[j].halo_recv_bytes(1 - side)
            buf = torch.empty(max(nbytes, 4), dtype=torch.uint8, device="cuda")
            ctxs[i].halo_pack(side, buf)
            ctxs[i].synchronize()
            ctxs[j].halo_unpack(1 - side, buf)
            ctxs[j].synchronize()
    got = np.zeros((H, W), np.uint8)
    for i, a in enumerate(ctxs):
        o = torch.empty((cuts[i + 1] - cuts[i], W), dtype=torch.uint8, device="cuda")
        a.band_finish(o)
        a.synchronize()
        got[cuts[i]:cuts[i + 1]] = o.cpu().numpy()
    assert int((got != ref).sum()) == 0


def test_too_thin_band_is_refused(torch_cuda):
    """A band thinner than the dependency radius (346 L0 rows for Occlusion4) would need a multi-hop
    exchange: libmeao refuses it loudly instead of computing garbage."""
    from miniengineao_b200 import AmbientOcclusion, Camera, MeaoError
    a = AmbientOcclusion(Camera(1920, 1088), device=0)
    with pytest.raises(MeaoError) as e:
        a.set_row_band(272, 544, 0, 816)
    assert e.value.code == -3


def test_fast_division_is_ieee_exact(torch_cuda):
    """The kernels replace nvcc's `a / b` expansion by its own fast path (MUFU.RCP + FMA refinement) under a
    stricter operand guard (csrc/common.cuh).  Brute force: 2^30 random in-range operand pairs, incl. the
    numerators 1, 3, 9 and quotients near 1, must give bit-identical results to the IEEE operators."""
    from miniengineao_b200 import AmbientOcclusion, Camera
    a = AmbientOcclusion(Camera(64, 64), device=0)
    assert a.selftest_div(1 << 30, seed=12345) == 0


def test_pipelined_host_batch_matches_oracle(torch_cuda):
    """meao_render_host_async / meao_host_wait (two staging slots): five distinct frames in flight, pinned buffers."""
    import ctypes as C
    from miniengineao_b200 import synth, _native as N
    W, H = 640, 360
    ao, orc = _mk(W, H, intensity=1.1)
    lib = N.lib()
    n = 5
    ptrs = [(lib.meao_host_alloc(W * H * 4), lib.meao_host_alloc(W * H)) for _ in range(n)]
    try:
        ds = [np.ctypeslib.as_array(C.cast(p[0], C.POINTER(C.c_float)), shape=(H, W)) for p in ptrs]
        os_ = [np.ctypeslib.as_array(C.cast(p[1], C.POINTER(C.c_uint8)), shape=(H, W)) for p in ptrs]
        for i in range(n):
            ds[i][...] = synth.lin01_to_raw(synth.random_depth(W, H, seed=100 + i))
            os_[i][...] = 0
        ao.render_host_batch(ds, os_)
        for i in range(n):
            assert np.array_equal(os_[i], orc.run(ds[i])), i
    finally:
        ao.close()
        for p in ptrs:
            lib.meao_host_free(p[0]); lib.meao_host_free(p[1])


@pytest.mark.parametrize("W,H,bands", [(1920, 1080, 2), (3840, 2160, 4)])
def test_graph_cached_band_phases_equal_whole_frame(torch_cuda, W, H, bands):
    """meao_band_phase_a / meao_band_phase_b (pack / unpack fused in, each half one CUDA graph): same equality,
    run twice so the second pass replays the cached graphs."""
    from miniengineao_b200 import AmbientOcclusion, Camera, rowtile, synth
    torch = torch_cuda
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    ref = AmbientOcclusion(Camera(W, H), device=0).render(torch.from_numpy(depth).cuda()).cpu().numpy()
    cuts = rowtile.partition(H, bands)
    ctxs, send, recv = [], [], []
    for i in range(bands):
        a = AmbientOcclusion(Camera(W, H), device=0)
        a.set_row_band(cuts[i], cuts[i + 1], *rowtile.neighbours(cuts, i))
        ctxs.append(a)
        mk = lambda n: torch.empty(int(n), dtype=torch.uint8, device="cuda")  # noqa: E731
        send.append([mk(a.halo_bytes(0)), mk(a.halo_bytes(1))])
        recv.append([mk(a.halo_recv_bytes(0)), mk(a.halo_recv_bytes(1))])
    dts = [torch.from_numpy(depth[cuts[i]:cuts[i + 1]]).cuda() for i in range(bands)]
    outs = [torch.zeros((cuts[i + 1] - cuts[i], W), dtype=torch.uint8, device="cuda") for i in range(bands)]
    for _ in range(2):
        for o in outs:
            o.zero_()
        for i, a in enumerate(ctxs):
            a.band_phase_a(dts[i], send[i][0], send[i][1])
        torch.cuda.synchronize()
        for i in range(bands):          # the "exchange": what i sends down is what i+1 receives from above
            if i + 1 < bands:
                recv[i + 1][0].copy_(send[i][1])
                recv[i][1].copy_(send[i + 1][0])
        torch.cuda.synchronize()
        for i, a in enumerate(ctxs):
            a.band_phase_b(recv[i][0], recv[i][1], outs[i])
        torch.cuda.synchronize()
        got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
        assert int((got != ref).sum()) == 0


@pytest.mark.parametrize("W,H", [(640, 360), (321, 203)])
def test_composite_passes_bit_exact(torch_cuda, W, H):
    """Blit.shader pass 2 (frame buffer *= ao) and pass 1 (G-buffer occlusion / ambient *= 1-(1-ao))
    on RGBA8 and RGBA16F targets, against the oracle's fp32 restatement of the output-merger blend."""
    from miniengineao_b200 import synth
    from oracle import oracle as O
    torch = torch_cuda
    ao, orc = _mk(W, H, intensity=1.1)
    depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=3))
    ref_ao = orc.run(depth)
    ao_dev = ao.render(torch.from_numpy(depth).cuda())
    rng = np.random.default_rng(0)
    c8 = rng.integers(0, 256, size=(H, W, 4), dtype=np.uint8)
    c16 = (rng.uniform(0, 4, size=(H, W, 4)) ** 3).astype(np.float16)
    for host in (c8, c16):
        dev = torch.from_numpy(host.copy()).cuda()
        ao.composite_framebuffer(ao_dev, dev)
        torch.cuda.synchronize()
        exp = O.composite_framebuffer(ref_ao, host)
        assert np.array_equal(dev.cpu().numpy().view(np.uint8), exp.view(np.uint8)), host.dtype
    for g3 in (c8, c16):
        g0d, g3d = torch.from_numpy(c8.copy()).cuda(), torch.from_numpy(g3.copy()).cuda()
        ao.composite_gbuffer(ao_dev, g0d, g3d)
        torch.cuda.synchronize()
        e0, e3 = O.composite_gbuffer(ref_ao, c8, g3)
        assert np.array_equal(g0d.cpu().numpy(), e0)
        assert np.array_equal(g3d.cpu().numpy().view(np.uint8), e3.view(np.uint8)), g3.dtype


@pytest.mark.parametrize("W,H", [(640, 360), (250, 131)])
def test_native_depth_formats(torch_cuda, W, H):
    """D16_UNORM and D24_UNORM_S8_UINT ingest (what Blit.shader pass 0 samples).  The oracle sees
    the float the D3D UNORM->FLOAT rule produces, code * (1 / (2^n - 1))."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    lin = synth.random_depth(W, H, seed=17)
    raw = synth.lin01_to_raw(lin).astype(np.float64)
    for bits, dt in ((16, np.uint16), (24, np.uint32)):
        ao, orc = _mk(W, H, intensity=1.1)
        full = (1 << bits) - 1
        codes = np.clip(np.rint(raw * full), 1, full).astype(np.uint32)          # no sky codes
        as_float = (codes.astype(np.float32) * np.float32(1.0 / full)).astype(np.float32)
        ref = orc.run(as_float)
        if bits == 16:
            dev = torch.from_numpy(codes.astype(np.uint16).view(np.int16)).cuda().view(torch.uint16)
            host = codes.astype(np.uint16)
        else:
            words = codes | (np.uint32(0xA5) << np.uint32(24))                    # stencil bits must be ignored
            dev = torch.from_numpy(words.view(np.int32)).cuda()
            host = words
        got = ao.render(dev).cpu().numpy()
        assert np.array_equal(got, ref), bits
        _compare_all(ao, orc, f"D{bits}")
        assert np.array_equal(ao.render_host(host), ref), bits


def test_pure_c_client_renders_golden_frame(torch_cuda, tmp_path):
    """The drop-in boundary from plain C: tests/c_abi/smoke.c renders a committed golden fixture through
    meao_render_host and must reproduce its AO bytes."""
    import os, subprocess
    from test_abi import _build_c_client
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "corridor_160x90.npz"))
    H, W = g["depth"].shape
    dpath, apath = os.path.join(str(tmp_path), "depth.f32"), os.path.join(str(tmp_path), "ao.u8")
    g["depth"].astype(np.float32).tofile(dpath)
    g["ao"].astype(np.uint8).tofile(apath)
    exe = _build_c_client(tmp_path)
    r = subprocess.run([exe, "render", str(W), str(H), dpath, apath, repr(float(g["params"][4]))], capture_output=True, text=True)
    assert r.returncode == 0 and " 0 mismatching pixels" in r.stdout, r.stdout + r.stderr


def test_random_configurations_bit_exact(torch_cuda):
    """32 seeded random (size, parameter, camera) configurations, full pipe + every intermediate buffer."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    rng = np.random.default_rng(20260923)
    for i in range(32):
        W, H = int(rng.integers(1, 420)), int(rng.integers(1, 300))
        params = dict(intensity=float(rng.uniform(0, 2)), thickness_modifier=float(rng.uniform(1, 10)),
                      blur_tolerance=float(rng.uniform(-8, -1)), upsample_tolerance=float(rng.uniform(-12, -1)),
                      noise_filter_tolerance=float(rng.uniform(-8, 0)), reversed_z=bool(rng.integers(0, 2)))
        ao, orc = _mk(W, H, **params)
        depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=int(rng.integers(0, 1 << 30))), reversed_z=params["reversed_z"])
        if i % 4 == 0 and W > 8 and H > 8:
            depth[H // 3: H // 3 + 3, W // 4: W // 4 + 5] = 0.0 if params["reversed_z"] else 1.0      # a patch of sky
        ref = orc.run(depth)
        got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
        assert np.array_equal(got, ref), (i, W, H, params)
        _compare_all(ao, orc, f"cfg {i}: {W}x{H} {params}")
        ao.close()


# ---- variants the reference ships but never dispatches ---------------------------------------
def _hq_ids(mask):
    return [17 + k for k in range(1, 5) if (mask >> (k - 1)) & 1]


VARIANTS = [
    dict(sample_exhaustively=True),
    dict(high_quality_mask=0b1111),
    dict(high_quality_mask=0b1000, intensity=1.1),
    dict(high_quality_mask=0b0110, sample_exhaustively=True, thickness_modifier=2.5, intensity=1.3),
    dict(single_pass_stereo=True),
    dict(single_pass_stereo=True, high_quality_mask=0b1111, sample_exhaustively=True, reversed_z=False),
]


@pytest.mark.parametrize("W,H", [(256, 256), (130, 70), (322, 190), (640, 360), (1000, 38), (38, 1000), (6, 4)])
@pytest.mark.parametrize("variant", VARIANTS)
def test_variants_full_pipe_bit_exact(torch_cuda, W, H, variant):
    """Render.compute kernel `main` (WIDE_SAMPLING) -> HighQuality<k>, SAMPLE_EXHAUSTIVELY, Upsample.compute main_premin*
    and the single-pass-stereo thickness: whole pipe + every intermediate against the oracle, bit for bit."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    ao, orc = _mk(W, H, **variant)
    rz = variant.get("reversed_z", True)
    depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=W * 5 + H), reversed_z=rz)
    ref = orc.run(depth)
    got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
    mask = variant.get("high_quality_mask", 0)
    assert ao.kernels_per_frame == 9 + bin(mask).count("1") and ao.launch_count == ao.kernels_per_frame
    assert int((got != ref).sum()) == 0, variant
    _compare_all(ao, orc, f"{W}x{H} {variant}", extra=_hq_ids(mask))


def test_variants_1080p_corridor_and_sky(torch_cuda):
    """BASELINE.json configs[1] geometry with every variant on, plus sky pixels (inf / NaN semantics in the f32 wide path)."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    W, H = 1920, 1080
    ao, orc = _mk(W, H, intensity=1.1, high_quality_mask=15, sample_exhaustively=True)
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    depth[100:180, 300:700] = 0.0
    depth[::37, ::29] = 0.0
    ref = orc.run(depth)
    got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
    assert int((got != ref).sum()) == 0
    _compare_all(ao, orc, "1080p variants", extra=[18, 19, 20, 21])


def test_variants_4k_matches_oracle(torch_cuda):
    from miniengineao_b200 import synth
    torch = torch_cuda
    W, H = 3840, 2160
    ao, orc = _mk(W, H, intensity=1.1, high_quality_mask=0b1100)
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    ref = orc.run(depth)
    got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
    assert int((got != ref).sum()) == 0
    for bid in (20, 21):
        assert np.array_equal(ao.debug_buffer(bid), orc.codes(bid)), bid


def test_stage_render_wide_and_premin_with_injected_inputs(torch_cuda):
    """The two new stage entry points alone, fed the ORACLE's inputs: meao_stage_render_wide per level, and
    meao_stage_upsample running the premin kernels (main_premin_blendout for 4->3..2->1, main_premin for 1->0)."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    W, H = 330, 170
    ao, orc = _mk(W, H, intensity=1.2, high_quality_mask=15)
    depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=31))
    orc.run(depth)
    ao.stage_downsample(torch.from_numpy(depth).cuda())
    for k in range(1, 5):
        ao.set_debug_buffer(1 + k, orc.buffer(1 + k))
        ao.stage_render_wide(k)
        assert np.array_equal(ao.debug_buffer(17 + k), orc.codes(17 + k)), f"render_wide level {k}"
    for lo in range(4, 0, -1):
        hi = lo - 1
        ao.set_debug_buffer(1 + lo, orc.buffer(1 + lo))
        ao.set_debug_buffer(13 if lo == 4 else 13 + lo, orc.codes(13 if lo == 4 else 13 + lo))
        ao.set_debug_buffer(17 + lo, orc.codes(17 + lo))
        if hi > 0:
            ao.set_debug_buffer(1 + hi, orc.buffer(1 + hi))
            ao.set_debug_buffer(9 + hi, orc.codes(9 + hi))
        else:
            ao.set_debug_buffer(1, orc.buffer(1).astype(np.float16))
        ao.stage_upsample(lo)
        out_id = 17 if hi == 0 else 13 + hi
        assert np.array_equal(ao.debug_buffer(out_id), orc.codes(out_id)), f"premin upsample {lo}->{hi}"


def test_variant_switch_replans_and_graph_counts(torch_cuda):
    """Toggling a variant between frames re-plans (drops the captured graph) and the next frame matches the oracle."""
    from miniengineao_b200 import synth
    from oracle.oracle import Oracle
    torch = torch_cuda
    W, H = 320, 200
    depth = synth.lin01_to_raw(synth.random_depth(W, H, 77))
    d = torch.from_numpy(depth).cuda()
    ao, _ = _mk(W, H)
    seq = [dict(), dict(high_quality_mask=5), dict(high_quality_mask=5, sample_exhaustively=True), dict(sample_exhaustively=True), dict()]
    launches = 0
    for v in seq:
        ao.highQualityMask = v.get("high_quality_mask", 0)
        ao.sampleExhaustively = v.get("sample_exhaustively", False)
        for _ in range(2):          # second call replays the graph
            got = ao.render(d).cpu().numpy()
            launches += 9 + bin(ao.highQualityMask).count("1")
        assert np.array_equal(got, Oracle(W, H, threads=8, **v).run(depth)), v
    assert ao.launch_count == launches


@pytest.mark.parametrize("W,H,bands", [(1920, 1080, 2), (2560, 1440, 3)])
def test_row_bands_with_variants_equal_whole_frame(torch_cuda, W, H, bands):
    """Tile-vs-whole equality with the high-quality passes on: HighQuality<k> needs only +-8 rows of LowDepth<k>,
    inside the halo the interleaved render already exchanges."""
    from miniengineao_b200 import AmbientOcclusion, Camera, rowtile, synth
    torch = torch_cuda
    depth = synth.lin01_to_raw(synth.corridor(W, H))

    def mk():
        a = AmbientOcclusion(Camera(W, H), device=0)
        a.highQualityMask, a.sampleExhaustively = 15, True
        return a
    ref = mk().render(torch.from_numpy(depth).cuda()).cpu().numpy()
    cuts = rowtile.partition(H, bands)
    ctxs, send, recv = [], [], []
    for i in range(bands):
        a = mk()
        a.set_row_band(cuts[i], cuts[i + 1], *rowtile.neighbours(cuts, i))
        ctxs.append(a)
        mkb = lambda n: torch.empty(int(n), dtype=torch.uint8, device="cuda")  # noqa: E731
        send.append([mkb(a.halo_bytes(0)), mkb(a.halo_bytes(1))])
        recv.append([mkb(a.halo_recv_bytes(0)), mkb(a.halo_recv_bytes(1))])
    dts = [torch.from_numpy(depth[cuts[i]:cuts[i + 1]]).cuda() for i in range(bands)]
    outs = [torch.zeros((cuts[i + 1] - cuts[i], W), dtype=torch.uint8, device="cuda") for i in range(bands)]
    for i, a in enumerate(ctxs):
        a.band_phase_a(dts[i], send[i][0], send[i][1])
    torch.cuda.synchronize()
    for i in range(bands - 1):
        recv[i + 1][0].copy_(send[i][1])
        recv[i][1].copy_(send[i + 1][0])
    torch.cuda.synchronize()
    for i, a in enumerate(ctxs):
        a.band_phase_b(recv[i][0], recv[i][1], outs[i])
    torch.cuda.synchronize()
    got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
    assert int((got != ref).sum()) == 0


@pytest.mark.parametrize("W,H", [(256, 256), (130, 70), (321, 203), (1920, 1080)])
def test_debug_views_bit_exact(torch_cuda, W, H):
    """PushDebugBlitCommands (AO.cs:787-820) for every value of the `debug` property (1..17) plus the
    HighQuality extension ids, incl. the Detile pass (Blit.shader:136-156) over the VIRTUAL atlases."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    ao, orc = _mk(W, H, intensity=1.1, high_quality_mask=0b0101)
    depth = synth.lin01_to_raw(synth.corridor(W, H) if W >= 1000 else synth.random_depth(W, H, seed=4))
    ref = orc.run(depth)
    out = ao.render(torch.from_numpy(depth).cuda())
    for bid in list(range(1, 18)) + [18, 20]:
        got = ao.debug_view(bid)
        ao.synchronize()
        assert np.array_equal(got.cpu().numpy(), orc.debug_view(bid)), (bid, ao.DEBUG_NAMES[bid])
    assert np.array_equal(out.cpu().numpy(), ref)                      # the frame's own output is untouched by the views


def test_debug_view_dump_is_a_valid_pgm(torch_cuda, tmp_path):
    from miniengineao_b200 import synth
    torch = torch_cuda
    W, H = 160, 90
    ao, orc = _mk(W, H)
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    orc.run(depth)
    ao.render(torch.from_numpy(depth).cuda())
    path = str(tmp_path / "tiled1.pgm")
    ao.dump_debug_view(6, path)
    raw = open(path, "rb").read()
    header = b"P5\n%d %d\n255\n" % (W, H)
    assert raw.startswith(header) and len(raw) == len(header) + W * H
    assert np.array_equal(np.frombuffer(raw[len(header):], np.uint8).reshape(H, W), orc.debug_view(6))


def test_composite_branch_selection(torch_cuda):
    """PushCompositeCommands (AO.cs:822-839): forward / non-HDR cameras multiply the frame buffer (pass 2), the
    ambient-only deferred HDR case multiplies GBuffer0.a and the ambient target (pass 1)."""
    from miniengineao_b200 import synth
    from oracle import oracle as O
    torch = torch_cuda
    W, H = 320, 180
    ao, orc = _mk(W, H)
    depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=8))
    ref_ao = orc.run(depth)
    ao_dev = ao.render(torch.from_numpy(depth).cuda())
    rng = np.random.default_rng(5)
    c16 = rng.uniform(0, 3, size=(H, W, 4)).astype(np.float16)
    g0 = rng.integers(0, 256, size=(H, W, 4), dtype=np.uint8)
    col = torch.from_numpy(c16.copy()).cuda()
    assert ao.composite(ao_dev, color=col) == "framebuffer"
    torch.cuda.synchronize()
    assert np.array_equal(col.cpu().numpy().view(np.uint16), O.composite_framebuffer(ref_ao, c16).view(np.uint16))
    ao.camera.actualRenderingPath = "DeferredShading"
    g0d, g3d = torch.from_numpy(g0.copy()).cuda(), torch.from_numpy(c16.copy()).cuda()
    assert ao.composite(ao_dev, gbuffer0=g0d, gbuffer3=g3d) == "gbuffer"
    torch.cuda.synchronize()
    e0, e3 = O.composite_gbuffer(ref_ao, g0, c16)
    assert np.array_equal(g0d.cpu().numpy(), e0) and np.array_equal(g3d.cpu().numpy().view(np.uint16), e3.view(np.uint16))


def _golden_paths():
    import glob
    import os
    return sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


@pytest.mark.parametrize("path", _golden_paths())
def test_cuda_reproduces_golden_fixture(torch_cuda, path):
    """The committed fixtures (tests/golden/, incl. the shader-variant ones): the CUDA path must reproduce the AO bytes and
    every stored intermediate bit for bit -- no oracle involved at test time."""
    import os
    from miniengineao_b200 import AmbientOcclusion, Camera
    torch = torch_cuda
    g = np.load(path)
    H, W = g["depth"].shape
    v = np.zeros(4, np.int32)
    if "variants" in g.files:
        v[:len(g["variants"])] = g["variants"]
    ao = AmbientOcclusion(Camera(W, H), device=0)
    (ao.noiseFilterTolerance, ao.blurTolerance, ao.upsampleTolerance, ao.thicknessModifier, ao.intensity) = (float(x) for x in g["params"])
    ao.sampleExhaustively, ao.highQualityMask, ao.singleScale = bool(v[1]), int(v[2]), bool(v[3])
    got = ao.render(torch.from_numpy(g["depth"]).cuda()).cpu().numpy()
    assert np.array_equal(got, g["ao"]), os.path.basename(path)
    for key in g.files:
        if not key.startswith("buf"):
            continue
        bid, ref = int(key[3:]), g[key]
        buf = ao.debug_buffer(bid)
        assert buf.dtype == ref.dtype and buf.shape == ref.shape, (bid, buf.dtype, buf.shape, ref.dtype, ref.shape)
        view = {1: np.uint8, 2: np.uint16, 4: np.uint32}[ref.dtype.itemsize]
        assert np.array_equal(buf.view(view), ref.view(view)), (os.path.basename(path), bid)


def test_debug_composite_replaces_the_camera_target(torch_cuda):
    """PushCompositeCommands with `debug` > 0 (AO.cs:826-829): the selected debug view lands on the camera target through
    Blit.shader pass 3 (rgba = view.rrrr, no blending); debug == 17 shows the AO texture itself."""
    from miniengineao_b200 import synth
    from oracle import oracle as O
    torch = torch_cuda
    W, H = 322, 190
    ao, orc = _mk(W, H, intensity=1.1)
    depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=6))
    ref_ao = orc.run(depth)
    ao_dev = ao.render(torch.from_numpy(depth).cuda())
    for dbg in (6, 10, 17):
        ao.debug = dbg
        view = ref_ao if dbg == 17 else orc.debug_view(dbg)
        for like in (np.zeros((H, W, 4), np.uint8), np.zeros((H, W, 4), np.float16)):
            target = torch.from_numpy(like.copy()).cuda()
            assert ao.composite(ao_dev, color=target) == "debug"
            torch.cuda.synchronize()
            assert np.array_equal(target.cpu().numpy().view(np.uint8), O.composite_debug(view, like).view(np.uint8)), (dbg, like.dtype)
    ao.debug = 0
    assert np.array_equal(ao.render(torch.from_numpy(depth).cuda()).cpu().numpy(), ref_ao)       # the debug property re-plans, nothing else


# ---- round 2 --------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("W,H", [(256, 256), (130, 70), (1920, 1080)])
def test_single_scale_plan_bit_exact(torch_cuda, W, H):
    """BASELINE.json configs[0]: Downsample -> Render level 1 -> final-style Upsample on Occlusion1 (3 kernels), vs the oracle's twin."""
    from miniengineao_b200 import AmbientOcclusion, Camera, synth
    from oracle.oracle import Oracle
    torch = torch_cuda
    lin = synth.flat_sphere(W, H) if (W, H) == (256, 256) else (synth.corridor(W, H) if W > 1000 else synth.random_depth(W, H, seed=2))
    depth = synth.lin01_to_raw(lin)
    orc = Oracle(W, H, threads=8, intensity=1.1, single_scale=True)
    ref = orc.run(depth)
    ao = AmbientOcclusion(Camera(W, H), device=0)
    ao.intensity, ao.singleScale = 1.1, True
    n0 = ao.launch_count
    got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
    assert np.array_equal(got, ref)
    assert ao.kernels_per_frame == 3 and ao.launch_count - n0 == 3
    assert np.array_equal(ao.debug_buffer(10), orc.codes(10))
    assert np.array_equal(ao.render_host(depth), ref)
    ao.singleScale = False                                  # back to the reference plan: re-plans, 9 kernels, the multi-scale answer
    full = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
    assert np.array_equal(full, Oracle(W, H, threads=8, intensity=1.1).run(depth)) and ao.kernels_per_frame == 9


@pytest.mark.parametrize("tile", ["0", "1", "2"])
def test_render_tile_variants_forced(torch_cuda, tile, monkeypatch):
    """render_ao_kernel<.., TH> for TH = 32 / 16 / 8 forced on EVERY level (MEAO_REN_TILE), TMA and gather tiles, all buffers."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    monkeypatch.setenv("MEAO_REN_TILE", tile)
    W, H = 1920, 1080
    ao, orc = _mk(W, H, intensity=1.1, high_quality_mask=0b0101)
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    ref = orc.run(depth)
    got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()
    assert np.array_equal(got, ref)
    _compare_all(ao, orc, f"MEAO_REN_TILE={tile}", extra=[18, 20])


def _native_bands(torch, W, H, bands, device=0, **attrs):
    """`bands` band contexts on ONE device, connected to each other through meao_band_export / meao_band_connect (in-process:
    direct pointers) -- the same kernels, flags and graph as across GPUs, minus NVLink."""
    from miniengineao_b200 import AmbientOcclusion, Camera, rowtile
    cuts = rowtile.partition(H, bands)
    ctxs = []
    for i in range(bands):
        a = AmbientOcclusion(Camera(W, H), device=device)
        for k, v in attrs.items():
            setattr(a, k, v)
        a.set_row_band(cuts[i], cuts[i + 1], *rowtile.neighbours(cuts, i))
        ctxs.append(a)
    handles = [a.band_export() for a in ctxs]
    for i, a in enumerate(ctxs):
        if i > 0:
            a.band_connect(0, handles[i - 1])
        if i + 1 < bands:
            a.band_connect(1, handles[i + 1])
    streams = [torch.cuda.Stream() for _ in range(bands)]      # one stream per band, like one GPU per band: the exchange kernels must overlap
    return cuts, ctxs, streams


@pytest.mark.parametrize("W,H,bands", [(1920, 1080, 2), pytest.param(3840, 2160, 3, marks=pytest.mark.run_last), (2560, 1440, 3)])
def test_native_exchange_bands_equal_oracle(torch_cuda, W, H, bands):
    """meao_band_step: ONE graph per band (prepare_depth -> band_exchange_kernel -> render x4 + upsample x4); the halo rows move by
    peer stores + epoch flags inside the graph.  Three frames with different depth through the same graphs / flags; the union of
    the bands must equal the ORACLE's whole frame bit for bit every time."""
    from miniengineao_b200 import synth
    from oracle.oracle import Oracle
    torch = torch_cuda
    cuts, ctxs, streams = _native_bands(torch, W, H, bands, intensity=1.1)
    outs = [torch.zeros((cuts[i + 1] - cuts[i], W), dtype=torch.uint8, device="cuda") for i in range(bands)]
    bufs = [torch.empty((cuts[i + 1] - cuts[i], W), dtype=torch.float32, device="cuda") for i in range(bands)]
    orc = Oracle(W, H, threads=8, intensity=1.1)
    for frame in range(3):
        depth = synth.lin01_to_raw(synth.corridor(W, H, frame=frame) if frame < 2 else synth.random_depth(W, H, seed=9))
        ref = orc.run(depth)
        for i in range(bands):
            bufs[i].copy_(torch.from_numpy(depth[cuts[i]:cuts[i + 1]]))
            outs[i].zero_()
        torch.cuda.synchronize()
        for i, a in enumerate(ctxs):                            # issue order is irrelevant: the kernels handshake on the device
            a.band_step(bufs[i], outs[i], stream=streams[i])
        torch.cuda.synchronize()
        assert [a.band_status()["error"] for a in ctxs] == [0] * bands, "a neighbour exchange timed out"
        got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
        assert int((got != ref).sum()) == 0, frame
        for a in ctxs:
            st = a.band_status()
            assert st["error"] == 0 and st["epoch"] == frame + 2, st
    assert ctxs[0].launch_count == 3 * 10 and ctxs[1].launch_count == 3 * 10


@pytest.mark.run_last
def test_native_exchange_8k_bands_equal_oracle(torch_cuda):
    """BASELINE.json configs[3] at full size against the ORACLE (not against our own single-GPU frame): 7680 x 4320, 3 bands.
    (Three, not eight: here all bands share ONE GPU and are driven by ONE host thread, while neighbours handshake through spinning
    kernels -- every band's kernels must be able to start while its neighbour's exchange kernel waits for them.  That needs a
    hardware queue per stream (CUDA_DEVICE_MAX_CONNECTIONS, at most 32) and a later band's first graph instantiation must not wait for
    the device; with 4 bands on one GPU this held whenever earlier tests had run in the process and failed (exchange time-out, stale
    halo rows) when these tests ran first; 2 and 3 bands have passed in every order.  With one band per GPU, as in bench.py --gpus
    2 / 4 / 8, the neighbours never share a device; those runs check their bands against the oracle too:
    configs.8k_single_frame.bands_match_oracle.)"""
    from miniengineao_b200 import synth
    from oracle.oracle import Oracle
    torch = torch_cuda
    W, H, bands = 7680, 4320, 3
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    ref = Oracle(W, H, threads=os.cpu_count() or 8, intensity=1.1).run(depth)
    cuts, ctxs, streams = _native_bands(torch, W, H, bands, intensity=1.1)
    outs = [torch.zeros((cuts[i + 1] - cuts[i], W), dtype=torch.uint8, device="cuda") for i in range(bands)]
    bufs = [torch.from_numpy(depth[cuts[i]:cuts[i + 1]]).cuda() for i in range(bands)]
    for rep in range(2):
        for i, a in enumerate(ctxs):
            a.band_step(bufs[i], outs[i], stream=streams[i])
        torch.cuda.synchronize()
        assert [a.band_status()["error"] for a in ctxs] == [0] * bands, "a neighbour exchange timed out"
        got = np.concatenate([o.cpu().numpy() for o in outs], axis=0)
        bad = np.nonzero((got != ref).any(axis=1))[0]
        assert bad.size == 0, (rep, f"{bad.size} rows differ, first {bad[:5]}, cuts {cuts}")
    del ctxs


def test_native_exchange_missing_neighbour_times_out(torch_cuda, monkeypatch):
    """A band whose neighbour never steps: the exchange kernel gives up after MEAO_BAND_TIMEOUT_MS, the step completes (on stale
    halo rows), the sticky error is visible in meao_band_status and the NEXT meao_band_step is refused with MEAO_ERR_PEER."""
    import time
    from miniengineao_b200 import MeaoError, synth
    from miniengineao_b200 import _native as N
    torch = torch_cuda
    monkeypatch.setenv("MEAO_BAND_TIMEOUT_MS", "100")
    W, H = 1280, 720
    cuts, ctxs, streams = _native_bands(torch, W, H, 2, intensity=1.1)
    depth = torch.from_numpy(synth.lin01_to_raw(synth.corridor(W, H))).cuda()
    out = torch.zeros((cuts[1], W), dtype=torch.uint8, device="cuda")
    t0 = time.time()
    ctxs[0].band_step(depth[:cuts[1]].contiguous(), out, stream=streams[0])
    torch.cuda.synchronize()
    assert 0.09 < time.time() - t0 < 5.0
    st = ctxs[0].band_status()
    assert st["error"] == 1 and st["epoch"] == 2, st
    with pytest.raises(MeaoError) as e:
        ctxs[0].band_step(depth[:cuts[1]].contiguous(), out, stream=streams[0])
    assert e.value.code == N.MEAO_ERR_PEER
    # an unconnected interior band is refused up front
    from miniengineao_b200 import AmbientOcclusion, Camera, rowtile
    lone = AmbientOcclusion(Camera(W, H), device=0)
    lone.set_row_band(cuts[0], cuts[1], *rowtile.neighbours(cuts, 0))
    with pytest.raises(MeaoError) as e2:
        lone.band_step(depth[:cuts[1]].contiguous(), out)
    assert e2.value.code == N.MEAO_ERR_INVALID


def test_resize_after_row_band_resets_the_band(torch_cuda):
    """meao_resize re-allocates and falls back to the whole frame; the host mirror must follow (it used to keep the old
    band height and hand band-sized tensors to a whole-frame context)."""
    from miniengineao_b200 import AmbientOcclusion, Camera, synth
    from oracle.oracle import Oracle
    torch = torch_cuda
    cam = Camera(640, 720)
    ao = AmbientOcclusion(cam, device=0)
    ao.set_row_band(0, 368, -1, 720)
    assert ao._band_rows() == 368
    cam.pixelHeight = 360                                        # camera size change
    depth = synth.lin01_to_raw(synth.random_depth(640, 360, seed=1))
    got = ao.render(torch.from_numpy(depth).cuda()).cpu().numpy()       # whole 640 x 360 frame: shapes are checked against the new size
    assert ao._band_rows() == 360 and got.shape == (360, 640)
    assert np.array_equal(got, Oracle(640, 360, threads=4).run(depth))
    assert ao.band_rows()["produce"][0] == (0, 360)


def test_graph_cache_retargets_instead_of_flushing(torch_cuda):
    """More distinct (depth, out) pairs than the 64-entry graph cache: the least recently used executable graph is re-targeted with
    cudaGraphExecUpdate (no device-wide synchronise + destroy-all), results stay exact and frames in flight are unaffected."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    W, H = 320, 200
    ao, orc = _mk(W, H, intensity=1.1)
    depths, refs = [], []
    for i in range(3):
        d = synth.lin01_to_raw(synth.random_depth(W, H, seed=20 + i))
        depths.append(torch.from_numpy(d).cuda()); refs.append(orc.run(d).copy())
    outs = [torch.zeros((H, W), dtype=torch.uint8, device="cuda") for _ in range(80)]
    for rnd in range(2):
        for i, o in enumerate(outs):                             # 80 output buffers x 3 depth buffers = 240 keys
            ao.render(depths[i % 3], o)
        torch.cuda.synchronize()
        for i, o in enumerate(outs):
            assert np.array_equal(o.cpu().numpy(), refs[i % 3]), (rnd, i)
            o.zero_()


def test_pdl_level_is_reported_and_results_do_not_depend_on_it(torch_cuda, monkeypatch):
    """Programmatic dependent launch inside the captured graph (MEAO_PDL caps the level): same bits with and without."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    W, H = 1920, 1080
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    d = torch.from_numpy(depth).cuda()
    results = []
    for lvl in ("0", "1", "2"):
        monkeypatch.setenv("MEAO_PDL", lvl)
        ao, orc = _mk(W, H, intensity=1.1)
        a = ao.render(d).cpu().numpy()
        b = ao.render(d).cpu().numpy()                           # replay of the cached graph
        assert np.array_equal(a, b)
        results.append(a)
        assert 0 <= ao.pdl_level <= int(lvl)
    assert np.array_equal(results[0], results[1]) and np.array_equal(results[0], results[2])
    assert np.array_equal(results[0], orc.run(depth))


def test_tolerances_outside_the_proven_range_take_the_ieee_path(torch_cuda):
    """upsample_tolerance = -17 (10^-17 < 2^-55) is outside the range for which the planner proves the fast final division
    (UpsampleArgs.fast_div_ok = 0): every thread then runs upsample8_slow, the plain IEEE operators -- same bits as the oracle."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    W, H = 322, 190
    ao, orc = _mk(W, H, intensity=1.1, upsample_tolerance=-17.0, noise_filter_tolerance=-8.0)
    depth = synth.lin01_to_raw(synth.random_depth(W, H, seed=6))
    ref = orc.run(depth)
    assert np.array_equal(ao.render(torch.from_numpy(depth).cuda()).cpu().numpy(), ref)
    _compare_all(ao, orc, "slow path")


def test_pure_c_client_drives_row_bands(torch_cuda, tmp_path):
    """Multi-GPU from plain C ("a C#/C caller of include/meao.h cannot do multi-GPU at all"): tests/c_abi/smoke.c splits
    a 1280 x 1200 frame into 2 / 3 row bands, connects them with meao_band_export / meao_band_connect and steps them with
    meao_band_step_host -- one single-threaded process, no NCCL, no Python.  All bands share device 0 here; with more devices the
    same binary spreads them (last argument)."""
    import os, subprocess
    from test_abi import _build_c_client
    from miniengineao_b200 import synth
    from oracle.oracle import Oracle
    torch = torch_cuda
    W, H = 1280, 1200                                        # three bands of 400 rows: deep enough for the level-4 halo
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    ref = Oracle(W, H, threads=8, intensity=1.1).run(depth)
    dpath, apath = os.path.join(str(tmp_path), "depth.f32"), os.path.join(str(tmp_path), "ao.u8")
    depth.tofile(dpath); ref.tofile(apath)
    exe = _build_c_client(tmp_path)
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32")
    ndev = min(torch.cuda.device_count(), 2)
    for nb, nd in ((2, 1), (3, 1), (2, ndev)):
        r = subprocess.run([exe, "bands", str(W), str(H), dpath, apath, "1.1", str(nb), str(nd)], capture_output=True, text=True, env=env, timeout=120)
        assert r.returncode == 0 and " 0 mismatching pixels" in r.stdout, r.stdout + r.stderr


def test_bad_depth_kind_is_refused_by_every_entry_point(torch_cuda):
    """only meao_render validated depth_kind; the check now sits in the downsample recorder all entry points share."""
    from miniengineao_b200 import AmbientOcclusion, Camera, rowtile
    from miniengineao_b200 import _native as N
    torch = torch_cuda
    W, H = 640, 736
    lib = N.lib()
    ao = AmbientOcclusion(Camera(W, H), device=0)
    ao.LateUpdate()
    d = torch.zeros((H, W), dtype=torch.float32, device="cuda")
    o = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert lib.meao_render(ao._ctx, d.data_ptr(), 7, o.data_ptr(), st) == N.MEAO_ERR_INVALID
    assert lib.meao_stage_downsample(ao._ctx, d.data_ptr(), -1, st) == N.MEAO_ERR_INVALID
    assert lib.meao_render_band_prepare(ao._ctx, d.data_ptr(), 4, st) == N.MEAO_ERR_INVALID
    cuts = rowtile.partition(H, 2)
    ao.set_row_band(cuts[0], cuts[1], *rowtile.neighbours(cuts, 0))
    send = torch.zeros(int(ao.halo_bytes(1)), dtype=torch.uint8, device="cuda")
    assert lib.meao_band_phase_a(ao._ctx, d.data_ptr(), 9, None, send.data_ptr(), st) == N.MEAO_ERR_INVALID
    hp = lib.meao_host_alloc(W * H * 4)
    try:
        assert lib.meao_render_host_async(ao._ctx, hp, 5, hp, 0) == N.MEAO_ERR_INVALID
    finally:
        lib.meao_host_free(hp)
    torch.cuda.synchronize()
    assert b"bad depth kind" in lib.meao_last_error(ao._ctx)


def test_persistent_tile_loop_forced_on_every_level(torch_cuda, monkeypatch):
    """blur_upsample's tile loop (one wave of CTAs pulling tiles from an atomic cursor, the next tile's TMA boxes prefetched into a
    second buffer pair) normally serves only launches with >= 2 tiles per CTA slot (the final level from 4K up -- covered by the 4K /
    8K tests); MEAO_UPS_PERSIST_MIN_WAVES forces it onto every level, where a CTA's tile sequence mixes interior and border tiles
    and most CTAs find the cursor exhausted at once.  Three frames: the counters must be re-armed by the last CTA each time."""
    from miniengineao_b200 import synth
    torch = torch_cuda
    monkeypatch.setenv("MEAO_UPS_PERSIST_MIN_WAVES", "0.0001")
    W, H = 1280, 720
    ao, orc = _mk(W, H, intensity=1.1, high_quality_mask=0b0101)
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    depth[300:340, 500:700] = 0.0
    ref = orc.run(depth)
    d = torch.from_numpy(depth).cuda()
    for frame in range(3):
        assert np.array_equal(ao.render(d).cpu().numpy(), ref), frame
    _compare_all(ao, orc, "forced tile loop", extra=[18, 20])


def test_native_exchange_reconnect_restarts_the_epochs(torch_cuda):
    """Bands that have stepped can be taken apart and connected again (e.g. after one of them timed out): partial reconnects are
    refused, a full disconnect + connect restarts every epoch at 1 and the frames are right again."""
    from miniengineao_b200 import MeaoError, synth
    from miniengineao_b200 import _native as N
    from oracle.oracle import Oracle
    torch = torch_cuda
    W, H, bands = 1280, 1200, 3
    cuts, ctxs, streams = _native_bands(torch, W, H, bands, intensity=1.1)
    depth = synth.lin01_to_raw(synth.corridor(W, H))
    ref = Oracle(W, H, threads=8, intensity=1.1).run(depth)
    outs = [torch.zeros((cuts[i + 1] - cuts[i], W), dtype=torch.uint8, device="cuda") for i in range(bands)]
    bufs = [torch.from_numpy(depth[cuts[i]:cuts[i + 1]]).cuda() for i in range(bands)]

    def frame():
        for o in outs:
            o.zero_()
        for i, a in enumerate(ctxs):
            a.band_step(bufs[i], outs[i], stream=streams[i])
        torch.cuda.synchronize()
        return np.concatenate([o.cpu().numpy() for o in outs], axis=0)
    for _ in range(2):
        assert np.array_equal(frame(), ref)
    assert ctxs[1].band_status()["epoch"] == 3
    handles = [a.band_export() for a in ctxs]
    with pytest.raises(MeaoError) as e:                          # the middle band has stepped and its lower side is still attached
        ctxs[1].band_connect(0, handles[0])
    assert e.value.code == N.MEAO_ERR_INVALID
    for i, a in enumerate(ctxs):                                 # take everything apart ...
        for side in (0, 1):
            a.band_connect(side, None)
    for i, a in enumerate(ctxs):                                 # ... and connect again: epochs restart at 1 everywhere
        if i > 0:
            a.band_connect(0, handles[i - 1])
        if i + 1 < bands:
            a.band_connect(1, handles[i + 1])
    assert [a.band_status()["epoch"] for a in ctxs] == [1, 1, 1]
    for _ in range(2):
        assert np.array_equal(frame(), ref)
    assert [a.band_status()["error"] for a in ctxs] == [0, 0, 0]
