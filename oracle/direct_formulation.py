"""Second, independent restatement of the MiniEngineAO compute path: GLOBAL per-pixel formulas in
vectorised numpy float32 -- no thread groups, no LDS tiles, no deinterleaved atlases.

TEST INFRASTRUCTURE ONLY (see oracle/meao_oracle.h).  PARITY UNPINNED.

Why it exists: the CUDA kernels restructure the reference (natural-layout depth with a *virtual*
atlas, blur defined on virtual low-res coordinates, one thread per output pixel).  This module
states exactly that restructured formulation; tests/test_oracle.py proves it bit-identical to the
literal thread-group oracle (meao_oracle.c, built with -DMEAO_ORACLE_NO_FMA because numpy has no
fused multiply-add), which shows the restructuring itself is exact before any GPU is involved.

Reference lines: Downsample1.compute:37-81, Downsample2.compute:32-51, Render.compute:60-177,
Upsample.compute:54-233, AmbientOcclusion.cs:511-531.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def _sat(x):
    with np.errstate(invalid="ignore"):
        return np.where(x > 0, np.where(x < 1, x, F(1)), F(0)).astype(F)      # NaN -> 0


def _f16r(x):
    with np.errstate(over="ignore"):
        return x.astype(np.float16).astype(F)                                   # RTNE, overflow -> inf


def _unorm8(x):
    c = _sat(x)
    return (c * F(255) + F(0.5)).astype(np.uint8)                               # truncation


def _u8load(k):
    return k.astype(F) * (F(1.0) / F(255.0))


def level_dims(W, H):
    return [((W + (1 << l) - 1) >> l, (H + (1 << l) - 1) >> l) for l in range(7)]


def prepare_depth(depth, zb, reversed_z=True, linear=False):
    """-> LinearDepth (f16-rounded f32), [None, Low1..Low4] (point samples lin(2^k i, 2^k j))."""
    d = depth.astype(F)
    if linear:
        lin = d
    else:
        with np.errstate(divide="ignore"):
            lin = (F(1) / (F(zb[0]) * d + F(zb[1]))).astype(F)                 # DS1:40 (unfused)
        lin = np.where(d == (0 if reversed_z else 1), F(1e5), lin).astype(F)   # DS1:41-45
    lows = [None] + [np.ascontiguousarray(lin[::1 << k, ::1 << k]) for k in range(1, 5)]
    return _f16r(lin), lows


def tiled_view(low_k, sw, sh, pad):
    """TiledDepth<k> [16, sh, sw] as Downsample1/2 write it (incl. padding texels)."""
    lh, lw = low_k.shape
    out = np.full((16, sh, sw), pad, F)
    for s in range(16):
        sx, sy = s & 3, s >> 2
        sub = low_k[sy::4, sx::4]
        out[s, :sub.shape[0], :sub.shape[1]] = sub
    return _f16r(out)


# (table slot, (x, y)) in call order
SAMPLES_CHECKER = ((1, (2, 0)), (3, (4, 0)), (4, (1, 1)), (8, (2, 2)), (11, (3, 3)), (6, (1, 3)), (10, (2, 4)))     # REN:162-168
SAMPLES_EXHAUSTIVE = ((0, (1, 0)), (1, (2, 0)), (2, (3, 0)), (3, (4, 0)), (4, (1, 1)), (8, (2, 2)), (11, (3, 3)),
                      (5, (1, 2)), (6, (1, 3)), (7, (1, 4)), (9, (2, 3)), (10, (2, 4)))                             # REN:148-159


def render_ao(low_k, sw, sh, pad, consts, exhaustive=False):
    """Occlusion<k> codes (kernel main_interleaved).  Pixel (X,Y): slice (X&3,Y&3), slice texel (X>>2,Y>>2); tap (di,dj)
    reads natural pixel (4*clamp(i+di)+sx, 4*clamp(j+dj)+sy), `pad` when outside the level."""
    lh, lw = low_k.shape
    lowh = _f16r(low_k)
    padh = _f16r(np.asarray(pad, F))
    Y, X = np.meshgrid(np.arange(lh), np.arange(lw), indexing="ij")
    I, J, SX, SY = X >> 2, Y >> 2, X & 3, Y & 3

    def tap(di, dj):
        cx = 4 * np.clip(I + di, 0, sw - 1) + SX
        cy = 4 * np.clip(J + dj, 0, sh - 1) + SY
        ok = (cx < lw) & (cy < lh)
        return np.where(ok, lowh[np.minimum(cy, lh - 1), np.minimum(cx, lw - 1)], padh).astype(F)

    return _render_core(tap, (lh, lw), consts, exhaustive)


def render_ao_wide(low_k, consts, exhaustive=False):
    """HighQuality<k> codes (kernel main, WIDE_SAMPLING): f32 source sampled in place, tap (di,dj) reads level pixel
    (clamp(X + 2*di), clamp(Y + 2*dj)) -- REN:79-82 doubles the offsets, REN:125 clamps per texel."""
    lh, lw = low_k.shape
    Y, X = np.meshgrid(np.arange(lh), np.arange(lw), indexing="ij")

    def tap(di, dj):
        return low_k[np.clip(Y + 2 * dj, 0, lh - 1), np.clip(X + 2 * di, 0, lw - 1)].astype(F)

    return _render_core(tap, (lh, lw), consts, exhaustive)


def _render_core(tap, shape, consts, exhaustive):
    lh, lw = shape
    rf = F(consts["reject_fadeoff"])
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        inv_depth = (F(1) / tap(0, 0)).astype(F)                                # REN:140

        def pair(ir, front, ox, oy):                                            # REN:60-75
            d1 = (tap(ox, oy) * ir - front).astype(F)
            d2 = (tap(-ox, -oy) * ir - front).astype(F)
            p1, p2 = _sat(rf * d1), _sat(rf * d2)
            c1 = np.fmin(np.fmax(d1, p2), F(1)).astype(F)
            c2 = np.fmin(np.fmax(d2, p1), F(1)).astype(F)
            return _sat((c1 + c2) - p1 * p2)

        def samples(x, y, it):                                                  # REN:77-110
            it = F(it)
            ir = (it * inv_depth).astype(F)
            front = F(it - F(0.5))
            if y == 0:
                return F(0.5) * (pair(ir, front, x, 0) + pair(ir, front, 0, x))
            if x == y:
                return F(0.5) * (pair(ir, front, -x, x) + pair(ir, front, x, x))
            return F(0.25) * (((pair(ir, front, x, y) + pair(ir, front, -x, y)) + pair(ir, front, y, x)) + pair(ir, front, -y, x))

        iT, sW = consts["inv_thickness"], consts["sample_weight"]
        ao = np.zeros((lh, lw), F)
        for idx, (x, y) in (SAMPLES_EXHAUSTIVE if exhaustive else SAMPLES_CHECKER):
            ao = (F(sW[idx]) * samples(x, y, iT[idx]) + ao).astype(F)
        out = (F(consts["intensity"]) * (ao - F(1)) + F(1)).astype(F)           # REN:176
    return _unorm8(out)


def _smart_blur(a, b, c, d, e, L, M, R):                                        # UPS:74-81
    b = np.where(L | M, b, c)
    a = np.where(L, a, b)
    d = np.where(R | M, d, c)
    e = np.where(R, e, d)
    return ((((a + e) * F(0.5) + b) + c + d) * F(0.25)).astype(F)


def _blur_1d(ao5, id5, step, kblur):
    """ao5/id5: the five taps (arrays) of AO and inverse depth along one axis."""
    dd = [(id5[i + 1] - id5[i]).astype(F) for i in range(4)]
    ll = [(dd[i] * dd[i] + step).astype(F) for i in range(4)]
    with np.errstate(invalid="ignore", over="ignore"):
        cc = []
        for i in range(3):                                                      # UPS:83-87
            t = (dd[i] * dd[i + 1] + step).astype(F)
            cc.append((t * t).astype(F) > ((ll[i] * ll[i + 1]).astype(F) * kblur).astype(F))
    return _smart_blur(ao5[0], ao5[1], ao5[2], ao5[3], ao5[4], cc[0], cc[1], cc[2])


def blur_upsample(lo_depth, lo_ao_codes, hi_depth, hi_ao_codes, consts, lo_ao2_codes=None):
    """AoResult codes at the hi level.  hi_ao_codes None => kernel "main" (UPS:223); lo_ao2_codes given => the
    main_premin variants: AO1 = min(AO1, LoResAO2) per texel before the blur (UPS:58-60)."""
    if lo_ao2_codes is not None:
        lo_ao_codes = np.minimum(lo_ao_codes, lo_ao2_codes)      # k -> k/255 is monotone, so min commutes with the load
    loh, low = lo_depth.shape
    hih, hiw = hi_depth.shape
    step, kblur = F(consts["step_size"]), F(consts["blur_tolerance"])
    tol, nfs = F(consts["upsample_tolerance"]), F(consts["noise_filter_strength"])
    with np.errstate(divide="ignore"):
        inv_d = (F(1) / lo_depth.astype(F)).astype(F)                           # UPS:67
    ao = _u8load(lo_ao_codes)

    # blurred AO on virtual coordinates vx in [-1, low], vy in [-1, loh]; reads clamp to the edge
    vx = np.arange(-1, low + 1)
    vy = np.arange(-1, loh + 1)

    def cx(off):
        return np.clip(vx + off, 0, low - 1)

    # horizontal pass on all rows that the vertical pass can touch: virtual rows [-3, loh+2] clamped
    rows = np.clip(np.arange(-3, loh + 3), 0, loh - 1)
    aoh = _blur_1d([ao[np.ix_(rows, cx(o))] for o in (-2, -1, 0, 1, 2)],
                   [inv_d[np.ix_(rows, cx(o))] for o in (-2, -1, 0, 1, 2)], step, kblur)
    idc = inv_d[np.ix_(rows, cx(0))]                                            # UPS:141-146 (same column)
    n = loh + 2
    blurred = _blur_1d([aoh[o:o + n] for o in range(5)], [idc[o:o + n] for o in range(5)], step, kblur)
    # blurred[r, c] <-> virtual (vx = c - 1, vy = r - 1)

    PY, PX = np.meshgrid(np.arange(hih), np.arange(hiw), indexing="ij")
    Xq, Yq = (PX + 1) >> 1, (PY + 1) >> 1                                       # quad X-1..X, Y-1..Y

    def lo_d(xx, yy):
        return lo_depth[np.clip(yy, 0, loh - 1), np.clip(xx, 0, low - 1)].astype(F)   # UPS:225 clamp

    def lo_a(xx, yy):
        return blurred[yy + 1, xx + 1]

    quad = {"bl": (Xq - 1, Yq), "br": (Xq, Yq), "tr": (Xq, Yq - 1), "tl": (Xq - 1, Yq - 1)}
    D = {k: lo_d(*v) for k, v in quad.items()}
    A = {k: lo_a(*v) for k, v in quad.items()}
    hd = hi_depth.astype(F)
    ha = _u8load(hi_ao_codes) if hi_ao_codes is not None else np.ones((hih, hiw), F)
    xo, yo = (PX & 1) == 1, (PY & 1) == 1
    orders = {(True, False): ("bl", "br", "tr", "tl"),      # UPS:229 (-1, 0)
              (False, False): ("br", "tr", "tl", "bl"),     # UPS:230 ( 0, 0)
              (False, True): ("tr", "tl", "bl", "br"),      # UPS:231 ( 0,-1)
              (True, True): ("tl", "bl", "br", "tr")}       # UPS:232 (-1,-1)
    res = np.zeros((hih, hiw), F)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        for (xodd, yodd), order in orders.items():
            m = (xo == xodd) & (yo == yodd)
            w = [(F(c) / (np.abs(hd - D[k]) + tol)).astype(F) for c, k in zip((9, 3, 1, 3), order)]   # UPS:179
            total = ((((w[0] + w[1]) + w[2]) + w[3]) + nfs).astype(F)                                   # UPS:180
            ws = (A[order[0]] * w[0]).astype(F)
            for i in (1, 2, 3):
                ws = (A[order[i]] * w[i] + ws).astype(F)
            ws = (ws + nfs).astype(F)                                                                   # UPS:181
            r = ((ha * ws).astype(F) / total).astype(F)                                                 # UPS:182
            res = np.where(m, r, res)
    return _unorm8(res)


def run(depth, render_consts, upsample_consts, zb, *, reversed_z=True, linear=False, return_all=False,
        exhaustive=False, high_quality_mask=0, render_consts_wide=None, single_scale=False):
    """Whole pipe (AO.cs:511-531).  render_consts / upsample_consts: dicts per level from the oracle.
    single_scale (BASELINE.json configs[0]): Downsample + Render level 1 + the final-style Upsample fed Occlusion1."""
    H, W = depth.shape
    dims = level_dims(W, H)
    lin_h, low = prepare_depth(depth, zb, reversed_z, linear)
    if linear:
        pad12 = F(0)
    else:
        pad12 = F(1e5) if reversed_z else F(F(1) / F(zb[1]))
    occ = [None] * 5
    if single_scale:
        sw, sh = dims[3]
        occ[1] = render_ao(low[1], sw, sh, pad12, render_consts[1], exhaustive)
        res = blur_upsample(low[1], occ[1], lin_h, None, upsample_consts[1], None)
        if return_all:
            return {"linear": lin_h, "low": low, "occ": occ, "comb": [res, None, None, None], "pad12": pad12, "hq": [None] * 5}
        return res
    for k in range(1, 5):
        sw, sh = dims[k + 2]
        occ[k] = render_ao(low[k], sw, sh, pad12 if k <= 2 else F(0), render_consts[k], exhaustive)
    hq = [None] * 5
    for k in range(1, 5):
        if (high_quality_mask >> (k - 1)) & 1:
            hq[k] = render_ao_wide(low[k], render_consts_wide[k], exhaustive)
    comb = [None] * 4
    lo_ao = occ[4]
    for lo in range(4, 0, -1):
        hi = lo - 1
        hi_depth = lin_h if hi == 0 else low[hi]
        hi_ao = None if hi == 0 else occ[hi]
        comb[hi] = blur_upsample(low[lo], lo_ao, hi_depth, hi_ao, upsample_consts[lo], hq[lo])
        lo_ao = comb[hi]
    if return_all:
        return {"linear": lin_h, "low": low, "occ": occ, "comb": comb, "pad12": pad12, "hq": hq}
    return comb[0]


def debug_view(buf, W, H):
    """PushDebugBlitCommands (AO.cs:787-820): the W x H R8 image of a buffer.  buf: [h, w] (stretch blit, point sampling at
    the pixel centre) or [16, h, w] (Blit.shader pass 4 :150-152: a 4 x 4 mosaic of the slices).  Values, not codes."""
    y, x = np.arange(H, dtype=np.int64), np.arange(W, dtype=np.int64)
    if buf.ndim == 3:
        _, sh, sw = buf.shape
        nx, ny = 8 * x + 4, 8 * y + 4
        qx, qy = nx // (2 * W), ny // (2 * H)
        tx, ty = (nx - 2 * W * qx) * sw // (2 * W), (ny - 2 * H * qy) * sh // (2 * H)
        v = buf[(qx[None, :] + 4 * qy[:, None]), ty[:, None], tx[None, :]]
    else:
        sh, sw = buf.shape
        v = buf[((2 * y + 1) * sh // (2 * H))[:, None], ((2 * x + 1) * sw // (2 * W))[None, :]]
    return _unorm8(v.astype(F))
