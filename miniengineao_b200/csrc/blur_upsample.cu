// blur_upsample.cu -- stage 3 of the SSAO pipe: depth-aware 5x5 separable blur of the low-res AO
// followed by a 4-tap bilateral upsample (optionally multiplied by the hi-res AO).
//
// Replaces Upsample.compute kernels main / main_blendout (PrefetchData :54-72, SmartBlur :74-81,
// CompareDeltas :83-87, BlurHorizontally :89-130, BlurVertically :132-170, BilateralUpsample
// :177-183, MAIN :185-233).
//
// Design (not a port): the reference maps one thread to a 2x2 output quad with an 8x8 group and a
// 16x16 LDS tile (3.5x apron overhead, 39/64 and 45/64 lanes active in the blur).  Here a CTA owns
// a 64x32 tile of HI-res outputs; the 38x22 low-res footprint (depth f32 + AO unorm8) arrives by two
// TMA box loads, the blur runs on 4-wide / 3-tall register runs so neighbouring outputs share
// their depth deltas, and the upsample streams hi-res depth / AO / result with 128-/64-bit
// accesses, 8 pixels per thread.
//
// The blurred value B(vx,vy) is a pure function of the low-res texels clamp(vx+dx), clamp(vy+dy)
// for |dx|,|dy| <= 2 (point + clamp Gather, UPS:56,67) and is defined for the virtual coordinates
// vx in [-1, low.w], so it does not depend on the reference's group tiling.  Hi-res pixel (px,py)
// uses the quad X-1..X, Y-1..Y with X = (px+1)>>1, Y = (py+1)>>1 and the weight order of
// UPS:229-232.
//
// Bound: mixed -- 5 IEEE divisions per output pixel make the final level issue-heavy next to
// its 2+1+1 B/px of HBM traffic.
//
// Variant PREMIN = kernels main_premin / main_premin_blendout (COMBINE_LOWER_RESOLUTIONS,
// UPS:23,25,32-34,58-60): a second low-res AO texture (LoResAO2 = HighQuality<lo>, the output of Render.compute
// kernel `main`) is min-combined with LoResAO1 texel by texel before the blur.  One more u8 TMA box; the min is
// taken on the unorm8 codes (k -> k/255 is monotone, so it commutes with the load conversion).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace meao {

namespace {

constexpr int kHW = 64, kHH = 32;               // hi-res outputs per CTA
constexpr int kRawW = 38, kRawH = 22;           // low-res footprint actually used
constexpr int kRawP = 40;                       // pitch of the raw arrays
constexpr int kLoDP = 42;                       // pitch of the lo_depth array: 4 rows apart = 168 words = 8 banks, so the four
                                                // row groups of a warp read distinct banks in the upsample phase
constexpr int kBoxDP = kUpsDepthBoxW;           // 40: depth TMA box width  (box column = raw column + kBoxDOff)
constexpr int kBoxAP = kUpsAoBoxW;              // 64: AO TMA box width     (box column = raw column + kBoxAOff)
// MEASURED on B200: cp.async.bulk.tensor (tiled, no swizzle) raises "illegal instruction" unless the
// innermost start coordinate * element size is a multiple of 16 bytes.  The raw tile starts at low-res
// column 32*bx - 3, so the boxes start at 32*bx - 4 (f32: 16 B aligned) and 32*bx - 16 (u8).
constexpr int kBoxDOff = 1, kBoxAOff = 13;
constexpr int kBlurW = 34, kBlurH = 18;         // blurred texels needed
constexpr int kBlurP = 36;                      // pitch of the blurred arrays
constexpr int kThreads = 256;
static_assert(kRawH == kUpsDepthBoxH && kRawH == kUpsAoBoxH, "TMA box mismatch");
constexpr int kBoxDElems = (kRawH * kBoxDP * 4 + 127) / 128 * 128 / 4;      // 3520 B -> 3584 B
static_assert((kRawH * kBoxAP) % 128 == 0, "the AO box copies must stay 128-byte aligned");

struct __align__(128) Smem {
    alignas(128) float box_depth[2][kBoxDElems];        // TMA destination: low-res depth box (LoResDB); two copies: tile i+1 is prefetched while tile i is
                                                        // processed.  Each copy padded to a multiple of 128 bytes: a TMA destination must be 128-byte aligned
    alignas(128) uint8_t box_ao[2][kRawH * kBoxAP];     // TMA destination: low-res AO codes box (LoResAO1)
    alignas(16) float lo_depth[2][kRawH * kLoDP];       // raw low-res depth, column 0 = virtual column lx0 (read until the end of phase 4: double-buffered)
    alignas(16) float inv_depth[kRawH * kRawP];     // DepthCache, UPS:67-71
    alignas(16) float ao[kRawH * kRawP];            // AOCache1 as loaded, UPS:62-65
    alignas(16) float hblur[kRawH * kBlurP];        // AOCache2, UPS:127-129
    alignas(16) float vblur[kBlurH * kBlurP];       // AOCache1 after the vertical pass, UPS:168-169
    struct alignas(16) Bar { uint64_t v; uint64_t pad_; } bar_[2];     // one mbarrier per box-buffer pair, each in its own 16-byte slot
    alignas(16) int4 tile[2];                       // this / the next iteration's tile, decoded by thread 0: hx0 (-1 = none), hy0, interior
};
struct SmemPremin : Smem {
    alignas(128) uint8_t box_ao2[2][kRawH * kBoxAP];    // TMA destination: second low-res AO codes box (LoResAO2)
};

// Upsample.compute:177-183 with the swizzled argument order of :229-232.
// FAST: the five divisions use div_fast (common.cuh) and `ok` collects the validity guard of the whole
// group; when it ends up false the caller recomputes with FAST = false (plain IEEE operators).
template <bool FAST, bool BLEND>
__device__ __forceinline__ float bilateral(float hi_depth, float hi_ao,
                                           float ld0, float ld1, float ld2, float ld3,
                                           float la0, float la1, float la2, float la3,
                                           float tol, float nfs, bool &ok)
{
    const float b0 = __fadd_rn(fabsf(__fadd_rn(hi_depth, -ld0)), tol);
    const float b1 = __fadd_rn(fabsf(__fadd_rn(hi_depth, -ld1)), tol);
    const float b2 = __fadd_rn(fabsf(__fadd_rn(hi_depth, -ld2)), tol);
    const float b3 = __fadd_rn(fabsf(__fadd_rn(hi_depth, -ld3)), tol);
    float w0, w1, w2, w3;
    if (FAST) {
        // every b_i >= tol >= 2^-60 (host-checked); their sum < 2^60 bounds them above and catches inf / NaN
        ok = ok & (__fadd_rn(__fadd_rn(b0, b1), __fadd_rn(b2, b3)) < 1152921504606846976.0f);
        w0 = div_fast(9.0f, b0); w1 = div_fast(3.0f, b1); w2 = div_fast(1.0f, b2); w3 = div_fast(3.0f, b3);
    } else {
        w0 = 9.0f / b0; w1 = 3.0f / b1; w2 = 1.0f / b2; w3 = 3.0f / b3;
    }
    const float total = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(w0, w1), w2), w3), nfs);
    const float wsum = __fadd_rn(fmaf(la3, w3, fmaf(la2, w2, fmaf(la1, w1, __fmul_rn(la0, w0)))), nfs);
    const float num = BLEND ? __fmul_rn(hi_ao, wsum) : wsum;      // HiSSAOs = 1 without blend (UPS:223): 1 * x == x
    if (FAST) {
        ok = ok & in_safe_range(total) & ((num == 0.0f) | in_safe_range(num));
        return div_fast(num, total);
    }
    return num / total;
}

// ---- packed (two-lane) 5-tap depth-aware blur -------------------------------------------------
// The two lanes are two independent rows (horizontal pass) or two independent columns (vertical pass);
// each lane performs exactly CompareDeltas (Upsample.compute:83-87) and SmartBlur (:74-81; /2 and /4 are exact scalings).
__device__ __forceinline__ void compare_deltas2(float2 d1, float2 d2, float2 l1, float2 l2, float2 step2, float2 kblur2, bool &cx, bool &cy)
{
    const float2 temp = __ffma2_rn(d1, d2, step2);
    const float2 tt = __fmul2_rn(temp, temp);
    const float2 lk = __fmul2_rn(__fmul2_rn(l1, l2), kblur2);
    cx = tt.x > lk.x; cy = tt.y > lk.y;
}
__device__ __forceinline__ float2 smart_blur2(float2 a, float2 b, float2 c, float2 d, float2 e,
                                              bool Lx, bool Mx, bool Rx, bool Ly, bool My, bool Ry)
{
    b.x = (Lx | Mx) ? b.x : c.x;  b.y = (Ly | My) ? b.y : c.y;
    a.x = Lx ? a.x : b.x;         a.y = Ly ? a.y : b.y;
    d.x = (Rx | Mx) ? d.x : c.x;  d.y = (Ry | My) ? d.y : c.y;
    e.x = Rx ? e.x : d.x;         e.y = Ry ? e.y : d.y;
    const float2 s = __fadd2_rn(__fadd2_rn(__fadd2_rn(__fmul2_rn(__fadd2_rn(a, e), make_float2(0.5f, 0.5f)), b), c), d);
    return __fmul2_rn(s, make_float2(0.25f, 0.25f));
}
// N outputs from N + 4 taps per lane
template <int N>
__device__ __forceinline__ void blur_run2(const float2 (&av)[N + 4], const float2 (&dv)[N + 4], float step, float kblur, float2 (&out)[N])
{
    const float2 m1 = make_float2(-1.0f, -1.0f), step2 = make_float2(step, step), k2 = make_float2(kblur, kblur);
    float2 dd[N + 3], ll[N + 3];
    bool cx[N + 2], cy[N + 2];
#pragma unroll
    for (int i = 0; i < N + 3; i++) { dd[i] = __ffma2_rn(dv[i], m1, dv[i + 1]); ll[i] = __ffma2_rn(dd[i], dd[i], step2); }   // d[i+1] - d[i]
#pragma unroll
    for (int i = 0; i < N + 2; i++) compare_deltas2(dd[i], dd[i + 1], ll[i], ll[i + 1], step2, k2, cx[i], cy[i]);
#pragma unroll
    for (int i = 0; i < N; i++)
        out[i] = smart_blur2(av[i], av[i + 1], av[i + 2], av[i + 3], av[i + 4], cx[i], cx[i + 1], cx[i + 2], cy[i], cy[i + 1], cy[i + 2]);
}

// ---- packed-f32x2 fast path -------------------------------------------------------------------
// Blackwell's FFMA2 / FADD2 / FMUL2 (PTX fma.rn.f32x2 ...) do two IEEE fp32 operations per lane and per
// issue slot.  The kernel is issue-bound, so the fast path evaluates TWO pixels of equal x parity
// (e, e+2: same operand order) with packed arithmetic.  Every lane of every packed instruction performs
// exactly the scalar operation of bilateral<true>, so the result is bit-identical; nb_i = -(|hi-lo_i| + tol)
// is formed directly in negated form (a sign flip is exact) because packed ops have no negate modifier.
__device__ __forceinline__ float2 div2_fast_neg(float2 num, float2 nden)      // num / (-nden), lane-wise div_fast
{
    float2 y = make_float2(rcp_approx(-nden.x), rcp_approx(-nden.y));
    const float2 one = make_float2(1.0f, 1.0f);
    const float2 e = __ffma2_rn(nden, y, one);
    y = __ffma2_rn(y, e, y);
    const float2 q = __fmul2_rn(num, y);
    const float2 r = __ffma2_rn(nden, q, num);
    return __ffma2_rn(y, r, q);
}

template <bool BLEND>
__device__ __forceinline__ float2 bilateral2(float2 hd, float2 ha,
                                             float2 ld0, float2 ld1, float2 ld2, float2 ld3,
                                             float2 la0, float2 la1, float2 la2, float2 la3,
                                             float tol, float nfs, bool &ok)
{
    const float2 m1 = make_float2(-1.0f, -1.0f);
    const float2 t0 = __ffma2_rn(ld0, m1, hd), t1 = __ffma2_rn(ld1, m1, hd);          // hd - ld_i (one rounding, == FADD)
    const float2 t2 = __ffma2_rn(ld2, m1, hd), t3 = __ffma2_rn(ld3, m1, hd);
    const float2 nb0 = make_float2(__fadd_rn(-fabsf(t0.x), -tol), __fadd_rn(-fabsf(t0.y), -tol));
    const float2 nb1 = make_float2(__fadd_rn(-fabsf(t1.x), -tol), __fadd_rn(-fabsf(t1.y), -tol));
    const float2 nb2 = make_float2(__fadd_rn(-fabsf(t2.x), -tol), __fadd_rn(-fabsf(t2.y), -tol));
    const float2 nb3 = make_float2(__fadd_rn(-fabsf(t3.x), -tol), __fadd_rn(-fabsf(t3.y), -tol));
    const float2 s = __fadd2_rn(__fadd2_rn(nb0, nb1), __fadd2_rn(nb2, nb3));
    ok = ok & (s.x > -1152921504606846976.0f) & (s.y > -1152921504606846976.0f);     // guard of bilateral<true>
    const float2 w0 = div2_fast_neg(make_float2(9.0f, 9.0f), nb0);
    const float2 w1 = div2_fast_neg(make_float2(3.0f, 3.0f), nb1);
    const float2 w2 = div2_fast_neg(make_float2(1.0f, 1.0f), nb2);
    const float2 w3 = div2_fast_neg(make_float2(3.0f, 3.0f), nb3);
    const float2 nfs2 = make_float2(nfs, nfs);
    const float2 total = __fadd2_rn(__fadd2_rn(__fadd2_rn(__fadd2_rn(w0, w1), w2), w3), nfs2);
    const float2 wsum = __fadd2_rn(__ffma2_rn(la3, w3, __ffma2_rn(la2, w2, __ffma2_rn(la1, w1, __fmul2_rn(la0, w0)))), nfs2);
    const float2 num = BLEND ? __fmul2_rn(ha, wsum) : wsum;
#if !MEAO_UPS_STATIC_GUARD
    ok = ok & in_safe_range(total.x) & in_safe_range(total.y)
            & ((num.x == 0.0f) | in_safe_range(num.x)) & ((num.y == 0.0f) | in_safe_range(num.y));
#endif
    // MEAO_UPS_STATIC_GUARD: the host sets fast_div_ok only when tol >= 2^-55 and 2^-52 <= nfs < 2^59 (true for every value in
    // the component's parameter ranges, AO.cs:20-42).  Once the guard above has passed, every b_i is in [tol, 2^60), so
    // w_i = c_i / b_i <= 9 / tol, total and wsum lie in [nfs, 16 / tol + nfs] (inside [2^-60, 2^60)) and num = ha * wsum is 0 or
    // >= nfs / 255 >= 2^-60: the range test of the final division can never fail and is not evaluated per pixel.
    return div2_fast_neg(num, __fmul2_rn(total, m1));
}

// ---- phase-4 restructure (MEAO_UPS_V2, default on) ------------------------------------------------------------------
// The kernel is issue-bound and the bilateral upsample is 62 % of its issue slots, so phase 4 was rebuilt around the slot count:
//   * a real 2-iteration loop over the thread's two 4-pixel halves (pairs (0,2) (1,3) | (4,6) (5,7)): half the live low-res
//     operands, no register spills under the 48-register cap, half the code;
//   * the range test of the final division is proven on the host from the two tolerances (what MEAO_UPS_STATIC_GUARD did) and
//     the one remaining run-time test -- every b_i finite and below 2^60 -- is accumulated as ONE integer max over the sign-
//     ordered bit patterns of the four pair sums (a NaN, -inf or too large a sum has a larger signed pattern than -2^60);
//   * w2 = 1 / b2 is the packed reciprocal (two roundings fewer instructions than the division; both are the correctly
//     rounded 1 / b2, so the bits are the same);
//   * the last step of the final division is issued as two FFMA.SAT (the saturate of the UNORM8 store rides on it) and the
//     + 0.5 of the store conversion runs packed.
// Threads that fail the test (sky: inf / NaN / zero / denormal operands), partial row ends and parameter sets outside the
// proven range take upsample8_slow(): the plain IEEE operators, out of line.
#ifndef MEAO_UPS_V2
#define MEAO_UPS_V2 1
#endif

// sign-ordered pattern of a NEGATIVE float: more negative (or NaN = 0x7fffffff) => larger signed integer
__device__ __forceinline__ int neg_order(float x) { return (int)__float_as_uint(x); }

template <bool BLEND>
__device__ __forceinline__ uint32_t bilateral2_v2(float2 hd, float2 ha,
                                                  float2 ld0, float2 ld1, float2 ld2, float2 ld3,
                                                  float2 la0, float2 la1, float2 la2, float2 la3,
                                                  float tol, float nfs, int &worst)
{
    const float2 m1 = make_float2(-1.0f, -1.0f);
    const float2 t0 = __ffma2_rn(ld0, m1, hd), t1 = __ffma2_rn(ld1, m1, hd);          // hd - ld_i (one rounding, == FADD)
    const float2 t2 = __ffma2_rn(ld2, m1, hd), t3 = __ffma2_rn(ld3, m1, hd);
    const float2 nb0 = make_float2(__fadd_rn(-fabsf(t0.x), -tol), __fadd_rn(-fabsf(t0.y), -tol));
    const float2 nb1 = make_float2(__fadd_rn(-fabsf(t1.x), -tol), __fadd_rn(-fabsf(t1.y), -tol));
    const float2 nb2 = make_float2(__fadd_rn(-fabsf(t2.x), -tol), __fadd_rn(-fabsf(t2.y), -tol));
    const float2 nb3 = make_float2(__fadd_rn(-fabsf(t3.x), -tol), __fadd_rn(-fabsf(t3.y), -tol));
    const float2 s = __fadd2_rn(__fadd2_rn(nb0, nb1), __fadd2_rn(nb2, nb3));
    worst = max(max(worst, neg_order(s.x)), neg_order(s.y));                          // guard of bilateral<true>: s > -2^60, finite
    const float2 w0 = div2_fast_neg(make_float2(9.0f, 9.0f), nb0);
    const float2 w1 = div2_fast_neg(make_float2(3.0f, 3.0f), nb1);
    const float2 w2 = rcp2_fast_neg(nb2);                                             // RN(1 / b2) == div_fast(1, b2)
    const float2 w3 = div2_fast_neg(make_float2(3.0f, 3.0f), nb3);
    const float2 nfs2 = make_float2(nfs, nfs);
    const float2 total = __fadd2_rn(__fadd2_rn(__fadd2_rn(__fadd2_rn(w0, w1), w2), w3), nfs2);
    const float2 wsum = __fadd2_rn(__ffma2_rn(la3, w3, __ffma2_rn(la2, w2, __ffma2_rn(la1, w1, __fmul2_rn(la0, w0)))), nfs2);
    const float2 num = BLEND ? __fmul2_rn(ha, wsum) : wsum;
    // num / total, lane-wise div_fast; the host has proven total and num inside the fast-division range (UpsampleArgs.fast_div_ok,
    // see the MEAO_UPS_STATIC_GUARD note in bilateral2), the saturate of the store conversion is fused into the last FMA
    const float2 nden = __fmul2_rn(total, m1);
    float2 y = make_float2(rcp_approx(total.x), rcp_approx(total.y));
    const float2 e = __ffma2_rn(nden, y, make_float2(1.0f, 1.0f));
    y = __ffma2_rn(y, e, y);
    const float2 q = __fmul2_rn(num, y);
    const float2 r = __ffma2_rn(nden, q, num);
    const float2 c = make_float2(__saturatef(fmaf(y.x, r.x, q.x)), __saturatef(fmaf(y.y, r.y, q.y)));
    // unorm8_code: c * 255 and + 0.5 are TWO roundings.  ptxas contracts a packed mul.rn.f32x2 feeding an add.rn.f32x2 into one FFMA2
    // (seen in the SASS; the explicit .rn does not protect the packed forms), so the product stays scalar -- mul.rn.f32 is never fused
    const float2 k = __fadd2_rn(make_float2(__fmul_rn(c.x, 255.0f), __fmul_rn(c.y, 255.0f)), make_float2(0.5f, 0.5f));
    return (uint32_t)k.x | ((uint32_t)k.y << 16);                                     // codes of pixels E (bits 0..7) and E + 2 (bits 16..23)
}

// The rare path: all eight pixels of a thread with the plain IEEE operators (UPS:177-183, 229-232), partial rows included.
// (scalar arguments, not the argument block by reference: taking its address would force a local-memory copy of the kernel parameters)
template <bool BLEND, bool HI_HALF>
__device__ __noinline__ void upsample8_slow(const void *hi_depth, int hi_dpitch, const uint8_t *hi_ao, int hi_apitch, uint8_t *out, int out_pitch,
                                            int out_row_origin, int hiw, float tol, float nfs,
                                            const float *vblur, const float *lo_depth, int rY, int j, int py, int px0)
{
    float bl_ao[2][6], lo_d[2][6];
#pragma unroll
    for (int rr = 0; rr < 2; rr++)
#pragma unroll
        for (int i = 0; i < 6; i++) {
            bl_ao[rr][i] = vblur[(rY - 1 + rr) * kBlurP + 4 * j + i];
            lo_d[rr][i] = lo_depth[(rY - 1 + rr + 2) * kLoDP + 4 * j + 2 + i];
        }
    const bool y_odd = (py & 1) != 0;
    uint8_t *dst = out + (size_t)(py - out_row_origin) * out_pitch + px0;
#pragma unroll 1
    for (int e = 0; e < 8; e++) {
        if (px0 + e >= hiw) break;
        float hd, ha = 1.0f;                                                                                     // UPS:223
        if (HI_HALF) hd = __half2float(reinterpret_cast<const __half *>(hi_depth)[(size_t)py * hi_dpitch + px0 + e]);
        else hd = __ldg(reinterpret_cast<const float *>(hi_depth) + (size_t)py * hi_dpitch + px0 + e);
        if (BLEND) ha = unorm8_load(__ldg(hi_ao + (size_t)py * hi_apitch + px0 + e));
        const int m = (e + 1) >> 1;
        const float tl_d = lo_d[0][m], tr_d = lo_d[0][m + 1], bl_d = lo_d[1][m], br_d = lo_d[1][m + 1];
        const float tl_a = bl_ao[0][m], tr_a = bl_ao[0][m + 1], bl_a = bl_ao[1][m], br_a = bl_ao[1][m + 1];
        bool unused = true;
        float r;
        if ((e & 1) != 0) {
            if (!y_odd) r = bilateral<false, BLEND>(hd, ha, bl_d, br_d, tr_d, tl_d, bl_a, br_a, tr_a, tl_a, tol, nfs, unused);   // UPS:229
            else        r = bilateral<false, BLEND>(hd, ha, tl_d, bl_d, br_d, tr_d, tl_a, bl_a, br_a, tr_a, tol, nfs, unused);   // UPS:232
        } else {
            if (!y_odd) r = bilateral<false, BLEND>(hd, ha, br_d, tr_d, tl_d, bl_d, br_a, tr_a, tl_a, bl_a, tol, nfs, unused);   // UPS:230
            else        r = bilateral<false, BLEND>(hd, ha, tr_d, tl_d, bl_d, br_d, tr_a, tl_a, bl_a, br_a, tol, nfs, unused);   // UPS:231
        }
        dst[e] = (uint8_t)unorm8_code(r);
    }
}

// Run lengths of the two blur passes (outputs per thread and lane).  Longer runs share more depth deltas between neighbouring
// outputs (fewer instructions in total), shorter runs put more of the CTA's eight warps to work and shorten the phase.
#ifndef MEAO_UPS_HRUN
#define MEAO_UPS_HRUN 4         // 4: 99 threads; 2: 187 threads
#endif
#ifndef MEAO_UPS_VRUN
#define MEAO_UPS_VRUN 6         // 6: 51 threads; 3: 102 threads; 2: 153 threads
#endif
#ifndef MEAO_UPS_V2_UNROLL
#define MEAO_UPS_V2_UNROLL 1        // 1: both 4-pixel halves of phase 4 in flight (more ILP; 4 bytes spilled at the 48-register cap); measured
                                    // with the tile loop, final level: 33.1 vs 35.1 us (0 = the rolled loop)
#endif
// Persistent tile loop with TMA prefetch of the next tile (see blur_upsample_kernel.inc); 0 = one CTA per tile as in round 1.
#ifndef MEAO_UPS_PERSIST
#define MEAO_UPS_PERSIST 1
#endif
#ifndef MEAO_UPS_MINB
#define MEAO_UPS_MINB 5
#endif
#define MEAO_UPS_PREMIN 0
#include "blur_upsample_kernel.inc"
#undef MEAO_UPS_PREMIN
#define MEAO_UPS_PREMIN 1
#include "blur_upsample_kernel.inc"
#undef MEAO_UPS_PREMIN

}  // namespace

cudaError_t launch_blur_upsample(const CUtensorMap &lo_depth_map, const CUtensorMap &lo_ao_map, const CUtensorMap *lo_ao2_map, bool use_tma,
                                 const UpsampleArgs &a_in, const uint8_t *lo_ao2, int lo_a2pitch, cudaStream_t s)
{
    if (a_in.row1 <= a_in.row0) return cudaSuccess;
    UpsampleArgs a = a_in;
    const int ybase = a.row0 & ~1;
    a.tiles_x = ceil_div(a.hiw, kHW); a.tiles_y = ceil_div(a.row1 - ybase, kHH);
    const int ntiles = a.tiles_x * a.tiles_y;
    // The tile loop pays when a CTA gets several tiles (the final level of a 4K frame: 5.5): the next tile's boxes are prefetched and
    // the launch ramp is paid once.  With ~1-2 tiles per CTA it loses (measured: 1020 tiles, 15.8 vs 14.7 us; 255 tiles, 8.4 vs 6.5 us):
    // the atomic fetch + staging sit on a short critical path and a 2-vs-1 split of tiles is the worst possible tail.
    constexpr int kWave = 148 * MEAO_UPS_MINB;
    const char *force = getenv("MEAO_UPS_PERSIST_MIN_WAVES");       // tuning aid: tiles / wave from which the loop is used (default 2)
    const double min_waves = force ? atof(force) : 2.0;
    const bool persist = MEAO_UPS_PERSIST && a.tile_ctr && ntiles >= (int)(min_waves * kWave);
    if (!persist) a.tile_ctr = nullptr;
    dim3 grid(persist ? kWave : ntiles);
    const int t = use_tma ? 1 : 0;
    if (!lo_ao2) {
        if (a.hi_ao) {
            if (a.hi_is_half) MEAO_LAUNCH((blur_upsample_kernel<true, true>), grid, kThreads, 0, s, lo_depth_map, lo_ao_map, a, t);
            else              MEAO_LAUNCH((blur_upsample_kernel<true, false>), grid, kThreads, 0, s, lo_depth_map, lo_ao_map, a, t);
        } else {
            if (a.hi_is_half) MEAO_LAUNCH((blur_upsample_kernel<false, true>), grid, kThreads, 0, s, lo_depth_map, lo_ao_map, a, t);
            else              MEAO_LAUNCH((blur_upsample_kernel<false, false>), grid, kThreads, 0, s, lo_depth_map, lo_ao_map, a, t);
        }
    } else {        // main_premin / main_premin_blendout
        if (!lo_ao2_map) return cudaErrorInvalidValue;
        const UpsamplePreminArgs pa{a, lo_ao2, lo_a2pitch};
        if (a.hi_ao) {
            if (a.hi_is_half) MEAO_LAUNCH((blur_upsample_premin_kernel<true, true>), grid, kThreads, 0, s, lo_depth_map, lo_ao_map, *lo_ao2_map, pa, t);
            else              MEAO_LAUNCH((blur_upsample_premin_kernel<true, false>), grid, kThreads, 0, s, lo_depth_map, lo_ao_map, *lo_ao2_map, pa, t);
        } else {
            if (a.hi_is_half) MEAO_LAUNCH((blur_upsample_premin_kernel<false, true>), grid, kThreads, 0, s, lo_depth_map, lo_ao_map, *lo_ao2_map, pa, t);
            else              MEAO_LAUNCH((blur_upsample_premin_kernel<false, false>), grid, kThreads, 0, s, lo_depth_map, lo_ao_map, *lo_ao2_map, pa, t);
        }
    }
    return cudaGetLastError();
}

#ifndef MEAO_EMULATE
cudaError_t preload_blur_upsample()
{
    cudaError_t e = cudaSuccess;
    auto t = [&](auto k) { if (e == cudaSuccess) e = preload_kernel(k); };
    t(blur_upsample_kernel<true, true>); t(blur_upsample_kernel<true, false>); t(blur_upsample_kernel<false, true>); t(blur_upsample_kernel<false, false>);
    t(blur_upsample_premin_kernel<true, true>); t(blur_upsample_premin_kernel<true, false>);
    t(blur_upsample_premin_kernel<false, true>); t(blur_upsample_premin_kernel<false, false>);
    return e;
}
#endif

}  // namespace meao
