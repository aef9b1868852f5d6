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

struct __align__(128) Smem {
    alignas(128) float box_depth[kRawH * kBoxDP];   // TMA destination: low-res depth box (LoResDB)
    alignas(128) uint8_t box_ao[kRawH * kBoxAP];    // TMA destination: low-res AO codes box (LoResAO1)
    alignas(16) float lo_depth[kRawH * kLoDP];      // raw low-res depth, column 0 = virtual column lx0
    alignas(16) float inv_depth[kRawH * kRawP];     // DepthCache, UPS:67-71
    alignas(16) float ao[kRawH * kRawP];            // AOCache1 as loaded, UPS:62-65
    alignas(16) float hblur[kRawH * kBlurP];        // AOCache2, UPS:127-129
    alignas(16) float vblur[kBlurH * kBlurP];       // AOCache1 after the vertical pass, UPS:168-169
    alignas(8) uint64_t bar;
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
    ok = ok & in_safe_range(total.x) & in_safe_range(total.y)
            & ((num.x == 0.0f) | in_safe_range(num.x)) & ((num.y == 0.0f) | in_safe_range(num.y));
    return div2_fast_neg(num, __fmul2_rn(total, m1));
}

#ifndef MEAO_UPS_MINB
#define MEAO_UPS_MINB 5
#endif
template <bool BLEND, bool HI_HALF>
__global__ void __launch_bounds__(kThreads, MEAO_UPS_MINB)
blur_upsample_kernel(const __grid_constant__ CUtensorMap lo_depth_map, const __grid_constant__ CUtensorMap lo_ao_map,
                     const UpsampleArgs a, const int use_tma)
{
#ifdef MEAO_DEVICE_OK
    __shared__ Smem sm;
    const int tid = threadIdx.x;
    const int hx0 = blockIdx.x * kHW;
    const int hy0 = (a.row0 & ~1) + blockIdx.y * kHH;
    const int lx0 = (hx0 >> 1) - 3, ly0 = (hy0 >> 1) - 3;      // virtual low-res coordinate of raw tile (0,0)

    // ---- this thread's 8 hi-res pixels (phase 4); their global loads are issued NOW so that the HBM/L2
    //      latency hides behind the TMA wait and the three blur phases (the kernel was stalling on them)
    // a warp takes rows w, w+8, w+16, w+24 of the tile: the row parity (which selects the operand order of
    // UPS:229-232) is then warp-uniform and the parity branch of phase 4 never diverges
    const int j = tid & 7, hy = ((tid >> 3) & 3) * 8 + (tid >> 5);
    const int py = hy0 + hy, px0 = hx0 + 8 * j;
    const bool active = !(py < a.row0 || py >= a.row1 || px0 >= a.hiw);
    const bool full = (px0 + 8 <= a.hiw);
    uint4 raw_d0 = make_uint4(0, 0, 0, 0), raw_d1 = make_uint4(0, 0, 0, 0);
    uint2 raw_a = make_uint2(0, 0);
    if (active && full) {
        if (HI_HALF) {
            raw_d0 = ldg_stream_u4(reinterpret_cast<const __half *>(a.hi_depth) + (size_t)py * a.hi_dpitch + px0);
        } else {
            const float *src = reinterpret_cast<const float *>(a.hi_depth) + (size_t)py * a.hi_dpitch + px0;
            raw_d0 = ldg_stream_u4(src);
            raw_d1 = ldg_stream_u4(src + 4);
        }
        if (BLEND) raw_a = ldg_stream_u2(a.hi_ao + (size_t)py * a.hi_apitch + px0);
    }

    const bool interior = use_tma && lx0 >= 0 && ly0 >= 0 && (lx0 + kRawW <= a.low) && (ly0 + kRawH <= a.loh);
    if (interior) {
        if (tid == 0) { mbar_init(&sm.bar, 1); fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) {
            mbar_arrive_expect_tx(&sm.bar, (uint32_t)(kRawH * kBoxDP * sizeof(float) + kRawH * kBoxAP));
            tma_load_2d(sm.box_depth, &lo_depth_map, lx0 - kBoxDOff, ly0, &sm.bar);
            tma_load_2d(sm.box_ao, &lo_ao_map, lx0 - kBoxAOff, ly0, &sm.bar);
        }
        mbar_wait(&sm.bar, 0);
    }
    // ---- phase 1: 22 rows x 10 groups of 4 texels: inverse depth (UPS:67) and AO codes -> float (UPS:56)
    if (tid < kRawH * 10) {
        const int r = tid / 10, c0 = (tid - r * 10) * 4;
        float d[4]; uint32_t k[4];
        if (interior) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                d[i] = (c0 + i + kBoxDOff < kBoxDP) ? sm.box_depth[r * kBoxDP + c0 + i + kBoxDOff] : 1.0f;   // column 39 is never consumed
                k[i] = sm.box_ao[r * kBoxAP + c0 + i + kBoxAOff];
            }
        } else {    // border tile: point + clamp addressing (UPS:56,67)
            const int sy = iclamp(ly0 + r, 0, a.loh - 1);
            const float *drow = a.lo_depth + (size_t)sy * a.lo_dpitch;
            const uint8_t *arow = a.lo_ao + (size_t)sy * a.lo_apitch;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int sx = iclamp(lx0 + c0 + i, 0, a.low - 1);
                d[i] = __ldg(drow + sx);
                k[i] = __ldg(arow + sx);
            }
        }
        *reinterpret_cast<float2 *>(&sm.lo_depth[r * kLoDP + c0]) = make_float2(d[0], d[1]);
        *reinterpret_cast<float2 *>(&sm.lo_depth[r * kLoDP + c0 + 2]) = make_float2(d[2], d[3]);
        *reinterpret_cast<float4 *>(&sm.inv_depth[r * kRawP + c0]) = make_float4(rcp_ieee(d[0]), rcp_ieee(d[1]), rcp_ieee(d[2]), rcp_ieee(d[3]));
        *reinterpret_cast<float4 *>(&sm.ao[r * kRawP + c0]) = make_float4(unorm8_load(k[0]), unorm8_load(k[1]), unorm8_load(k[2]), unorm8_load(k[3]));
    }
    __syncthreads();

    const float step = a.step_size, kblur = a.blur_tolerance;

    // ---- horizontal blur, UPS:89-130: 11 row pairs (the two packed lanes) x 9 runs of 4 outputs;
    //      output c is centred on raw column c+2
    if (tid < (kRawH / 2) * 9) {
        const int rp = tid / 9, c0 = (tid - rp * 9) * 4, r = 2 * rp;
        float2 av[8], dv[8], o[4];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const float4 A0 = *reinterpret_cast<const float4 *>(&sm.ao[r * kRawP + c0 + 4 * q]);
            const float4 A1 = *reinterpret_cast<const float4 *>(&sm.ao[(r + 1) * kRawP + c0 + 4 * q]);
            const float4 D0 = *reinterpret_cast<const float4 *>(&sm.inv_depth[r * kRawP + c0 + 4 * q]);
            const float4 D1 = *reinterpret_cast<const float4 *>(&sm.inv_depth[(r + 1) * kRawP + c0 + 4 * q]);
            av[4 * q] = make_float2(A0.x, A1.x); av[4 * q + 1] = make_float2(A0.y, A1.y);
            av[4 * q + 2] = make_float2(A0.z, A1.z); av[4 * q + 3] = make_float2(A0.w, A1.w);
            dv[4 * q] = make_float2(D0.x, D1.x); dv[4 * q + 1] = make_float2(D0.y, D1.y);
            dv[4 * q + 2] = make_float2(D0.z, D1.z); dv[4 * q + 3] = make_float2(D0.w, D1.w);
        }
        blur_run2<4>(av, dv, step, kblur, o);
        *reinterpret_cast<float4 *>(&sm.hblur[r * kBlurP + c0]) = make_float4(o[0].x, o[1].x, o[2].x, o[3].x);
        *reinterpret_cast<float4 *>(&sm.hblur[(r + 1) * kBlurP + c0]) = make_float4(o[0].y, o[1].y, o[2].y, o[3].y);
    }
    __syncthreads();

    // ---- vertical blur, UPS:132-170: 17 column pairs (the two packed lanes) x 3 runs of 6 outputs;
    //      output r centred on row r+2, depth column offset +2 (UPS:141-146)
    if (tid < (kBlurW / 2) * 3) {
        const int run = tid / (kBlurW / 2), c = 2 * (tid - run * (kBlurW / 2)), r0 = run * 6;
        float2 av[10], dv[10], o[6];
#pragma unroll
        for (int i = 0; i < 10; i++) {
            av[i] = *reinterpret_cast<const float2 *>(&sm.hblur[(r0 + i) * kBlurP + c]);
            dv[i] = *reinterpret_cast<const float2 *>(&sm.inv_depth[(r0 + i) * kRawP + c + 2]);
        }
        blur_run2<6>(av, dv, step, kblur, o);
#pragma unroll
        for (int i = 0; i < 6; i++) *reinterpret_cast<float2 *>(&sm.vblur[(r0 + i) * kBlurP + c]) = o[i];
    }
    __syncthreads();

    // ---- bilateral upsample, UPS:213-232: thread -> 8 consecutive hi-res pixels of one row ----------
    if (!active) return;

    // blurred index of X-1 for the first pixel is 4j; quad rows: rY-1, rY with rY = ((hy+1)>>1)+1
    const int rY = ((hy + 1) >> 1) + 1;
    float bl_ao[2][6], lo_d[2][6];      // [0] = row Y-1 (top), [1] = row Y (bottom)
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        const float *vb = &sm.vblur[(rY - 1 + rr) * kBlurP + 4 * j];
        const float *ld = &sm.lo_depth[(rY - 1 + rr + 2) * kLoDP + 4 * j + 2];
        const float4 v4 = *reinterpret_cast<const float4 *>(vb);
        const float2 v2 = *reinterpret_cast<const float2 *>(vb + 4);
        bl_ao[rr][0] = v4.x; bl_ao[rr][1] = v4.y; bl_ao[rr][2] = v4.z; bl_ao[rr][3] = v4.w; bl_ao[rr][4] = v2.x; bl_ao[rr][5] = v2.y;
        const float2 d0 = *reinterpret_cast<const float2 *>(ld);
        const float2 d1 = *reinterpret_cast<const float2 *>(ld + 2);
        const float2 d2 = *reinterpret_cast<const float2 *>(ld + 4);
        lo_d[rr][0] = d0.x; lo_d[rr][1] = d0.y; lo_d[rr][2] = d1.x; lo_d[rr][3] = d1.y; lo_d[rr][4] = d2.x; lo_d[rr][5] = d2.y;
    }

    float hd[8], ha[8];
    if (HI_HALF) {
        const __half *src = reinterpret_cast<const __half *>(a.hi_depth) + (size_t)py * a.hi_dpitch + px0;
        if (full) {
            const __half2 *h = reinterpret_cast<const __half2 *>(&raw_d0);
#pragma unroll
            for (int e = 0; e < 4; e++) { const float2 f = __half22float2(h[e]); hd[2 * e] = f.x; hd[2 * e + 1] = f.y; }
        } else {
#pragma unroll
            for (int e = 0; e < 8; e++) hd[e] = (px0 + e < a.hiw) ? __half2float(src[e]) : 1.0f;
        }
    } else {
        const float *src = reinterpret_cast<const float *>(a.hi_depth) + (size_t)py * a.hi_dpitch + px0;
        if (full) {
            const float4 q0 = *reinterpret_cast<const float4 *>(&raw_d0), q1 = *reinterpret_cast<const float4 *>(&raw_d1);
            hd[0] = q0.x; hd[1] = q0.y; hd[2] = q0.z; hd[3] = q0.w; hd[4] = q1.x; hd[5] = q1.y; hd[6] = q1.z; hd[7] = q1.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; e++) hd[e] = (px0 + e < a.hiw) ? __ldg(src + e) : 1.0f;
        }
    }
    if (BLEND) {
        const uint8_t *src = a.hi_ao + (size_t)py * a.hi_apitch + px0;
        if (full) {
            const uint2 q = raw_a;
#pragma unroll
            for (int e = 0; e < 4; e++) { ha[e] = unorm8_load((q.x >> (8 * e)) & 0xffu); ha[4 + e] = unorm8_load((q.y >> (8 * e)) & 0xffu); }
        } else {
#pragma unroll
            for (int e = 0; e < 8; e++) ha[e] = (px0 + e < a.hiw) ? unorm8_load(__ldg(src + e)) : 1.0f;
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; e++) ha[e] = 1.0f;                                            // UPS:223
    }

    const float tol = a.upsample_tolerance, nfs = a.noise_filter_strength;
    const bool y_odd = (py & 1) != 0;     // py = 2Y-1 (odd) or 2Y (even)
    uint32_t code[8];
    // X - 1 -> local index m, X -> m + 1, with m = (e + 1) >> 1 relative to blurred column 4j
#define MEAO_UPS_PIXELS(FAST, YODD, OK)                                                                                         \
    _Pragma("unroll") for (int e = 0; e < 8; e++) {                                                                             \
        const int m = (e + 1) >> 1;                                                                                             \
        const float tl_d = lo_d[0][m], tr_d = lo_d[0][m + 1], bl_d = lo_d[1][m], br_d = lo_d[1][m + 1];                         \
        const float tl_a = bl_ao[0][m], tr_a = bl_ao[0][m + 1], bl_a = bl_ao[1][m], br_a = bl_ao[1][m + 1];                     \
        float r;                                                                                                                \
        if ((e & 1) != 0) { /* px odd = 2X-1 */                                                                                 \
            if (!(YODD)) r = bilateral<FAST, BLEND>(hd[e], ha[e], bl_d, br_d, tr_d, tl_d, bl_a, br_a, tr_a, tl_a, tol, nfs, OK); /* UPS:229 (-1, 0) .xyzw */ \
            else         r = bilateral<FAST, BLEND>(hd[e], ha[e], tl_d, bl_d, br_d, tr_d, tl_a, bl_a, br_a, tr_a, tol, nfs, OK); /* UPS:232 (-1,-1) .wxyz */ \
        } else {            /* px even = 2X */                                                                                  \
            if (!(YODD)) r = bilateral<FAST, BLEND>(hd[e], ha[e], br_d, tr_d, tl_d, bl_d, br_a, tr_a, tl_a, bl_a, tol, nfs, OK); /* UPS:230 ( 0, 0) .yzwx */ \
            else         r = bilateral<FAST, BLEND>(hd[e], ha[e], tr_d, tl_d, bl_d, br_d, tr_a, tl_a, bl_a, br_a, tol, nfs, OK); /* UPS:231 ( 0,-1) .zwxy */ \
        }                                                                                                                       \
        code[e] = unorm8_code(r);                                                                                               \
    }
    bool ok = a.fast_div_ok != 0;
    if (ok) {
        // packed fast path: pixel pairs (0,2) (1,3) (4,6) (5,7); blurred column of X-1 is m = (e+1)>>1
#define MEAO_UPS_PAIR(E, YODD)                                                                                                  \
        {                                                                                                                       \
            constexpr int ma = ((E) + 1) >> 1, mb = ((E) + 3) >> 1;                                                             \
            const float2 tl_d = make_float2(lo_d[0][ma], lo_d[0][mb]), tr_d = make_float2(lo_d[0][ma + 1], lo_d[0][mb + 1]);   \
            const float2 bl_d = make_float2(lo_d[1][ma], lo_d[1][mb]), br_d = make_float2(lo_d[1][ma + 1], lo_d[1][mb + 1]);   \
            const float2 tl_a = make_float2(bl_ao[0][ma], bl_ao[0][mb]), tr_a = make_float2(bl_ao[0][ma + 1], bl_ao[0][mb + 1]); \
            const float2 bl_a = make_float2(bl_ao[1][ma], bl_ao[1][mb]), br_a = make_float2(bl_ao[1][ma + 1], bl_ao[1][mb + 1]); \
            const float2 hd2 = make_float2(hd[E], hd[(E) + 2]), ha2 = make_float2(ha[E], ha[(E) + 2]);                          \
            float2 r;                                                                                                           \
            if (((E) & 1) != 0) {                                                                                               \
                if (!(YODD)) r = bilateral2<BLEND>(hd2, ha2, bl_d, br_d, tr_d, tl_d, bl_a, br_a, tr_a, tl_a, tol, nfs, ok);     \
                else         r = bilateral2<BLEND>(hd2, ha2, tl_d, bl_d, br_d, tr_d, tl_a, bl_a, br_a, tr_a, tol, nfs, ok);     \
            } else {                                                                                                            \
                if (!(YODD)) r = bilateral2<BLEND>(hd2, ha2, br_d, tr_d, tl_d, bl_d, br_a, tr_a, tl_a, bl_a, tol, nfs, ok);     \
                else         r = bilateral2<BLEND>(hd2, ha2, tr_d, tl_d, bl_d, br_d, tr_a, tl_a, bl_a, br_a, tol, nfs, ok);     \
            }                                                                                                                   \
            code[E] = unorm8_code(r.x);                                                                                         \
            code[(E) + 2] = unorm8_code(r.y);                                                                                   \
        }
        if (y_odd) { MEAO_UPS_PAIR(0, true) MEAO_UPS_PAIR(1, true) MEAO_UPS_PAIR(4, true) MEAO_UPS_PAIR(5, true) }
        else       { MEAO_UPS_PAIR(0, false) MEAO_UPS_PAIR(1, false) MEAO_UPS_PAIR(4, false) MEAO_UPS_PAIR(5, false) }
#undef MEAO_UPS_PAIR
    }
    if (!ok) {      // rare: inf / NaN / zero / denormal operands somewhere in this thread's 8 pixels -> plain IEEE operators
        bool unused = true;
        MEAO_UPS_PIXELS(false, y_odd, unused)
    }
#undef MEAO_UPS_PIXELS

    uint8_t *dst = a.out + (size_t)(py - a.out_row_origin) * a.out_pitch + px0;
    if (full && a.out_vec_ok) {
        uint2 pk;
        pk.x = code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24);
        pk.y = code[4] | (code[5] << 8) | (code[6] << 16) | (code[7] << 24);
        *reinterpret_cast<uint2 *>(dst) = pk;
    } else {
#pragma unroll
        for (int e = 0; e < 8; e++)
            if (px0 + e < a.hiw) dst[e] = (uint8_t)code[e];
    }
#endif
}

}  // namespace

cudaError_t launch_blur_upsample(const CUtensorMap &lo_depth_map, const CUtensorMap &lo_ao_map, bool use_tma,
                                 const UpsampleArgs &a, cudaStream_t s)
{
    if (a.row1 <= a.row0) return cudaSuccess;
    const int ybase = a.row0 & ~1;
    dim3 grid(ceil_div(a.hiw, kHW), ceil_div(a.row1 - ybase, kHH));
    const int t = use_tma ? 1 : 0;
    if (a.hi_ao) {
        if (a.hi_is_half) blur_upsample_kernel<true, true><<<grid, kThreads, 0, s>>>(lo_depth_map, lo_ao_map, a, t);
        else              blur_upsample_kernel<true, false><<<grid, kThreads, 0, s>>>(lo_depth_map, lo_ao_map, a, t);
    } else {
        if (a.hi_is_half) blur_upsample_kernel<false, true><<<grid, kThreads, 0, s>>>(lo_depth_map, lo_ao_map, a, t);
        else              blur_upsample_kernel<false, false><<<grid, kThreads, 0, s>>>(lo_depth_map, lo_ao_map, a, t);
    }
    return cudaGetLastError();
}

}  // namespace meao
