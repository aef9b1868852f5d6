// AmbientOcclusionScalar.cs -- the scalar C# CPU implementation of the three SSAO compute stages that BASELINE.json's
// north_star asks to time next to the GPU path ("a scalar C# CPU implementation of the same three stages").
//
// SOURCE ONLY.  This image (and the GPU box, which runs the same image) has no C# toolchain -- dotnet, mono, mcs and csc are
// all absent (SURVEY.md 8c) -- so this file has never been compiled or run here.  It is the line-for-line C# twin of
// oracle/meao_oracle.c, which IS compiled, pinned by the CPU tests and timed by bench.py as the cpu_baseline (kind "port").
// Keep the two in step: every method below names the oracle function and the reference lines it restates
// (paths relative to Assets/MiniEngineAO/).
//
// It follows the reference thread-group structure literally (group ids, group-shared arrays, barriers), one thread at a
// time; `threads` > 1 stripes the thread-group rows over System.Threading.Tasks.Parallel.For (results are identical for
// any thread count -- groups are independent).
//
// Fixed-function conventions (D3D11 behaviour the HLSL relies on; identical to oracle/meao_oracle.h):
//   * fp32 round-to-nearest-even; a*b+c written in one HLSL expression is a fused mad -> Mad() below;
//   * x / y is IEEE; f32 -> f16 store RTNE with overflow to +inf; f32 -> UNORM8 store = (uint)(saturate(x) * 255 + 0.5),
//     NaN -> 0; UNORM8 load = k * (1/255); out-of-bounds load -> 0, out-of-bounds store dropped;
//   * Gather = point + clamp; footprint of texel-corner c is texels (c-1, c); .w=(x0,y0) .z=(x1,y0) .x=(x0,y1) .y=(x1,y1);
//   * saturate(NaN) = 0; min / max return the non-NaN operand.
// C# note: every intermediate is cast to (float) so that no runtime may keep excess precision (ECMA-335 I.12.1.3).

using System;
using System.Diagnostics;
using System.Threading.Tasks;

namespace MiniEngineAO.ScalarCpu
{
    public struct AoParams          // AmbientOcclusion.cs:20-58
    {
        public float noiseFilterTolerance, blurTolerance, upsampleTolerance, thicknessModifier, intensity;
        public static AoParams Default()
        {
            return new AoParams { noiseFilterTolerance = 0f, blurTolerance = -4.6f, upsampleTolerance = -12f, thicknessModifier = 1f, intensity = 1f };
        }
    }

    public struct AoCamera          // AmbientOcclusion.cs:561-573
    {
        public float nearClip, farClip, tanHalfFovH;
        public bool reversedZ;
    }

    public sealed class AoScalar
    {
        public readonly int W, H;
        public readonly int[] lw = new int[7], lh = new int[7];     // AmbientOcclusion.cs:276-281
        public AoParams p = AoParams.Default();
        public AoCamera cam = new AoCamera { nearClip = 0.3f, farClip = 100f, tanHalfFovH = 1f, reversedZ = true };
        public int threads = 1;
        public bool singleScale;        // BASELINE.json configs[0]: Downsample1 -> Render level 1 -> the final-style Upsample fed with Occlusion1

        // post-quantisation values: an f16 buffer holds floats representable in f16, a UNORM8 buffer holds k * (1/255)
        public float[] linearDepth;                         // id 1       L0      f16
        public readonly float[][] lowDepth = new float[5][];    // id 2..5    L1..L4  f32      [1..4]
        public readonly float[][] tiledDepth = new float[5][];  // id 6..9    L3..L6 x 16 slices, f16   [1..4]
        public readonly float[][] occlusion = new float[5][];   // id 10..13  L1..L4  unorm8   [1..4]
        public readonly float[][] combined = new float[4][];    // id 14..16  L1..L3  unorm8   [1..3]
        public float[] result;                              // id 17      L0      unorm8

        public AoScalar(int width, int height)              // meao_oracle_create
        {
            W = width; H = height;
            for (int l = 0; l < 7; l++) { int div = 1 << l; lw[l] = (width + (div - 1)) / div; lh[l] = (height + (div - 1)) / div; }
            linearDepth = new float[lw[0] * lh[0]];
            result = new float[lw[0] * lh[0]];
            for (int k = 1; k <= 4; k++)
            {
                lowDepth[k] = new float[lw[k] * lh[k]];
                tiledDepth[k] = new float[16 * lw[k + 2] * lh[k + 2]];
                occlusion[k] = new float[lw[k] * lh[k]];
                if (k <= 3) combined[k] = new float[lw[k] * lh[k]];
            }
        }

        // ---- arithmetic conventions ------------------------------------------------------------------------------
#if NETCOREAPP3_0_OR_GREATER
        static float Mad(float a, float b, float c) { return MathF.FusedMultiplyAdd(a, b, c); }
#else
        // the product of two floats is exact in double; the single rounding of the sum to double can double-round on
        // rare ties -- use a runtime with MathF.FusedMultiplyAdd for bit parity with the oracle
        static float Mad(float a, float b, float c) { return (float)((double)a * (double)b + (double)c); }
#endif
        static float Sat(float x) { return x > 0f ? (x < 1f ? x : 1f) : 0f; }                       // NaN -> 0
        static float HMax(float a, float b) { return float.IsNaN(a) ? b : (float.IsNaN(b) ? a : (a > b ? a : b)); }
        static float HMin(float a, float b) { return float.IsNaN(a) ? b : (float.IsNaN(b) ? a : (a < b ? a : b)); }
        static float HClamp(float x, float lo, float hi) { return HMin(HMax(x, lo), hi); }
        static int IClamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

        // ---- storage conversions (formats: AmbientOcclusion.cs:262-273) ---------------------------------------------
        public static ushort F32ToF16Bits(float x)          // meao_oracle_f32_to_f16_bits: RTNE, overflow -> inf
        {
            uint u = (uint)BitConverter.SingleToInt32Bits(x);
            uint sign = (u >> 16) & 0x8000u, absu = u & 0x7fffffffu;
            if (absu > 0x7f800000u) return (ushort)(sign | 0x7e00u);
            if (absu >= 0x477ff000u) return (ushort)(sign | 0x7c00u);      // >= 65520 rounds to inf (65520 is the tie -> even = inf)
            if (absu < 0x33000000u) return (ushort)sign;                   // < 2^-25 -> 0
            int e = (int)(absu >> 23) - 127;
            uint m = (absu & 0x7fffffu) | 0x800000u;
            int shift; uint bas;
            if (e >= -14) { shift = 13; bas = (uint)(e + 15) << 10; m &= 0x7fffffu; }
            else { shift = 13 + (-14 - e); bas = 0; }
            if (shift > 24) return (ushort)sign;
            uint q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
            if (rem > half || (rem == half && (q & 1u) != 0)) q++;
            return (ushort)(sign | (bas + q));
        }

        public static float F16BitsToF32(ushort h)          // meao_oracle_f16_bits_to_f32
        {
            uint sign = (uint)(h & 0x8000u) << 16, e = (uint)(h >> 10) & 0x1fu, m = (uint)h & 0x3ffu, u;
            if (e == 0)
            {
                if (m == 0) u = sign;
                else { float f = (float)m * 5.9604644775390625e-8f; u = (uint)BitConverter.SingleToInt32Bits(f) | sign; }
            }
            else if (e == 31) u = sign | 0x7f800000u | (m << 13);
            else u = sign | ((e + 112u) << 23) | (m << 13);
            return BitConverter.Int32BitsToSingle((int)u);
        }

        static float StHalf(float x) { return F16BitsToF32(F32ToF16Bits(x)); }
        public static byte Unorm8Code(float x) { float c = Sat(x); return (byte)(uint)((float)(c * 255.0f) + 0.5f); }
        static float StUnorm8(float x) { return (float)Unorm8Code(x) * (1.0f / 255.0f); }

        // ---- CPU-side constants ------------------------------------------------------------------------------------
        public float[] ZBufferParams()                      // AmbientOcclusion.cs:561-568
        {
            float fpn = cam.farClip / cam.nearClip;
            return cam.reversedZ ? new[] { fpn - 1f, 1f, 0f, 0f } : new[] { 1f - fpn, fpn, 0f, 0f };
        }

        static float MathfSqrt(float f) { return (float)Math.Sqrt(f); }
        static float MathfPow(float f, float q) { return (float)Math.Pow(f, q); }

        public static float[] SampleThickness()             // AmbientOcclusion.cs:577-590
        {
            return new[] {
                MathfSqrt(1 - 0.2f * 0.2f), MathfSqrt(1 - 0.4f * 0.4f), MathfSqrt(1 - 0.6f * 0.6f), MathfSqrt(1 - 0.8f * 0.8f),
                MathfSqrt(1 - 0.2f * 0.2f - 0.2f * 0.2f), MathfSqrt(1 - 0.2f * 0.2f - 0.4f * 0.4f),
                MathfSqrt(1 - 0.2f * 0.2f - 0.6f * 0.6f), MathfSqrt(1 - 0.2f * 0.2f - 0.8f * 0.8f),
                MathfSqrt(1 - 0.4f * 0.4f - 0.4f * 0.4f), MathfSqrt(1 - 0.4f * 0.4f - 0.6f * 0.6f),
                MathfSqrt(1 - 0.4f * 0.4f - 0.8f * 0.8f), MathfSqrt(1 - 0.6f * 0.6f - 0.6f * 0.6f) };
        }

        // AmbientOcclusion.cs:660-734 for source = TiledDepth<level> (mip level + 2, tiled)
        void RenderConstants(int level, float[] invThickness, float[] sampleWeight, out float rejectFadeoff, out float intensity)
        {
            float[] thick = SampleThickness();
            const float ScreenspaceDiameter = 10;                                                   // :669
            float thicknessMultiplier = 2 * cam.tanHalfFovH * ScreenspaceDiameter / (float)lw[level + 2];  // :678
            float inverseRangeFactor = 1 / thicknessMultiplier;                                     // :683
            for (int i = 0; i < 12; i++) invThickness[i] = inverseRangeFactor / thick[i];           // :687-688
            float[] mult = { 4, 4, 4, 4, 4, 8, 8, 8, 4, 8, 8, 4 };                                  // :696-707
            for (int i = 0; i < 12; i++) sampleWeight[i] = mult[i] * thick[i];
            sampleWeight[0] = 0; sampleWeight[2] = 0; sampleWeight[5] = 0; sampleWeight[7] = 0; sampleWeight[9] = 0;   // :711-715
            float total = 0.0f;
            for (int i = 0; i < 12; i++) total += sampleWeight[i];                                  // :718-721
            for (int i = 0; i < 12; i++) sampleWeight[i] /= total;                                  // :723-724
            rejectFadeoff = -1 / p.thicknessModifier;                                               // :733
            intensity = p.intensity;                                                                // :734
        }

        // AmbientOcclusion.cs:757-771
        void UpsampleConstants(int loLevel, out float noiseFilterStrength, out float stepSize, out float kBlurTolerance, out float kUpsampleTolerance)
        {
            stepSize = 1920.0f / (float)lw[loLevel];                                                // :760
            float blurTolerance = 1 - MathfPow(10, p.blurTolerance) * stepSize;                     // :761
            kBlurTolerance = blurTolerance * blurTolerance;                                         // :762
            kUpsampleTolerance = MathfPow(10, p.upsampleTolerance);                                 // :763
            noiseFilterStrength = 1 / (MathfPow(10, p.noiseFilterTolerance) + kUpsampleTolerance);  // :764
        }

        void Striped(int n, Action<int, int> body)          // run_striped
        {
            int t = Math.Max(1, Math.Min(threads, n));
            if (t == 1) { body(0, n); return; }
            Parallel.For(0, t, i => body((int)((long)n * i / t), (int)((long)n * (i + 1) / t)));
        }

        // ---- Downsample1.compute / Downsample2.compute -----------------------------------------------------------------
        float Linearize(float[] depth, float[] zb, int x, int y)     // ds1_linearize, Downsample1.compute:37-48
        {
            bool inb = x < W && y < H;
            float d = inb ? depth[y * W + x] : 0.0f;                 // :39, OOB load -> 0
            float dist = 1.0f / Mad(zb[0], d, zb[1]);                // :40
            if (cam.reversedZ) { if (d == 0) dist = 1e5f; }          // :41-42
            else { if (d == 1) dist = 1e5f; }                        // :43-44
            if (inb) linearDepth[y * W + x] = StHalf(dist);          // :46, OOB store dropped
            return dist;
        }

        static int SliceOf(int sx, int sy) { return ((sx & 3) | (sy << 2)) & 15; }     // Downsample1.compute:69

        public void Downsample(float[] depth)               // meao_oracle_downsample, AmbientOcclusion.cs:604-658
        {
            float[] zb = ZBufferParams();
            int L1w = lw[1], L1h = lh[1], L2w = lw[2], L2h = lh[2], A1w = lw[3], A1h = lh[3], A2w = lw[4], A2h = lh[4];
            // Downsample1.compute:52-81, dispatch (tiled2.w, tiled2.h, 1) AmbientOcclusion.cs:643
            Striped(lh[4], (gy0, gy1) =>
            {
                float[] cache = new float[256];                                              // :50 g_CacheW
                for (int gy = gy0; gy < gy1; gy++)
                for (int gx = 0; gx < lw[4]; gx++)
                {
                    for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++)
                    {
                        int sx = (gx << 4) | tx, sy = (gy << 4) | ty, dest = (ty << 4) | tx;  // :55-56
                        cache[dest + 0] = Linearize(depth, zb, sx | 0, sy | 0);               // :57-60
                        cache[dest + 8] = Linearize(depth, zb, sx | 8, sy | 0);
                        cache[dest + 128] = Linearize(depth, zb, sx | 0, sy | 8);
                        cache[dest + 136] = Linearize(depth, zb, sx | 8, sy | 8);
                    }
                    // GroupMemoryBarrierWithGroupSync :62
                    for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++)
                    {
                        int GI = ty * 8 + tx;
                        float w1 = cache[(tx << 1) | (ty << 5)];                              // :64-66
                        int stx = gx * 8 + tx, sty = gy * 8 + ty;                             // :68
                        int slice = SliceOf(stx, sty);
                        if (stx < L1w && sty < L1h) lowDepth[1][sty * L1w + stx] = w1;        // :70
                        if ((stx >> 2) < A1w && (sty >> 2) < A1h)                             // :71
                            tiledDepth[1][(slice * A1h + (sty >> 2)) * A1w + (stx >> 2)] = StHalf(w1);
                        if ((GI & 9) == 0)                                                    // :73  (011 is OCTAL = 9)
                        {
                            int s2x = stx >> 1, s2y = sty >> 1;                               // :75
                            slice = SliceOf(s2x, s2y);
                            if (s2x < L2w && s2y < L2h) lowDepth[2][s2y * L2w + s2x] = w1;    // :77
                            if ((s2x >> 2) < A2w && (s2y >> 2) < A2h)                         // :78
                                tiledDepth[2][(slice * A2h + (s2y >> 2)) * A2w + (s2x >> 2)] = StHalf(w1);
                        }
                    }
                }
            });
            // Downsample2.compute:32-51, dispatch (tiled4.w, tiled4.h, 1) AmbientOcclusion.cs:657
            int L3w = lw[3], L3h = lh[3], L4w = lw[4], L4h = lh[4], A3w = lw[5], A3h = lh[5], A4w = lw[6], A4h = lh[6];
            Striped(lh[6], (gy0, gy1) =>
            {
                for (int gy = gy0; gy < gy1; gy++)
                for (int gx = 0; gx < lw[6]; gx++)
                for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++)
                {
                    int GI = ty * 8 + tx, stx = gx * 8 + tx, sty = gy * 8 + ty, rx = stx << 1, ry = sty << 1;
                    float m1 = (rx < L2w && ry < L2h) ? lowDepth[2][ry * L2w + rx] : 0.0f;    // :35
                    int slice = SliceOf(stx, sty);                                            // :37-39
                    if (stx < L3w && sty < L3h) lowDepth[3][sty * L3w + stx] = m1;            // :40
                    if ((stx >> 2) < A3w && (sty >> 2) < A3h)                                 // :41
                        tiledDepth[3][(slice * A3h + (sty >> 2)) * A3w + (stx >> 2)] = StHalf(m1);
                    if ((GI & 9) == 0)                                                        // :43
                    {
                        int s2x = stx >> 1, s2y = sty >> 1;
                        slice = SliceOf(s2x, s2y);
                        if (s2x < L4w && s2y < L4h) lowDepth[4][s2y * L4w + s2x] = m1;        // :48
                        if ((s2x >> 2) < A4w && (s2y >> 2) < A4h)                             // :49
                            tiledDepth[4][(slice * A4h + (s2y >> 2)) * A4w + (s2x >> 2)] = StHalf(m1);
                    }
                }
            });
        }

        // ---- Gather: point + clamp; returns (x, y, z, w) ----------------------------------------------------------------
        static void Gather4(float[] buf, int off, int w, int h, int cx, int cy, out float gx, out float gy, out float gz, out float gw)
        {
            int x0 = IClamp(cx - 1, 0, w - 1), x1 = IClamp(cx, 0, w - 1), y0 = IClamp(cy - 1, 0, h - 1), y1 = IClamp(cy, 0, h - 1);
            gw = buf[off + y0 * w + x0]; gz = buf[off + y0 * w + x1]; gx = buf[off + y1 * w + x0]; gy = buf[off + y1 * w + x1];
        }

        // ---- Render.compute, kernel main_interleaved (TILE_DIM 16, 8x8 threads) -------------------------------------------
        const int T = 16;                                                                       // Render.compute:53

        static float TestSamplePair(float[] DS, float rf, float frontDepth, float invRange, int bas, int offset)   // :60-75
        {
            float d1 = Mad(DS[bas + offset], invRange, -frontDepth);                            // :65
            float d2 = Mad(DS[bas - offset], invRange, -frontDepth);                            // :66
            float p1 = Sat((float)(rf * d1)), p2 = Sat((float)(rf * d2));                       // :68-69
            float s = (float)(HClamp(d1, p2, 1.0f) + HClamp(d2, p1, 1.0f));
            return Sat(Mad(-p1, p2, s));                                                        // :71-74
        }

        static float TestSamples(float[] DS, float rf, int centerIdx, int x, int y, float invDepth, float invThickness)   // :77-110
        {
            float invRange = (float)(invThickness * invDepth);                                  // :84
            float frontDepth = (float)(invThickness - 0.5f);                                    // :85
            if (y == 0)
                return (float)(0.5f * (float)(TestSamplePair(DS, rf, frontDepth, invRange, centerIdx, x) +
                                              TestSamplePair(DS, rf, frontDepth, invRange, centerIdx, x * T)));
            if (x == y)
                return (float)(0.5f * (float)(TestSamplePair(DS, rf, frontDepth, invRange, centerIdx, x * T - x) +
                                              TestSamplePair(DS, rf, frontDepth, invRange, centerIdx, x * T + x)));
            float a = TestSamplePair(DS, rf, frontDepth, invRange, centerIdx, y * T + x);
            float b = TestSamplePair(DS, rf, frontDepth, invRange, centerIdx, y * T - x);
            float c = TestSamplePair(DS, rf, frontDepth, invRange, centerIdx, x * T + y);
            float d = TestSamplePair(DS, rf, frontDepth, invRange, centerIdx, x * T - y);
            return (float)(0.25f * (float)((float)((float)(a + b) + c) + d));
        }

        public void Render(int k)                           // meao_oracle_render, Render.compute:112-177, AmbientOcclusion.cs:660-748
        {
            float[] iT = new float[12], sW = new float[12];
            float rf, intensity;
            RenderConstants(k, iT, sW, out rf, out intensity);
            int sw = lw[k + 2], sh = lh[k + 2], ow = lw[k], oh = lh[k], ngx = (sw + 7) / 8, ngy = (sh + 7) / 8;
            Striped(16 * ngy, (r0, r1) =>                    // dispatch ceil(w/8) x ceil(h/8) x 16, :739-747
            {
                float[] DS = new float[T * T];               // Render.compute:58
                for (int r = r0; r < r1; r++)
                {
                    int z = r / ngy, gy = r % ngy, sliceOff = z * sw * sh;
                    for (int gx = 0; gx < ngx; gx++)
                    {
                        for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++)
                        {
                            float dx, dy, dz, dw;
                            Gather4(tiledDepth[k], sliceOff, sw, sh, gx * 8 + tx + tx - 3, gy * 8 + ty + ty - 3, out dx, out dy, out dz, out dw);   // :118,123
                            int dest = tx * 2 + ty * 2 * T;                                     // :127
                            DS[dest] = dw; DS[dest + 1] = dz; DS[dest + T] = dx; DS[dest + T + 1] = dy;   // :128-131
                        }
                        // GroupMemoryBarrierWithGroupSync :133
                        for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++)
                        {
                            int thisIdx = tx + ty * T + 4 * T + 4;                              // :138
                            float invThisDepth = 1.0f / DS[thisIdx];                            // :140
                            float ao = 0.0f;                                                    // :142
                            ao = Mad(sW[1], TestSamples(DS, rf, thisIdx, 2, 0, invThisDepth, iT[1]), ao);     // :162-168
                            ao = Mad(sW[3], TestSamples(DS, rf, thisIdx, 4, 0, invThisDepth, iT[3]), ao);
                            ao = Mad(sW[4], TestSamples(DS, rf, thisIdx, 1, 1, invThisDepth, iT[4]), ao);
                            ao = Mad(sW[8], TestSamples(DS, rf, thisIdx, 2, 2, invThisDepth, iT[8]), ao);
                            ao = Mad(sW[11], TestSamples(DS, rf, thisIdx, 3, 3, invThisDepth, iT[11]), ao);
                            ao = Mad(sW[6], TestSamples(DS, rf, thisIdx, 1, 3, invThisDepth, iT[6]), ao);
                            ao = Mad(sW[10], TestSamples(DS, rf, thisIdx, 2, 4, invThisDepth, iT[10]), ao);
                            int ox = ((gx * 8 + tx) << 2) | (z & 3), oy = ((gy * 8 + ty) << 2) | (z >> 2);    // :172
                            if (ox < ow && oy < oh)
                                occlusion[k][oy * ow + ox] = StUnorm8(Mad(intensity, (float)(ao - 1.0f), 1.0f));   // :176 lerp(1, ao, I)
                        }
                    }
                }
            });
        }

        // ---- Upsample.compute, kernels main / main_blendout ------------------------------------------------------------
        static float SmartBlur(float a, float b, float c, float d, float e, bool left, bool middle, bool right)   // :74-81
        {
            b = (left | middle) ? b : c;
            a = left ? a : b;
            d = (right | middle) ? d : c;
            e = right ? e : d;
            return (float)((float)((float)((float)((float)((float)(a + e) / 2.0f) + b) + c) + d) / 4.0f);
        }

        static bool CompareDeltas(float stepSize, float kBlurTolerance, float d1, float d2, float l1, float l2)   // :83-87
        {
            float temp = Mad(d1, d2, stepSize);
            return (float)(temp * temp) > (float)((float)(l1 * l2) * kBlurTolerance);
        }

        static void BlurHorizontally(float S, float kB, float[] AO1, float[] DC, float[] AO2, int i)              // :89-130
        {
            float d01 = (float)(DC[i + 1] - DC[i]), d12 = (float)(DC[i + 2] - DC[i + 1]), d23 = (float)(DC[i + 3] - DC[i + 2]);
            float d34 = (float)(DC[i + 4] - DC[i + 3]), d45 = (float)(DC[i + 5] - DC[i + 4]), d56 = (float)(DC[i + 6] - DC[i + 5]);
            float l01 = Mad(d01, d01, S), l12 = Mad(d12, d12, S), l23 = Mad(d23, d23, S), l34 = Mad(d34, d34, S), l45 = Mad(d45, d45, S), l56 = Mad(d56, d56, S);
            bool c02 = CompareDeltas(S, kB, d01, d12, l01, l12), c13 = CompareDeltas(S, kB, d12, d23, l12, l23);
            bool c24 = CompareDeltas(S, kB, d23, d34, l23, l34), c35 = CompareDeltas(S, kB, d34, d45, l34, l45);
            bool c46 = CompareDeltas(S, kB, d45, d56, l45, l56);
            float o0 = SmartBlur(AO1[i], AO1[i + 1], AO1[i + 2], AO1[i + 3], AO1[i + 4], c02, c13, c24);
            float o1 = SmartBlur(AO1[i + 1], AO1[i + 2], AO1[i + 3], AO1[i + 4], AO1[i + 5], c13, c24, c35);
            float o2 = SmartBlur(AO1[i + 2], AO1[i + 3], AO1[i + 4], AO1[i + 5], AO1[i + 6], c24, c35, c46);
            AO2[i] = o0; AO2[i + 1] = o1; AO2[i + 2] = o2;
        }

        static void BlurVertically(float S, float kB, float[] AO1, float[] DC, float[] AO2, int i)                // :132-170
        {
            float a0 = AO2[i], a1 = AO2[i + 16], a2 = AO2[i + 32], a3 = AO2[i + 48], a4 = AO2[i + 64], a5 = AO2[i + 80];
            float d01 = (float)(DC[i + 18] - DC[i + 2]), d12 = (float)(DC[i + 34] - DC[i + 18]), d23 = (float)(DC[i + 50] - DC[i + 34]);
            float d34 = (float)(DC[i + 66] - DC[i + 50]), d45 = (float)(DC[i + 82] - DC[i + 66]);
            float l01 = Mad(d01, d01, S), l12 = Mad(d12, d12, S), l23 = Mad(d23, d23, S), l34 = Mad(d34, d34, S), l45 = Mad(d45, d45, S);
            bool c02 = CompareDeltas(S, kB, d01, d12, l01, l12), c13 = CompareDeltas(S, kB, d12, d23, l12, l23);
            bool c24 = CompareDeltas(S, kB, d23, d34, l23, l34), c35 = CompareDeltas(S, kB, d34, d45, l34, l45);
            float r1 = SmartBlur(a0, a1, a2, a3, a4, c02, c13, c24), r2 = SmartBlur(a1, a2, a3, a4, a5, c13, c24, c35);
            AO1[i] = r1; AO1[i + 16] = r2;
        }

        static float BilateralUpsample(float tol, float nfs, float hiDepth, float hiAO,                         // :177-183
                                       float ld0, float ld1, float ld2, float ld3, float la0, float la1, float la2, float la3)
        {
            float w0 = 9.0f / (float)(Math.Abs((float)(hiDepth - ld0)) + tol), w1 = 3.0f / (float)(Math.Abs((float)(hiDepth - ld1)) + tol);
            float w2 = 1.0f / (float)(Math.Abs((float)(hiDepth - ld2)) + tol), w3 = 3.0f / (float)(Math.Abs((float)(hiDepth - ld3)) + tol);
            float totalWeight = (float)((float)((float)((float)(w0 + w1) + w2) + w3) + nfs);
            float weightedSum = (float)(Mad(la3, w3, Mad(la2, w2, Mad(la1, w1, (float)(la0 * w0)))) + nfs);
            return (float)((float)(hiAO * weightedSum) / totalWeight);
        }

        public void Upsample(int lo)                        // meao_oracle_upsample, wiring AmbientOcclusion.cs:528-531
        {
            int hi = lo - 1, low = lw[lo], loh = lh[lo], hiw = lw[hi], hih = lh[hi];
            float[] loDepth = lowDepth[lo], loAo = (lo == 4 || (singleScale && lo == 1)) ? occlusion[lo] : combined[lo];   // meao_oracle.c: single_scale
            float[] hiDepth = (hi == 0) ? linearDepth : lowDepth[hi], hiAo = (hi == 0) ? null : occlusion[hi];
            float[] dest = (hi == 0) ? result : combined[hi];
            float nfs, S, kB, tol;
            UpsampleConstants(lo, out nfs, out S, out kB, out tol);
            Action<int, int, float> store = (x, y, v) => { if (x >= 0 && y >= 0 && x < hiw && y < hih) dest[y * hiw + x] = StUnorm8(v); };
            int ngx = (hiw + 17) / 16;
            Striped((hih + 17) / 16, (gy0, gy1) =>          // dispatch ((hi.w+17)/16, (hi.h+17)/16, 1), :782-784
            {
                float[] DC = new float[256], AO1 = new float[256], AO2 = new float[256];          // Upsample.compute:50-52
                for (int gy = gy0; gy < gy1; gy++)
                for (int gx = 0; gx < ngx; gx++)
                {
                    Array.Clear(AO2, 0, 256);       // row 13 is read (:139) but never written; it feeds only an unconsumed output
                    for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++)
                    {
                        int index = (tx << 1) | (ty << 5), cx = gx * 8 + tx + tx - 2, cy = gy * 8 + ty + ty - 2;   // :191
                        float ax, ay, az, aw, dx, dy, dz, dw;
                        Gather4(loAo, 0, low, loh, cx, cy, out ax, out ay, out az, out aw);         // :56
                        AO1[index] = aw; AO1[index + 1] = az; AO1[index + 16] = ax; AO1[index + 17] = ay;          // :62-65
                        Gather4(loDepth, 0, low, loh, cx, cy, out dx, out dy, out dz, out dw);      // :67
                        DC[index] = 1.0f / dw; DC[index + 1] = 1.0f / dz; DC[index + 16] = 1.0f / dx; DC[index + 17] = 1.0f / dy;
                    }
                    for (int GI = 0; GI < 39; GI++) BlurHorizontally(S, kB, AO1, DC, AO2, (GI / 3) * 16 + (GI % 3) * 3);   // :199-200
                    for (int GI = 0; GI < 45; GI++) BlurVertically(S, kB, AO1, DC, AO2, (GI / 9) * 32 + GI % 9);          // :206-207
                    for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++)
                    {
                        int X = gx * 8 + tx, Y = gy * 8 + ty, idx0 = tx + ty * 16;                  // :213
                        float lsx = AO1[idx0 + 16], lsy = AO1[idx0 + 17], lsz = AO1[idx0 + 1], lsw = AO1[idx0];   // :214
                        float hx = 1f, hy = 1f, hz = 1f, hw = 1f;                                   // :223
                        if (hiAo != null) Gather4(hiAo, 0, hiw, hih, 2 * X, 2 * Y, out hx, out hy, out hz, out hw);   // :221
                        float lx, ly, lz, lwv, ex, ey, ez, ew;
                        Gather4(loDepth, 0, low, loh, X, Y, out lx, out ly, out lz, out lwv);       // :225
                        Gather4(hiDepth, 0, hiw, hih, 2 * X, 2 * Y, out ex, out ey, out ez, out ew); // :226
                        int ox = X << 1, oy = Y << 1;                                               // :228
                        store(ox - 1, oy, BilateralUpsample(tol, nfs, ex, hx, lx, ly, lz, lwv, lsx, lsy, lsz, lsw));       // :229
                        store(ox, oy, BilateralUpsample(tol, nfs, ey, hy, ly, lz, lwv, lx, lsy, lsz, lsw, lsx));           // :230
                        store(ox, oy - 1, BilateralUpsample(tol, nfs, ez, hz, lz, lwv, lx, ly, lsz, lsw, lsx, lsy));       // :231
                        store(ox - 1, oy - 1, BilateralUpsample(tol, nfs, ew, hw, lwv, lx, ly, lz, lsw, lsx, lsy, lsz));   // :232
                    }
                }
            });
        }

        public byte[] Run(float[] depth)                    // meao_oracle_run, record order AmbientOcclusion.cs:511-531
        {
            Downsample(depth);
            int kmax = singleScale ? 1 : 4;                 // single-scale: nothing coarser than level 1 contributes
            for (int k = 1; k <= kmax; k++) Render(k);
            for (int lo = kmax; lo >= 1; lo--) Upsample(lo);
            byte[] ao = new byte[W * H];
            for (int i = 0; i < ao.Length; i++) ao[i] = (byte)(uint)((float)(result[i] * 255.0f) + 0.5f);
            return ao;
        }

        // The timed CPU baseline of north_star: Mpixels/s of whole frames on `threads` host threads (median of `reps`).
        public static double TimeFrames(int width, int height, float[] depth, int threads, int reps)
        {
            var ao = new AoScalar(width, height) { threads = threads };
            ao.cam.tanHalfFovH = (float)((double)width / height * Math.Tan(Math.PI / 6));   // 60 degree vertical fov
            ao.p.intensity = 1.1f;                                                          // Sponza.unity:969
            ao.Run(depth);                                                                  // warm-up
            var times = new double[reps];
            for (int r = 0; r < reps; r++) { var sw = Stopwatch.StartNew(); ao.Run(depth); times[r] = sw.Elapsed.TotalSeconds; }
            Array.Sort(times);
            return (double)width * height / times[reps / 2] / 1e6;
        }
    }
}
