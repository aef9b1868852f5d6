/*
 * meao_oracle.h -- CPU oracle for the multi-scale SSAO hot path of keijiro/MiniEngineAO.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker / the timed CPU baseline.  The product path (libmeao.so, CUDA)
 * never links, imports or calls this code.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or captures for this path
 *  and its HLSL/C# cannot be executed in this image (no dotnet /
 * mono / dxc / Unity).  The oracle is therefore pinned only by (i) the analytic identities
 * that follow from the reference code (constant depth => AO 255, intensity 0 => 255,
 * sum of sample weights == 1, ...), (ii) the constant tables printed in SURVEY.md 8(a),
 * and (iii) an independent second restatement (oracle/direct_formulation.py).
 *
 * The oracle follows the reference thread-group structure literally (group ids, group
 * shared arrays, barriers), one thread at a time, so every line can be checked against
 * the HLSL.  Reference files (under /root/reference/Assets/MiniEngineAO/):
 *   Shaders/Downsample1.compute:37-81, Shaders/Downsample2.compute:32-51,
 *   Shaders/Render.compute:60-177 (variant main_interleaved: INTERLEAVE_RESULT, TILE_DIM 16; plus the
 *     undispatched variants: kernel main = WIDE_SAMPLING / TILE_DIM 32, and SAMPLE_EXHAUSTIVELY),
 *   Shaders/Upsample.compute:54-233 (variants main / main_blendout; plus main_premin / main_premin_blendout),
 *   AmbientOcclusion.cs:262-281 (formats, sizes), :561-593 (constants), :604-785 (dispatch).
 *
 * Conventions for the fixed-function behaviour the HLSL relies on (D3D11, not in the repo):
 *   - fp32 arithmetic, round-to-nearest-even; a*b+c written in ONE HLSL expression is
 *     contracted to a fused mad (explicit fmaf below; compile with -ffp-contract=off so
 *     nothing else fuses).  Build with -DMEAO_ORACLE_NO_FMA to get the unfused variant.
 *   - x / y and 1 / y are IEEE correctly rounded.  (D3D11 guarantees only 1 ULP: -DMEAO_ORACLE_DIV_MODE=1 builds
 *     x / y := x * (1 / y), =2 builds truncated quotients, to measure what that freedom can change -- DESIGN.md section 3.)
 *   - f32 -> f16 store: round-to-nearest-even, overflow -> +inf.  f16 -> f32 load: exact.
 *     (-DMEAO_ORACLE_F16_RTZ builds the truncating store the D3D spec also allowed.)
 *   - f32 -> UNORM8 store: NaN -> 0, clamp to [0,1], k = (uint)(x * 255.0f + 0.5f).
 *     UNORM8 -> f32 load: (float)k * (1.0f / 255.0f).
 *   - out-of-bounds texture Load -> 0; out-of-bounds UAV store -> dropped.
 *   - Gather(): point + clamp addressing; the footprint of texel-corner coordinate c is
 *     texels (c-1, c) per axis; .w=(x0,y0) .z=(x1,y0) .x=(x0,y1) .y=(x1,y1).
 *   - saturate(NaN) = 0; min/max return the non-NaN operand.
 *   - UNITY_REVERSED_Z defined (D3D11/12) unless reversed_z == 0.
 */
#ifndef MEAO_ORACLE_H
#define MEAO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* AmbientOcclusion.cs:20-58 -- the serialized parameter surface (defaults in comments). */
typedef struct {
    float noise_filter_tolerance;   /* [-8, 0]   default 0     AO.cs:20 */
    float blur_tolerance;           /* [-8,-1]   default -4.6  AO.cs:28 */
    float upsample_tolerance;       /* [-12,-1]  default -12   AO.cs:36 */
    float thickness_modifier;       /* [1, 10]   default 1     AO.cs:44 */
    float intensity;                /* [0, 2]    default 1     AO.cs:52 */
} MeaoOracleParams;

typedef struct {
    float near_clip, far_clip;      /* AO.cs:563 */
    float tan_half_fov_h;           /* AO.cs:570-573: 1 / projectionMatrix[0,0] */
    int   reversed_z;               /* AO.cs:564 / UNITY_REVERSED_Z (DS1:41) */
} MeaoOracleCamera;

/* All buffers are float arrays holding POST-quantisation values (an f16 buffer holds
 * floats exactly representable in f16, a UNORM8 buffer holds k * (1/255)).  With
 * quantize_storage == 0 the storage conversions become identities (the "storage=float32"
 * diagnostic switch of SURVEY.md 8c). */
typedef struct MeaoOracle {
    int W, H;
    int lw[7], lh[7];               /* AO.cs:276-281: ceil(base / 2^level) */
    int quantize_storage;
    int depth_is_linear;            /* 0: raw depth through Linearize (reference); 1: input already linear */
    MeaoOracleParams params;
    MeaoOracleCamera camera;
    /* ---- variants the reference ships in its shaders but never dispatches; all 0 = reference behaviour */
    int sample_exhaustively;        /* Render.compute:144-159 (#define SAMPLE_EXHAUSTIVELY, 68 taps); weight zeroing AO.cs:709-715 skipped */
    int single_pass_stereo;         /* AO.cs:680: ThicknessMultiplier *= 2 (the width doubling of AO.cs:339 is the caller's) */
    int high_quality_mask;          /* bit k-1: level k also runs Render.compute kernel "main" (WIDE_SAMPLING, non-tiled source
                                       LowDepth<k>, AO.cs:679) into HighQuality<k>, and the upsample whose LOW level is k runs
                                       main_premin / main_premin_blendout with LoResAO2 = HighQuality<k> (Upsample.compute:23,25,58-60) */
    int single_scale;               /* BASELINE.json configs[0]: meao_oracle_run = Downsample + Render level 1 + final-style Upsample with
                                       LoResAO1 = Occlusion1 (three of the ten dispatches of AO.cs:511-531) */
    float *linear_depth;            /* id 1      L0      f16   */
    float *low_depth[5];            /* id 2..5   L1..L4  f32   [1..4] */
    float *tiled_depth[5];          /* id 6..9   L3..L6 x16 slices  f16   [1..4] */
    float *occlusion[5];            /* id 10..13 L1..L4  unorm8 [1..4] */
    float *combined[4];             /* id 14..16 L1..L3  unorm8 [1..3] */
    float *result;                  /* id 17     L0      unorm8 */
    float *high_quality[5];         /* id 18..21 L1..L4  unorm8 [1..4]  (extension ids: not in AO.cs:787-808) */
} MeaoOracle;

MeaoOracle *meao_oracle_create(int width, int height);
void meao_oracle_destroy(MeaoOracle *o);

/* storage conversions (exposed for tests) */
float    meao_oracle_f16_round(float x);          /* f32 -> f16 (RTNE) -> f32 */
uint16_t meao_oracle_f32_to_f16_bits(float x);
float    meao_oracle_f16_bits_to_f32(uint16_t h);
uint8_t  meao_oracle_unorm8_code(float x);        /* store conversion, returns k */

/* CPU-side constants, AmbientOcclusion.cs:561-593, 660-734, 750-771 */
void meao_oracle_zbuffer_params(const MeaoOracleCamera *cam, float out4[4]);
void meao_oracle_sample_thickness(float out12[12]);
void meao_oracle_render_constants(const MeaoOracle *o, int level /*1..4*/,
                                  float inv_thickness[12], float sample_weight[12],
                                  float inv_slice_dim[2], float *reject_fadeoff, float *intensity);
/* same for Render.compute kernel "main" (WIDE_SAMPLING): source = LowDepth<level>, not tiled => AO.cs:679 doubles the thickness */
void meao_oracle_render_constants_wide(const MeaoOracle *o, int level /*1..4*/,
                                       float inv_thickness[12], float sample_weight[12],
                                       float inv_slice_dim[2], float *reject_fadeoff, float *intensity);
void meao_oracle_upsample_constants(const MeaoOracle *o, int lo_level /*1..4*/,
                                    float inv_low[2], float inv_high[2], float *noise_filter_strength,
                                    float *step_size, float *blur_tolerance, float *upsample_tolerance);

/* stages (record order AO.cs:511-531).  'threads' > 1 splits thread groups over std threads
 * row-striped; results are identical for any thread count. */
void meao_oracle_downsample(MeaoOracle *o, const float *depth, int threads);   /* DS1 + DS2 */
void meao_oracle_render(MeaoOracle *o, int level /*1..4*/, int threads);       /* REN main_interleaved */
void meao_oracle_render_wide(MeaoOracle *o, int level /*1..4*/, int threads);  /* REN main (WIDE_SAMPLING) -> HighQuality<level> */
void meao_oracle_upsample(MeaoOracle *o, int lo_level /*4..1*/, int threads);  /* UPS main(_blendout); main_premin(_blendout) when bit lo_level-1 of high_quality_mask is set */
void meao_oracle_run(MeaoOracle *o, const float *depth, int threads);          /* steps 1..10 (+ the high-quality renders selected by high_quality_mask) */

/* Debug views: PushDebugBlitCommands, AO.cs:787-820, followed by Blit.shader pass 3 (:116-134).
 * Writes the W x H R8 image the _result target holds after the debug blit of buffer <debug_id> (1..17):
 *   non-tiled source: cmd.Blit(rt, _result) = point-sampled stretch (RTs are FilterMode.Point, AO.cs:206,228,236),
 *                     texel = floor(uv * size) at the pixel centre uv = ((x+.5)/W, (y+.5)/H), evaluated in exact integers;
 *   tiled source:     Blit.shader pass 4 (:136-156): uv4 = uv*4, slice = floor(uv4.x) + 4*floor(uv4.y), point sample at frac(uv4);
 *   id 17: the AO result itself.  The R8 store applies the UNORM8 rule to the sampled value. */
void meao_oracle_debug_view(const MeaoOracle *o, int debug_id, uint8_t *out_codes);
/* Blit.shader pass 3 (:116-134) without blending, recorded at AO.cs:826-829: rgba = view.rrrr.  R8 -> RGBA8 keeps the code on
 * every channel; R8 -> RGBA16F stores f16(code * (1/255)), RTNE. */
void meao_oracle_composite_debug(const uint8_t *view_codes, void *rgba, int is_half, size_t npix);

/* Composite passes.  Fixed-function output-merger blending restated in fp32:
 *   pass 2, Blit.shader:84-101 + "Blend Zero SrcAlpha"                          : dst.rgba *= ao
 *   pass 1, Blit.shader:66-92 + "Blend Zero OneMinusSrcColor, Zero OneMinusSrcAlpha": gbuffer0.a *= 1-(1-ao), gbuffer3.rgb *= 1-(1-ao)
 * ao = code * (1/255); RGBA8 targets load code*(1/255) and store through the UNORM8 rule; RGBA16F load exact, store RTNE. */
void meao_oracle_composite_framebuffer(const uint8_t *ao_codes, void *rgba, int is_half, size_t npix);
void meao_oracle_composite_gbuffer(const uint8_t *ao_codes, uint8_t *gbuffer0_rgba8, void *gbuffer3_rgba, int g3_is_half, size_t npix);

/* debug ids 1..17 (AO.cs:787-808).  Returns pointer + dims (depth = 16 for tiled). */
const float *meao_oracle_get_buffer(const MeaoOracle *o, int debug_id, int *w, int *h, int *slices);
float *meao_oracle_buffer_mut(MeaoOracle *o, int debug_id, int *w, int *h, int *slices);

#ifdef __cplusplus
}
#endif
#endif
