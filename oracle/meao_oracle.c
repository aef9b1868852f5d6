/*
 * meao_oracle.c -- CPU restatement of the MiniEngineAO compute path.  See meao_oracle.h for
 * the conventions and the test-infrastructure-only notice.  PARITY UNPINNED (no reference
 * golden vectors exist; see header).
 *
 * Build: gcc -O2 -fno-tree-vectorize -ffp-contract=off -fPIC -shared (oracle/Makefile).
 * Each function cites the reference lines it follows; paths are relative to
 * /root/reference/Assets/MiniEngineAO/.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE          /* syscall(), for the worker pool's futex waits */
#endif
#include "meao_oracle.h"

#include <linux/futex.h>
#include <math.h>
#include <pthread.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * arithmetic conventions
 * ---------------------------------------------------------------------------------------- */
#ifdef MEAO_ORACLE_NO_FMA
static inline float mad(float a, float b, float c) { volatile float p = a * b; return p + c; }
#else
static inline float mad(float a, float b, float c) { return fmaf(a, b, c); }
#endif

/* Shader-side division and reciprocal (`x / y`, `1.0 / y` in the HLSL).  Default: IEEE correctly rounded -- the convention the
 * CUDA path is held to.  D3D11 only guarantees 1 ULP for both, so two LEGAL alternatives can be built to measure how much
 * that freedom can move the output (oracle/Makefile):
 *   MEAO_ORACLE_DIV_MODE 1: x / y := x * (1 / y), the reciprocal correctly rounded (what `mul(x, rcp(y))` hardware does: two roundings);
 *   MEAO_ORACLE_DIV_MODE 2: quotient and reciprocal truncated (round toward zero) instead of rounded to nearest. */
#ifndef MEAO_ORACLE_DIV_MODE
#define MEAO_ORACLE_DIV_MODE 0
#endif
static inline float rtz_from_double(double q)
{
    float f = (float)q;                                         /* nearest */
    if (f != f || f == 0.0f || isinf(f)) return f;
    if (fabs((double)f) > fabs(q)) f = nextafterf(f, 0.0f);     /* step back toward zero */
    return f;
}
static inline float grcp(float y)
{
#if MEAO_ORACLE_DIV_MODE == 2
    return rtz_from_double(1.0 / (double)y);                   /* the double quotient of two floats never ties a float boundary inexactly */
#else
    return 1.0f / y;
#endif
}
static inline float gdiv(float x, float y)
{
#if MEAO_ORACLE_DIV_MODE == 1
    volatile float r = 1.0f / y; return x * r;
#elif MEAO_ORACLE_DIV_MODE == 2
    return rtz_from_double((double)x / (double)y);
#else
    return x / y;
#endif
}

/* HLSL saturate: NaN -> 0 */
static inline float sat(float x) { return x > 0.0f ? (x < 1.0f ? x : 1.0f) : 0.0f; }
/* HLSL min/max: return the non-NaN operand */
static inline float hmax(float a, float b) { return (a != a) ? b : ((b != b) ? a : (a > b ? a : b)); }
static inline float hmin(float a, float b) { return (a != a) ? b : ((b != b) ? a : (a < b ? a : b)); }
static inline float hclamp(float x, float lo, float hi) { return hmin(hmax(x, lo), hi); }
static inline int iclamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* ------------------------------------------------------------------------------------------
 * storage conversions (formats: AmbientOcclusion.cs:262-273)
 * ---------------------------------------------------------------------------------------- */
/* Default f32 -> f16 store: round to nearest even, overflow -> inf (the convention the CUDA path is held to).
 * -DMEAO_ORACLE_F16_RTZ builds the other rounding the D3D spec historically allowed for float16 UAV / RT writes: truncation,
 * where finite values above 65504 store as 65504 (only a true inf stores inf). */
uint16_t meao_oracle_f32_to_f16_bits(float x)
{
    uint32_t u; memcpy(&u, &x, 4);
    uint32_t sign = (u >> 16) & 0x8000u;
    uint32_t absu = u & 0x7fffffffu;
    if (absu > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);           /* NaN */
#ifdef MEAO_ORACLE_F16_RTZ
    if (absu == 0x7f800000u) return (uint16_t)(sign | 0x7c00u);          /* inf */
    if (absu >= 0x477fe000u) return (uint16_t)(sign | 0x7bffu);          /* >= 65504: largest finite half */
    if (absu < 0x33800000u) return (uint16_t)sign;                       /* < 2^-24: below the smallest subnormal half */
    {
        int e = (int)(absu >> 23) - 127;
        uint32_t m = (absu & 0x7fffffu) | 0x800000u;
        if (e >= -14) return (uint16_t)(sign | (((uint32_t)(e + 15) << 10) + ((m & 0x7fffffu) >> 13)));
        int shift = 13 + (-14 - e);
        return (uint16_t)(sign | (shift > 24 ? 0u : (m >> shift)));
    }
#endif
    if (absu >= 0x47800000u) {                                           /* >= 65536 (incl. inf) */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (absu >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);          /* rounds up to 65536 -> inf (65520 is the tie, to even = inf) */
    if (absu < 0x33000000u) return (uint16_t)sign;                       /* < 2^-25 -> 0 (2^-25 itself is a tie -> even = 0, handled below) */
    int e = (int)(absu >> 23) - 127;
    uint32_t m = (absu & 0x7fffffu) | 0x800000u;                         /* 24-bit significand */
    int shift;                                                           /* bits dropped from m */
    uint32_t base;
    if (e >= -14) { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7fffffu; }
    else { shift = 13 + (-14 - e); base = 0; }                           /* subnormal half */
    if (shift > 24) return (uint16_t)sign;
    uint32_t q = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    return (uint16_t)(sign | (base + q));                                /* carry into exponent is correct by construction */
}

float meao_oracle_f16_bits_to_f32(uint16_t h)
{
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu, u;
    if (e == 0) {
        if (m == 0) u = sign;
        else { float f = (float)m * 5.9604644775390625e-8f; /* 2^-24 */ memcpy(&u, &f, 4); u |= sign; }
    } else if (e == 31) u = sign | 0x7f800000u | (m << 13);
    else u = sign | ((e + 112u) << 23) | (m << 13);
    float f; memcpy(&f, &u, 4); return f;
}

float meao_oracle_f16_round(float x) { return meao_oracle_f16_bits_to_f32(meao_oracle_f32_to_f16_bits(x)); }

uint8_t meao_oracle_unorm8_code(float x)
{
    float c = sat(x);                       /* NaN -> 0, clamp */
    return (uint8_t)(c * 255.0f + 0.5f);    /* truncation */
}

static inline float st_half(const MeaoOracle *o, float x) { return o->quantize_storage ? meao_oracle_f16_round(x) : x; }
static inline float st_unorm8(const MeaoOracle *o, float x)
{
    if (!o->quantize_storage) return sat(x);
    return (float)meao_oracle_unorm8_code(x) * (1.0f / 255.0f);
}

/* ------------------------------------------------------------------------------------------
 * geometry / allocation  (AmbientOcclusion.cs:124-131, 276-281, 453-475)
 * ---------------------------------------------------------------------------------------- */
MeaoOracle *meao_oracle_create(int width, int height)
{
    if (width <= 0 || height <= 0) return NULL;
    MeaoOracle *o = (MeaoOracle *)calloc(1, sizeof(*o));
    o->W = width; o->H = height;
    for (int l = 0; l < 7; l++) {
        int div = 1 << l;                                   /* AO.cs:278-280 */
        o->lw[l] = (width + (div - 1)) / div;
        o->lh[l] = (height + (div - 1)) / div;
    }
    o->quantize_storage = 1;
    o->params.noise_filter_tolerance = 0.0f;                /* AO.cs:20 */
    o->params.blur_tolerance = -4.6f;                       /* AO.cs:28 */
    o->params.upsample_tolerance = -12.0f;                  /* AO.cs:36 */
    o->params.thickness_modifier = 1.0f;                    /* AO.cs:44 */
    o->params.intensity = 1.0f;                             /* AO.cs:52 */
    o->camera.near_clip = 0.3f; o->camera.far_clip = 100.0f;
    o->camera.tan_half_fov_h = 1.0f; o->camera.reversed_z = 1;
#define ALLOC(n) ((float *)calloc((size_t)(n), sizeof(float)))
    o->linear_depth = ALLOC((size_t)o->lw[0] * o->lh[0]);
    for (int k = 1; k <= 4; k++) {
        o->low_depth[k] = ALLOC((size_t)o->lw[k] * o->lh[k]);
        o->tiled_depth[k] = ALLOC((size_t)16 * o->lw[k + 2] * o->lh[k + 2]);
        o->occlusion[k] = ALLOC((size_t)o->lw[k] * o->lh[k]);
        if (k <= 3) o->combined[k] = ALLOC((size_t)o->lw[k] * o->lh[k]);
        o->high_quality[k] = ALLOC((size_t)o->lw[k] * o->lh[k]);
    }
    o->result = ALLOC((size_t)o->lw[0] * o->lh[0]);
#undef ALLOC
    return o;
}

void meao_oracle_destroy(MeaoOracle *o)
{
    if (!o) return;
    free(o->linear_depth); free(o->result);
    for (int k = 1; k <= 4; k++) {
        free(o->low_depth[k]); free(o->tiled_depth[k]); free(o->occlusion[k]); free(o->high_quality[k]);
        if (k <= 3) free(o->combined[k]);
    }
    free(o);
}

float *meao_oracle_buffer_mut(MeaoOracle *o, int id, int *w, int *h, int *slices)
{
    int lvl, s = 1; float *p;
    if (id == 1) { lvl = 0; p = o->linear_depth; }
    else if (id >= 2 && id <= 5) { lvl = id - 1; p = o->low_depth[id - 1]; }
    else if (id >= 6 && id <= 9) { lvl = id - 5 + 2; p = o->tiled_depth[id - 5]; s = 16; }
    else if (id >= 10 && id <= 13) { lvl = id - 9; p = o->occlusion[id - 9]; }
    else if (id >= 14 && id <= 16) { lvl = id - 13; p = o->combined[id - 13]; }
    else if (id == 17) { lvl = 0; p = o->result; }
    else if (id >= 18 && id <= 21) { lvl = id - 17; p = o->high_quality[id - 17]; }   /* extension: HighQuality1..4 */
    else return NULL;
    if (w) *w = o->lw[lvl];
    if (h) *h = o->lh[lvl];
    if (slices) *slices = s;
    return p;
}

const float *meao_oracle_get_buffer(const MeaoOracle *o, int id, int *w, int *h, int *slices)
{
    return meao_oracle_buffer_mut((MeaoOracle *)o, id, w, h, slices);
}

/* ------------------------------------------------------------------------------------------
 * CPU-side constants
 * ---------------------------------------------------------------------------------------- */
/* AmbientOcclusion.cs:561-568 */
void meao_oracle_zbuffer_params(const MeaoOracleCamera *cam, float out4[4])
{
    float fpn = cam->far_clip / cam->near_clip;
    if (cam->reversed_z) { out4[0] = fpn - 1.0f; out4[1] = 1.0f; }
    else                 { out4[0] = 1.0f - fpn; out4[1] = fpn; }
    out4[2] = 0.0f; out4[3] = 0.0f;
}

/* Mathf.Sqrt(f) == (float)Math.Sqrt((double)f);  Mathf.Pow(f,p) == (float)Math.Pow(f,p) */
static float mathf_sqrt(float f) { return (float)sqrt((double)f); }
static float mathf_pow(float f, float p) { return (float)pow((double)f, (double)p); }

/* AmbientOcclusion.cs:577-590 */
void meao_oracle_sample_thickness(float t[12])
{
    t[0]  = mathf_sqrt(1 - 0.2f * 0.2f);
    t[1]  = mathf_sqrt(1 - 0.4f * 0.4f);
    t[2]  = mathf_sqrt(1 - 0.6f * 0.6f);
    t[3]  = mathf_sqrt(1 - 0.8f * 0.8f);
    t[4]  = mathf_sqrt(1 - 0.2f * 0.2f - 0.2f * 0.2f);
    t[5]  = mathf_sqrt(1 - 0.2f * 0.2f - 0.4f * 0.4f);
    t[6]  = mathf_sqrt(1 - 0.2f * 0.2f - 0.6f * 0.6f);
    t[7]  = mathf_sqrt(1 - 0.2f * 0.2f - 0.8f * 0.8f);
    t[8]  = mathf_sqrt(1 - 0.4f * 0.4f - 0.4f * 0.4f);
    t[9]  = mathf_sqrt(1 - 0.4f * 0.4f - 0.6f * 0.6f);
    t[10] = mathf_sqrt(1 - 0.4f * 0.4f - 0.8f * 0.8f);
    t[11] = mathf_sqrt(1 - 0.6f * 0.6f - 0.6f * 0.6f);
}

/* AmbientOcclusion.cs:660-734 for a source of src_w x src_h texels; tiled: source.isTiled (AO.cs:679) */
static void render_constants_impl(const MeaoOracle *o, int src_w, int src_h, int tiled,
                                  float inv_thickness[12], float sample_weight[12],
                                  float inv_slice_dim[2], float *reject_fadeoff, float *intensity)
{
    float thick[12];
    meao_oracle_sample_thickness(thick);
    const float ScreenspaceDiameter = 10;                                   /* AO.cs:669 */
    float ThicknessMultiplier = 2 * o->camera.tan_half_fov_h * ScreenspaceDiameter / (float)src_w; /* AO.cs:678 */
    if (!tiled) ThicknessMultiplier *= 2;                                   /* AO.cs:679 */
    if (o->single_pass_stereo) ThicknessMultiplier *= 2;                    /* AO.cs:680 */
    float InverseRangeFactor = 1 / ThicknessMultiplier;                     /* AO.cs:683 */
    for (int i = 0; i < 12; i++) inv_thickness[i] = InverseRangeFactor / thick[i];   /* AO.cs:687-688 */
    static const float mult[12] = {4, 4, 4, 4, 4, 8, 8, 8, 4, 8, 8, 4};     /* AO.cs:696-707 */
    for (int i = 0; i < 12; i++) sample_weight[i] = mult[i] * thick[i];
    if (!o->sample_exhaustively) {                                          /* AO.cs:709-715 ("FIXME: should we support SAMPLE_EXHAUSTIVELY mode?") */
        sample_weight[0] = 0; sample_weight[2] = 0; sample_weight[5] = 0;
        sample_weight[7] = 0; sample_weight[9] = 0;
    }
    float total = 0.0f;                                                     /* AO.cs:718-724 */
    for (int i = 0; i < 12; i++) total += sample_weight[i];
    for (int i = 0; i < 12; i++) sample_weight[i] /= total;
    inv_slice_dim[0] = 1.0f / (float)src_w;                                 /* AO.cs:169-172, 732 */
    inv_slice_dim[1] = 1.0f / (float)src_h;
    *reject_fadeoff = -1 / o->params.thickness_modifier;                    /* AO.cs:733 */
    *intensity = o->params.intensity;                                       /* AO.cs:734 */
}

/* source = TiledDepth<level>, i.e. mip level+2, tiled (the only calls AmbientOcclusion.cs makes, AO.cs:519-522) */
void meao_oracle_render_constants(const MeaoOracle *o, int level,
                                  float inv_thickness[12], float sample_weight[12],
                                  float inv_slice_dim[2], float *reject_fadeoff, float *intensity)
{
    render_constants_impl(o, o->lw[level + 2], o->lh[level + 2], 1, inv_thickness, sample_weight, inv_slice_dim, reject_fadeoff, intensity);
}

/* source = LowDepth<level>, not tiled (kernel "main"; PushRenderCommands would take the AO.cs:679 branch) */
void meao_oracle_render_constants_wide(const MeaoOracle *o, int level,
                                       float inv_thickness[12], float sample_weight[12],
                                       float inv_slice_dim[2], float *reject_fadeoff, float *intensity)
{
    render_constants_impl(o, o->lw[level], o->lh[level], 0, inv_thickness, sample_weight, inv_slice_dim, reject_fadeoff, intensity);
}

/* AmbientOcclusion.cs:757-771 */
void meao_oracle_upsample_constants(const MeaoOracle *o, int lo_level,
                                    float inv_low[2], float inv_high[2], float *noise_filter_strength,
                                    float *step_size, float *blur_tolerance, float *upsample_tolerance)
{
    int lo_w = o->lw[lo_level], lo_h = o->lh[lo_level];
    int hi_w = o->lw[lo_level - 1], hi_h = o->lh[lo_level - 1];
    float stepSize = 1920.0f / (float)lo_w;                                             /* AO.cs:760 */
    float blurTolerance = 1 - mathf_pow(10, o->params.blur_tolerance) * stepSize;       /* AO.cs:761 */
    blurTolerance *= blurTolerance;                                                     /* AO.cs:762 */
    float upsampleTolerance = mathf_pow(10, o->params.upsample_tolerance);              /* AO.cs:763 */
    float noiseFilterWeight = 1 / (mathf_pow(10, o->params.noise_filter_tolerance) + upsampleTolerance); /* AO.cs:764 */
    inv_low[0] = 1.0f / (float)lo_w; inv_low[1] = 1.0f / (float)lo_h;
    inv_high[0] = 1.0f / (float)hi_w; inv_high[1] = 1.0f / (float)hi_h;
    *noise_filter_strength = noiseFilterWeight;
    *step_size = stepSize;
    *blur_tolerance = blurTolerance;
    *upsample_tolerance = upsampleTolerance;
}

/* ------------------------------------------------------------------------------------------
 * row-striped thread helper: run fn(ctx, gy0, gy1) over [0, ny) thread-group rows
 * ---------------------------------------------------------------------------------------- */
typedef void (*stripe_fn)(void *ctx, int gy0, int gy1);

/* A persistent worker pool (the first version created and joined `threads` pthreads for every one of the ten stages of a frame).
 * Workers sleep on a futex word (no mutex to re-acquire: all of them wake in parallel), take chunks of thread-group rows through an
 * atomic cursor, and the last one to finish wakes the caller.  Rows are independent, so the result does not depend on the thread
 * count or on who runs which chunk. */
#define POOL_MAX 512
static struct {
    pthread_mutex_t region;                     /* one parallel region at a time */
    pthread_t tid[POOL_MAX];
    int nworkers;                               /* threads created so far */
    stripe_fn fn; void *ctx; int ny, chunk, active, xparts;
    char pad0[64];
    volatile int generation;                    /* futex word: bumped once per region */
    char pad1[60];
    volatile int running;                       /* futex word: workers that have not finished the current region */
    char pad2[60];
    volatile int cursor;
    char pad3[60];
} g_pool = { .region = PTHREAD_MUTEX_INITIALIZER };

static long pool_futex(volatile int *addr, int op, int val) { return syscall(SYS_futex, addr, op, val, NULL, NULL, 0); }

/* When a stage has fewer thread-group rows than ~4 per thread (the two big stages of a 4K frame have 135 rows: on 128 threads a
 * pure row split leaves half the cores idle), every row is additionally cut into `xparts` column ranges; the stripe functions read
 * their range through XLO / XHI. */
static __thread int t_xpart = 0, t_xparts = 1;
#define XLO(n) ((int)((long long)(n) * t_xpart / t_xparts))
#define XHI(n) ((int)((long long)(n) * (t_xpart + 1) / t_xparts))

/* MEAO_ORACLE_POOL: 2 (default) = units handed out through an atomic cursor; 1 = every worker owns a FIXED contiguous range of units
 * (a thread touches the same rows frame after frame); 0 = round 1: one pthread per row stripe, created and joined for every stage.
 * On the two-socket 128-thread B200 host (shared with other tenants: medians are noisy), one 4K frame, best / median of 8 runs on 64
 * threads: mode 0 166 / 126, mode 1 201 / 102, mode 2 245 / 102 Mpx/s (gpurun_out r2g, before the spin-then-sleep wait was added). */
#ifndef MEAO_ORACLE_POOL
#define MEAO_ORACLE_POOL 2
#endif

__attribute__((unused)) static void pool_run_unit_range(int u0, int u1)
{
    const int xparts = g_pool.xparts;
    if (xparts == 1) { if (u1 > u0) g_pool.fn(g_pool.ctx, u0, u1); return; }
    for (int u = u0; u < u1; u++) {
        t_xpart = u % xparts; t_xparts = xparts;
        g_pool.fn(g_pool.ctx, u / xparts, u / xparts + 1);
    }
    t_xpart = 0; t_xparts = 1;
}

static void pool_drain(int worker)
{
    const int xparts = g_pool.xparts, units = g_pool.ny * xparts;
#if MEAO_ORACLE_POOL == 1
    const int T = g_pool.active + 1;                /* workers 0 .. active-1 plus the caller (= worker `active`) */
    pool_run_unit_range((int)((long long)units * worker / T), (int)((long long)units * (worker + 1) / T));
    return;
#endif
    (void)worker;
    for (;;) {
        int u0 = __atomic_fetch_add(&g_pool.cursor, g_pool.chunk, __ATOMIC_RELAXED);
        if (u0 >= units) break;
        if (xparts == 1) {
            int u1 = u0 + g_pool.chunk; if (u1 > units) u1 = units;
            g_pool.fn(g_pool.ctx, u0, u1);
        } else {                                    /* chunk == 1: one (row, column range) unit at a time */
            t_xpart = u0 % xparts; t_xparts = xparts;
            g_pool.fn(g_pool.ctx, u0 / xparts, u0 / xparts + 1);
        }
    }
    t_xpart = 0; t_xparts = 1;
}

__attribute__((unused)) static void *pool_worker(void *arg)
{
    const int index = (int)(intptr_t)arg;
    int seen = 0;
    for (;;) {
        int g;
        /* the ten stages of a frame follow each other within microseconds: spin briefly before going to sleep */
        for (int spin = 0; (g = __atomic_load_n(&g_pool.generation, __ATOMIC_ACQUIRE)) == seen; spin++) {
            if (spin < 4000) __builtin_ia32_pause();
            else pool_futex(&g_pool.generation, FUTEX_WAIT_PRIVATE, seen);
        }
        seen = g;
        if (index >= g_pool.active) continue;
        pool_drain(index);
        if (__atomic_sub_fetch(&g_pool.running, 1, __ATOMIC_ACQ_REL) == 0) pool_futex(&g_pool.running, FUTEX_WAKE_PRIVATE, 1);
    }
    return NULL;
}

#if MEAO_ORACLE_POOL == 0
typedef struct { stripe_fn fn; void *ctx; int gy0, gy1; } stripe_job;
static void *stripe_main(void *p) { stripe_job *j = (stripe_job *)p; j->fn(j->ctx, j->gy0, j->gy1); return NULL; }
static void run_striped(stripe_fn fn, void *ctx, int ny, int threads)       /* round 1: one pthread per stripe, created and joined per stage */
{
    if (threads < 1) threads = 1;
    if (threads > ny) threads = ny > 0 ? ny : 1;
    if (threads == 1) { fn(ctx, 0, ny); return; }
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    stripe_job *jobs = (stripe_job *)malloc(sizeof(stripe_job) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].gy0 = (int)((long long)ny * t / threads);
        jobs[t].gy1 = (int)((long long)ny * (t + 1) / threads);
        pthread_create(&tid[t], NULL, stripe_main, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    free(tid); free(jobs);
}
#else
static void run_striped(stripe_fn fn, void *ctx, int ny, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > ny * 16) threads = ny > 0 ? ny * 16 : 1;
    if (threads > POOL_MAX) threads = POOL_MAX;
    if (threads == 1 || ny <= 0) { if (ny > 0) fn(ctx, 0, ny); return; }
    pthread_mutex_lock(&g_pool.region);
    while (g_pool.nworkers < threads - 1) {                 /* the calling thread is worker number `threads` */
        pthread_attr_t at; pthread_attr_init(&at); pthread_attr_setdetachstate(&at, PTHREAD_CREATE_DETACHED);
        if (pthread_create(&g_pool.tid[g_pool.nworkers], &at, pool_worker, (void *)(intptr_t)g_pool.nworkers) != 0) { pthread_attr_destroy(&at); break; }
        pthread_attr_destroy(&at);
        g_pool.nworkers++;
    }
    g_pool.fn = fn; g_pool.ctx = ctx; g_pool.ny = ny;
    g_pool.chunk = ny / (threads * 4); if (g_pool.chunk < 1) g_pool.chunk = 1;
    g_pool.xparts = (ny >= threads * 4) ? 1 : (threads * 4 + ny - 1) / ny;
    if (g_pool.xparts > 16) g_pool.xparts = 16;
    if (g_pool.xparts > 1) g_pool.chunk = 1;
    g_pool.cursor = 0;
    g_pool.active = g_pool.nworkers < threads - 1 ? g_pool.nworkers : threads - 1;
    __atomic_store_n(&g_pool.running, g_pool.active, __ATOMIC_RELEASE);
    __atomic_add_fetch(&g_pool.generation, 1, __ATOMIC_ACQ_REL);
    pool_futex(&g_pool.generation, FUTEX_WAKE_PRIVATE, POOL_MAX);
    pool_drain(g_pool.active);
    for (int spin = 0;; spin++) {
        int r = __atomic_load_n(&g_pool.running, __ATOMIC_ACQUIRE);
        if (r == 0) break;
        if (spin < 4000) __builtin_ia32_pause();
        else pool_futex(&g_pool.running, FUTEX_WAIT_PRIVATE, r);
    }
    pthread_mutex_unlock(&g_pool.region);
}
#endif

/* ------------------------------------------------------------------------------------------
 * Downsample1.compute
 * ---------------------------------------------------------------------------------------- */
typedef struct { MeaoOracle *o; const float *depth; float zb[4]; } ds_ctx;

/* Downsample1.compute:37-48 */
static float ds1_linearize(const ds_ctx *c, int x, int y)
{
    const MeaoOracle *o = c->o;
    int inb = (x < o->W && y < o->H);
    float depth = inb ? c->depth[(size_t)y * o->W + x] : 0.0f;           /* DS1:39, OOB load -> 0 */
    float dist;
    if (o->depth_is_linear) {
        dist = depth;                                                    /* not in the reference: linear-depth ingest */
    } else {
        dist = grcp(mad(c->zb[0], depth, c->zb[1]));                     /* DS1:40 */
        if (o->camera.reversed_z) { if (depth == 0) dist = 1e5f; }       /* DS1:41-42 */
        else                      { if (depth == 1) dist = 1e5f; }       /* DS1:43-44 */
    }
    if (inb) c->o->linear_depth[(size_t)y * o->W + x] = st_half(o, dist); /* DS1:46, OOB store dropped */
    return dist;
}

static inline unsigned slice_of(unsigned sx, unsigned sy) { return ((sx & 3u) | (sy << 2)) & 15u; }  /* DS1:69 */

/* Downsample1.compute:52-81; dispatch (tiled2.w, tiled2.h, 1) AO.cs:643 */
static void ds1_stripe(void *vc, int gy0, int gy1)
{
    ds_ctx *c = (ds_ctx *)vc; MeaoOracle *o = c->o;
    const int L1w = o->lw[1], L1h = o->lh[1], L2w = o->lw[2], L2h = o->lh[2];
    const int A1w = o->lw[3], A1h = o->lh[3], A2w = o->lw[4], A2h = o->lh[4];
    float cache[256];                                                    /* DS1:50 g_CacheW */
    for (int gy = gy0; gy < gy1; gy++)
    for (int gx = XLO(o->lw[4]); gx < XHI(o->lw[4]); gx++) {
        for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++) {
            int sx = (gx << 4) | tx, sy = (gy << 4) | ty;                /* DS1:55 */
            int dest = (ty << 4) | tx;                                   /* DS1:56 */
            cache[dest +   0] = ds1_linearize(c, sx | 0, sy | 0);        /* DS1:57-60 */
            cache[dest +   8] = ds1_linearize(c, sx | 8, sy | 0);
            cache[dest + 128] = ds1_linearize(c, sx | 0, sy | 8);
            cache[dest + 136] = ds1_linearize(c, sx | 8, sy | 8);
        }
        /* GroupMemoryBarrierWithGroupSync DS1:62 */
        for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++) {
            int GI = ty * 8 + tx;
            int lds = (tx << 1) | (ty << 5);                             /* DS1:64 */
            float w1 = cache[lds];                                       /* DS1:66 */
            int stx = gx * 8 + tx, sty = gy * 8 + ty;                    /* DS1:68 DTid.xy */
            unsigned slice = slice_of((unsigned)stx, (unsigned)sty);
            if (stx < L1w && sty < L1h) o->low_depth[1][(size_t)sty * L1w + stx] = w1;                   /* DS1:70 */
            if ((stx >> 2) < A1w && (sty >> 2) < A1h)                                                     /* DS1:71 */
                o->tiled_depth[1][((size_t)slice * A1h + (sty >> 2)) * A1w + (stx >> 2)] = st_half(o, w1);
            if ((GI & 011) == 0) {                                       /* DS1:73 (octal) */
                int s2x = stx >> 1, s2y = sty >> 1;                      /* DS1:75 */
                slice = slice_of((unsigned)s2x, (unsigned)s2y);
                if (s2x < L2w && s2y < L2h) o->low_depth[2][(size_t)s2y * L2w + s2x] = w1;               /* DS1:77 */
                if ((s2x >> 2) < A2w && (s2y >> 2) < A2h)                                                 /* DS1:78 */
                    o->tiled_depth[2][((size_t)slice * A2h + (s2y >> 2)) * A2w + (s2x >> 2)] = st_half(o, w1);
            }
        }
    }
}

/* Downsample2.compute:32-51; dispatch (tiled4.w, tiled4.h, 1) AO.cs:657 */
static void ds2_stripe(void *vc, int gy0, int gy1)
{
    ds_ctx *c = (ds_ctx *)vc; MeaoOracle *o = c->o;
    const int L2w = o->lw[2], L2h = o->lh[2], L3w = o->lw[3], L3h = o->lh[3], L4w = o->lw[4], L4h = o->lh[4];
    const int A3w = o->lw[5], A3h = o->lh[5], A4w = o->lw[6], A4h = o->lh[6];
    for (int gy = gy0; gy < gy1; gy++)
    for (int gx = XLO(o->lw[6]); gx < XHI(o->lw[6]); gx++)
    for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++) {
        int GI = ty * 8 + tx;
        int stx = gx * 8 + tx, sty = gy * 8 + ty;
        int rx = stx << 1, ry = sty << 1;
        float m1 = (rx < L2w && ry < L2h) ? o->low_depth[2][(size_t)ry * L2w + rx] : 0.0f;              /* DS2:35 */
        unsigned slice = slice_of((unsigned)stx, (unsigned)sty);                                         /* DS2:37-39 */
        if (stx < L3w && sty < L3h) o->low_depth[3][(size_t)sty * L3w + stx] = m1;                       /* DS2:40 */
        if ((stx >> 2) < A3w && (sty >> 2) < A3h)                                                        /* DS2:41 */
            o->tiled_depth[3][((size_t)slice * A3h + (sty >> 2)) * A3w + (stx >> 2)] = st_half(o, m1);
        if ((GI & 011) == 0) {                                                                           /* DS2:43 */
            int s2x = stx >> 1, s2y = sty >> 1;
            slice = slice_of((unsigned)s2x, (unsigned)s2y);
            if (s2x < L4w && s2y < L4h) o->low_depth[4][(size_t)s2y * L4w + s2x] = m1;                   /* DS2:48 */
            if ((s2x >> 2) < A4w && (s2y >> 2) < A4h)                                                    /* DS2:49 */
                o->tiled_depth[4][((size_t)slice * A4h + (s2y >> 2)) * A4w + (s2x >> 2)] = st_half(o, m1);
        }
    }
}

/* AmbientOcclusion.cs:604-658 (minus the raster depth copy, which the caller supplies) */
void meao_oracle_downsample(MeaoOracle *o, const float *depth, int threads)
{
    ds_ctx c; c.o = o; c.depth = depth;
    meao_oracle_zbuffer_params(&o->camera, c.zb);
    run_striped(ds1_stripe, &c, o->lh[4], threads);
    run_striped(ds2_stripe, &c, o->lh[6], threads);
}

/* ------------------------------------------------------------------------------------------
 * Gather (point, clamp): footprint of texel-corner coordinate (cx,cy) is texels (c-1, c)
 * ---------------------------------------------------------------------------------------- */
typedef struct { float x, y, z, w; } float4_t;
static inline float4_t gather4(const float *buf, int w, int h, int cx, int cy)
{
    int x0 = iclamp(cx - 1, 0, w - 1), x1 = iclamp(cx, 0, w - 1);
    int y0 = iclamp(cy - 1, 0, h - 1), y1 = iclamp(cy, 0, h - 1);
    float4_t r;
    r.w = buf[(size_t)y0 * w + x0];
    r.z = buf[(size_t)y0 * w + x1];
    r.x = buf[(size_t)y1 * w + x0];
    r.y = buf[(size_t)y1 * w + x1];
    return r;
}

/* ------------------------------------------------------------------------------------------
 * Render.compute: kernel main_interleaved (INTERLEAVE_RESULT => TILE_DIM 16, 8x8 threads, 16-slice
 * Texture2DArray source) and kernel main (WIDE_SAMPLING => TILE_DIM 32, 16x16 threads, plain
 * Texture2D source).  SAMPLE_EXHAUSTIVELY (REN:144-159) is a compile-time switch of both.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    MeaoOracle *o; int level;
    float inv_thickness[12], sample_weight[12], inv_slice_dim[2], reject_fadeoff, intensity;
} ren_ctx;

#define REN_TILE_DIM 16        /* REN:53 */
#define REN_TILE_DIM_WIDE 32   /* REN:48 */

/* Render.compute:60-75 */
static inline float ren_test_sample_pair(const float *DS, float rf, float frontDepth, float invRange, unsigned base, int offset)
{
    float disocclusion1 = mad(DS[(int)base + offset], invRange, -frontDepth);    /* REN:65 */
    float disocclusion2 = mad(DS[(int)base - offset], invRange, -frontDepth);    /* REN:66 */
    float pseudo1 = sat(rf * disocclusion1);                                      /* REN:68 */
    float pseudo2 = sat(rf * disocclusion2);                                      /* REN:69 */
    float s = hclamp(disocclusion1, pseudo2, 1.0f) + hclamp(disocclusion2, pseudo1, 1.0f);
    return sat(mad(-pseudo1, pseudo2, s));                                        /* REN:71-74 */
}

/* Render.compute:77-110; T = TILE_DIM, wide = WIDE_SAMPLING */
static inline float ren_test_samples(const float *DS, float rf, unsigned T, int wide, unsigned centerIdx, unsigned x, unsigned y, float invDepth, float invThickness)
{
    if (wide) { x <<= 1; y <<= 1; }                                               /* REN:79-82 */
    float invRange = invThickness * invDepth;                                     /* REN:84 */
    float frontDepth = invThickness - 0.5f;                                       /* REN:85 */
    if (y == 0) {                                                                 /* REN:87-93 axial */
        return 0.5f * (ren_test_sample_pair(DS, rf, frontDepth, invRange, centerIdx, (int)x) +
                       ren_test_sample_pair(DS, rf, frontDepth, invRange, centerIdx, (int)(x * T)));
    } else if (x == y) {                                                          /* REN:94-100 diagonal */
        return 0.5f * (ren_test_sample_pair(DS, rf, frontDepth, invRange, centerIdx, (int)(x * T - x)) +
                       ren_test_sample_pair(DS, rf, frontDepth, invRange, centerIdx, (int)(x * T + x)));
    } else {                                                                      /* REN:101-109 L-shaped */
        return 0.25f * (ren_test_sample_pair(DS, rf, frontDepth, invRange, centerIdx, (int)(y * T + x)) +
                        ren_test_sample_pair(DS, rf, frontDepth, invRange, centerIdx, (int)(y * T - x)) +
                        ren_test_sample_pair(DS, rf, frontDepth, invRange, centerIdx, (int)(x * T + y)) +
                        ren_test_sample_pair(DS, rf, frontDepth, invRange, centerIdx, (int)(x * T - y)));
    }
}

/* Render.compute:142-169: the weighted sum over the sample set.  Table slots: float4[3] arrays indexed [i/4][i%4] (REN:39-40). */
static inline float ren_accumulate(const ren_ctx *c, const float *DS, unsigned T, int wide, unsigned thisIdx, float invThisDepth)
{
    const float *iT = c->inv_thickness, *sW = c->sample_weight;
    const float rf = c->reject_fadeoff;
    float ao = 0.0f;                                                              /* REN:142 */
#define TS(slot, x, y) ao = mad(sW[slot], ren_test_samples(DS, rf, T, wide, thisIdx, x, y, invThisDepth, iT[slot]), ao)
    if (c->o->sample_exhaustively) {
        /* REN:146-159, 68 samples: all cells within a circular radius of 5 */
        TS(0, 1, 0); TS(1, 2, 0); TS(2, 3, 0); TS(3, 4, 0);                       /* REN:148-151  [0].xyzw */
        TS(4, 1, 1); TS(8, 2, 2); TS(11, 3, 3);                                   /* REN:152-154  [1].x [2].x [2].w */
        TS(5, 1, 2); TS(6, 1, 3); TS(7, 1, 4);                                    /* REN:155-157  [1].yzw */
        TS(9, 2, 3); TS(10, 2, 4);                                                /* REN:158-159  [2].yz */
    } else {
        /* REN:162-168, 36-sample checker pattern */
        TS(1, 2, 0); TS(3, 4, 0); TS(4, 1, 1); TS(8, 2, 2); TS(11, 3, 3); TS(6, 1, 3); TS(10, 2, 4);
    }
#undef TS
    return ao;
}

/* Render.compute:112-177 as main_interleaved; dispatch ceil(w/8) x ceil(h/8) x 16, AO.cs:739-747.
 * Stripes run over (slice z, group row gy) pairs flattened as z * ngy + gy. */
static void ren_stripe(void *vc, int r0, int r1)
{
    ren_ctx *c = (ren_ctx *)vc; MeaoOracle *o = c->o;
    const int k = c->level;
    const int sw = o->lw[k + 2], sh = o->lh[k + 2];
    const int ow = o->lw[k], oh = o->lh[k];
    const int ngx = (sw + 7) / 8, ngy = (sh + 7) / 8;
    float DS[REN_TILE_DIM * REN_TILE_DIM];                        /* REN:58 */
    for (int r = r0; r < r1; r++) {
        int z = r / ngy, gy = r % ngy;
        const float *slice = o->tiled_depth[k] + (size_t)z * sw * sh;
        for (int gx = XLO(ngx); gx < XHI(ngx); gx++) {
            for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++) {
                int cx = gx * 8 + tx + tx - 3, cy = gy * 8 + ty + ty - 3;            /* REN:118 (DTid + GTid - 3) * invDim */
                float4_t d = gather4(slice, sw, sh, cx, cy);                         /* REN:123 */
                int dest = tx * 2 + ty * 2 * REN_TILE_DIM;                           /* REN:127 */
                DS[dest] = d.w; DS[dest + 1] = d.z;                                  /* REN:128-129 */
                DS[dest + REN_TILE_DIM] = d.x; DS[dest + REN_TILE_DIM + 1] = d.y;    /* REN:130-131 */
            }
            /* GroupMemoryBarrierWithGroupSync REN:133 */
            for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++) {
                unsigned thisIdx = (unsigned)(tx + ty * REN_TILE_DIM + 4 * REN_TILE_DIM + 4);   /* REN:138 */
                const float invThisDepth = grcp(DS[thisIdx]);                                    /* REN:140 */
                float ao = ren_accumulate(c, DS, REN_TILE_DIM, 0, thisIdx, invThisDepth);        /* REN:142-169 */
                int ox = ((gx * 8 + tx) << 2) | (z & 3), oy = ((gy * 8 + ty) << 2) | (z >> 2); /* REN:172 */
                if (ox < ow && oy < oh)
                    o->occlusion[k][(size_t)oy * ow + ox] = st_unorm8(o, mad(c->intensity, ao - 1.0f, 1.0f)); /* REN:176 lerp(1, ao, I) */
            }
        }
    }
}

/* AmbientOcclusion.cs:660-748 */
void meao_oracle_render(MeaoOracle *o, int level, int threads)
{
    ren_ctx c; c.o = o; c.level = level;
    meao_oracle_render_constants(o, level, c.inv_thickness, c.sample_weight, c.inv_slice_dim, &c.reject_fadeoff, &c.intensity);
    int ngy = (o->lh[level + 2] + 7) / 8;
    run_striped(ren_stripe, &c, 16 * ngy, threads);
}

/* Render.compute:112-177 as kernel "main": WIDE_SAMPLING, TILE_DIM 32, 16x16 threads (REN:46-50), source =
 * Texture2D<float> LowDepth<level> (f32, point + clamp Gather REN:125), output at DTid.xy (REN:174).
 * Dispatch by the generic formula of AO.cs:742-747 with the kernel's own group size: ceil(w/16) x ceil(h/16) x 1. */
static void ren_wide_stripe(void *vc, int gy0, int gy1)
{
    ren_ctx *c = (ren_ctx *)vc; MeaoOracle *o = c->o;
    const int k = c->level;
    const int sw = o->lw[k], sh = o->lh[k];
    const int ngx = (sw + 15) / 16;
    const float *src = o->low_depth[k];
    float DS[REN_TILE_DIM_WIDE * REN_TILE_DIM_WIDE];                                 /* REN:58 */
    for (int gy = gy0; gy < gy1; gy++)
    for (int gx = XLO(ngx); gx < XHI(ngx); gx++) {
        for (int ty = 0; ty < 16; ty++) for (int tx = 0; tx < 16; tx++) {
            int cx = gx * 16 + tx + tx - 7, cy = gy * 16 + ty + ty - 7;              /* REN:116 (DTid + GTid - 7) * invDim */
            float4_t d = gather4(src, sw, sh, cx, cy);                               /* REN:125 */
            int dest = tx * 2 + ty * 2 * REN_TILE_DIM_WIDE;                          /* REN:127 */
            DS[dest] = d.w; DS[dest + 1] = d.z;                                      /* REN:128-129 */
            DS[dest + REN_TILE_DIM_WIDE] = d.x; DS[dest + REN_TILE_DIM_WIDE + 1] = d.y;   /* REN:130-131 */
        }
        /* GroupMemoryBarrierWithGroupSync REN:133 */
        for (int ty = 0; ty < 16; ty++) for (int tx = 0; tx < 16; tx++) {
            unsigned thisIdx = (unsigned)(tx + ty * REN_TILE_DIM_WIDE + 8 * REN_TILE_DIM_WIDE + 8);   /* REN:136 */
            const float invThisDepth = grcp(DS[thisIdx]);                                              /* REN:140 */
            float ao = ren_accumulate(c, DS, REN_TILE_DIM_WIDE, 1, thisIdx, invThisDepth);             /* REN:142-169 */
            int ox = gx * 16 + tx, oy = gy * 16 + ty;                                                  /* REN:174 */
            if (ox < sw && oy < sh)
                o->high_quality[k][(size_t)oy * sw + ox] = st_unorm8(o, mad(c->intensity, ao - 1.0f, 1.0f));  /* REN:176 */
        }
    }
}

void meao_oracle_render_wide(MeaoOracle *o, int level, int threads)
{
    ren_ctx c; c.o = o; c.level = level;
    meao_oracle_render_constants_wide(o, level, c.inv_thickness, c.sample_weight, c.inv_slice_dim, &c.reject_fadeoff, &c.intensity);
    run_striped(ren_wide_stripe, &c, (o->lh[level] + 15) / 16, threads);
}

/* ------------------------------------------------------------------------------------------
 * Upsample.compute, kernels main (no hi-res AO) and main_blendout
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    MeaoOracle *o; int lo_level;
    const float *lo_depth, *hi_depth, *lo_ao, *lo_ao2, *hi_ao; float *dest;   /* lo_ao2: LoResAO2 (COMBINE_LOWER_RESOLUTIONS) or NULL */
    int low, loh, hiw, hih;
    float NoiseFilterStrength, StepSize, kBlurTolerance, kUpsampleTolerance;
} ups_ctx;

/* Upsample.compute:74-81 */
static inline float ups_smart_blur(float a, float b, float c, float d, float e, int Left, int Middle, int Right)
{
    b = (Left | Middle) ? b : c;
    a = Left ? a : b;
    d = (Right | Middle) ? d : c;
    e = Right ? e : d;
    return ((a + e) / 2.0f + b + c + d) / 4.0f;
}

/* Upsample.compute:83-87 */
static inline int ups_compare_deltas(const ups_ctx *c, float d1, float d2, float l1, float l2)
{
    float temp = mad(d1, d2, c->StepSize);
    return temp * temp > l1 * l2 * c->kBlurTolerance;
}

/* Upsample.compute:89-130 */
static void ups_blur_h(const ups_ctx *c, const float *AO1, const float *DC, float *AO2, unsigned i)
{
    float a0 = AO1[i], a1 = AO1[i + 1], a2 = AO1[i + 2], a3 = AO1[i + 3], a4 = AO1[i + 4], a5 = AO1[i + 5], a6 = AO1[i + 6];
    float d0 = DC[i], d1 = DC[i + 1], d2 = DC[i + 2], d3 = DC[i + 3], d4 = DC[i + 4], d5 = DC[i + 5], d6 = DC[i + 6];
    float d01 = d1 - d0, d12 = d2 - d1, d23 = d3 - d2, d34 = d4 - d3, d45 = d5 - d4, d56 = d6 - d5;
    float S = c->StepSize;
    float l01 = mad(d01, d01, S), l12 = mad(d12, d12, S), l23 = mad(d23, d23, S);
    float l34 = mad(d34, d34, S), l45 = mad(d45, d45, S), l56 = mad(d56, d56, S);
    int c02 = ups_compare_deltas(c, d01, d12, l01, l12);
    int c13 = ups_compare_deltas(c, d12, d23, l12, l23);
    int c24 = ups_compare_deltas(c, d23, d34, l23, l34);
    int c35 = ups_compare_deltas(c, d34, d45, l34, l45);
    int c46 = ups_compare_deltas(c, d45, d56, l45, l56);
    AO2[i]     = ups_smart_blur(a0, a1, a2, a3, a4, c02, c13, c24);
    AO2[i + 1] = ups_smart_blur(a1, a2, a3, a4, a5, c13, c24, c35);
    AO2[i + 2] = ups_smart_blur(a2, a3, a4, a5, a6, c24, c35, c46);
}

/* Upsample.compute:132-170 */
static void ups_blur_v(const ups_ctx *c, float *AO1, const float *DC, const float *AO2, unsigned i)
{
    float a0 = AO2[i], a1 = AO2[i + 16], a2 = AO2[i + 32], a3 = AO2[i + 48], a4 = AO2[i + 64], a5 = AO2[i + 80];
    float d0 = DC[i + 2], d1 = DC[i + 18], d2 = DC[i + 34], d3 = DC[i + 50], d4 = DC[i + 66], d5 = DC[i + 82];
    float d01 = d1 - d0, d12 = d2 - d1, d23 = d3 - d2, d34 = d4 - d3, d45 = d5 - d4;
    float S = c->StepSize;
    float l01 = mad(d01, d01, S), l12 = mad(d12, d12, S), l23 = mad(d23, d23, S), l34 = mad(d34, d34, S), l45 = mad(d45, d45, S);
    int c02 = ups_compare_deltas(c, d01, d12, l01, l12);
    int c13 = ups_compare_deltas(c, d12, d23, l12, l23);
    int c24 = ups_compare_deltas(c, d23, d34, l23, l34);
    int c35 = ups_compare_deltas(c, d34, d45, l34, l45);
    float r1 = ups_smart_blur(a0, a1, a2, a3, a4, c02, c13, c24);
    float r2 = ups_smart_blur(a1, a2, a3, a4, a5, c13, c24, c35);
    AO1[i] = r1;
    AO1[i + 16] = r2;
}

/* Upsample.compute:177-183.  dot(weights,1) and dot(LowAO,weights) are evaluated x,y,z,w in
 * order with mad contraction. */
static inline float ups_bilateral(const ups_ctx *c, float HiDepth, float HiAO,
                                  float ld0, float ld1, float ld2, float ld3,
                                  float la0, float la1, float la2, float la3)
{
    float t = c->kUpsampleTolerance;
    float w0 = gdiv(9.0f, fabsf(HiDepth - ld0) + t);
    float w1 = gdiv(3.0f, fabsf(HiDepth - ld1) + t);
    float w2 = gdiv(1.0f, fabsf(HiDepth - ld2) + t);
    float w3 = gdiv(3.0f, fabsf(HiDepth - ld3) + t);
    float TotalWeight = (((w0 + w1) + w2) + w3) + c->NoiseFilterStrength;
    float WeightedSum = mad(la3, w3, mad(la2, w2, mad(la1, w1, la0 * w0))) + c->NoiseFilterStrength;
    return gdiv(HiAO * WeightedSum, TotalWeight);
}

static inline void ups_store(const ups_ctx *c, int x, int y, float v)
{
    if (x >= 0 && y >= 0 && x < c->hiw && y < c->hih)
        c->dest[(size_t)y * c->hiw + x] = st_unorm8(c->o, v);
}

/* Upsample.compute:185-233; dispatch ((hi.w+17)/16, (hi.h+17)/16, 1), AO.cs:782-784 */
static void ups_stripe(void *vc, int gy0, int gy1)
{
    ups_ctx *c = (ups_ctx *)vc;
    const int ngx = (c->hiw + 17) / 16;
    float DC[256], AO1[256], AO2[256];                                   /* UPS:50-52 */
    for (int gy = gy0; gy < gy1; gy++)
    for (int gx = XLO(ngx); gx < XHI(ngx); gx++) {
        memset(AO2, 0, sizeof(AO2));   /* row 13 is read (UPS:139) but never written; feeds only an unconsumed output */
        for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++) {
            unsigned index = (unsigned)((tx << 1) | (ty << 5));          /* UPS:191 */
            int cx = gx * 8 + tx + tx - 2, cy = gy * 8 + ty + ty - 2;    /* (DTid + GTid - 2) * InvLowResolution */
            float4_t A = gather4(c->lo_ao, c->low, c->loh, cx, cy);      /* UPS:56 */
            if (c->lo_ao2) {                                             /* UPS:58-60 COMBINE_LOWER_RESOLUTIONS */
                float4_t B = gather4(c->lo_ao2, c->low, c->loh, cx, cy);
                A.x = hmin(A.x, B.x); A.y = hmin(A.y, B.y); A.z = hmin(A.z, B.z); A.w = hmin(A.w, B.w);
            }
            AO1[index] = A.w; AO1[index + 1] = A.z; AO1[index + 16] = A.x; AO1[index + 17] = A.y;   /* UPS:62-65 */
            float4_t D = gather4(c->lo_depth, c->low, c->loh, cx, cy);   /* UPS:67 */
            DC[index] = grcp(D.w); DC[index + 1] = grcp(D.z); DC[index + 16] = grcp(D.x); DC[index + 17] = grcp(D.y);
        }
        /* barrier UPS:192 */
        for (int GI = 0; GI < 39; GI++)                                  /* UPS:199-200 */
            ups_blur_h(c, AO1, DC, AO2, (unsigned)((GI / 3) * 16 + (GI % 3) * 3));
        /* barrier UPS:201 */
        for (int GI = 0; GI < 45; GI++)                                  /* UPS:206-207 */
            ups_blur_v(c, AO1, DC, AO2, (unsigned)((GI / 9) * 32 + GI % 9));
        /* barrier UPS:208 */
        for (int ty = 0; ty < 8; ty++) for (int tx = 0; tx < 8; tx++) {
            int X = gx * 8 + tx, Y = gy * 8 + ty;                        /* DTid.xy */
            unsigned Idx0 = (unsigned)(tx + ty * 16);                    /* UPS:213 */
            float lsx = AO1[Idx0 + 16], lsy = AO1[Idx0 + 17], lsz = AO1[Idx0 + 1], lsw = AO1[Idx0];   /* UPS:214 */
            float4_t Hs = {1.0f, 1.0f, 1.0f, 1.0f};                      /* UPS:223 */
            if (c->hi_ao) Hs = gather4(c->hi_ao, c->hiw, c->hih, 2 * X, 2 * Y);      /* UPS:221 */
            float4_t Ld = gather4(c->lo_depth, c->low, c->loh, X, Y);                /* UPS:225 */
            float4_t Hd = gather4(c->hi_depth, c->hiw, c->hih, 2 * X, 2 * Y);        /* UPS:226 */
            int ox = X << 1, oy = Y << 1;                                            /* UPS:228 */
            ups_store(c, ox - 1, oy,     ups_bilateral(c, Hd.x, Hs.x, Ld.x, Ld.y, Ld.z, Ld.w, lsx, lsy, lsz, lsw));  /* UPS:229 */
            ups_store(c, ox,     oy,     ups_bilateral(c, Hd.y, Hs.y, Ld.y, Ld.z, Ld.w, Ld.x, lsy, lsz, lsw, lsx));  /* UPS:230 */
            ups_store(c, ox,     oy - 1, ups_bilateral(c, Hd.z, Hs.z, Ld.z, Ld.w, Ld.x, Ld.y, lsz, lsw, lsx, lsy));  /* UPS:231 */
            ups_store(c, ox - 1, oy - 1, ups_bilateral(c, Hd.w, Hs.w, Ld.w, Ld.x, Ld.y, Ld.z, lsw, lsx, lsy, lsz));  /* UPS:232 */
        }
    }
}

/* AmbientOcclusion.cs:528-531 (argument wiring) and :750-785 */
void meao_oracle_upsample(MeaoOracle *o, int lo_level, int threads)
{
    ups_ctx c; memset(&c, 0, sizeof(c));
    c.o = o; c.lo_level = lo_level;
    int hi = lo_level - 1;
    c.lo_depth = o->low_depth[lo_level];
    c.lo_ao = (lo_level == 4) ? o->occlusion[4] : o->combined[lo_level];
    if (o->single_scale && lo_level == 1) c.lo_ao = o->occlusion[1];     /* single-scale: LoResAO1 = Occlusion1, nothing coarser contributes */
    c.lo_ao2 = ((o->high_quality_mask >> (lo_level - 1)) & 1) ? o->high_quality[lo_level] : NULL;   /* kernels main_premin / main_premin_blendout */
    c.hi_depth = (hi == 0) ? o->linear_depth : o->low_depth[hi];
    c.hi_ao = (hi == 0) ? NULL : o->occlusion[hi];
    c.dest = (hi == 0) ? o->result : o->combined[hi];
    c.low = o->lw[lo_level]; c.loh = o->lh[lo_level];
    c.hiw = o->lw[hi]; c.hih = o->lh[hi];
    float il[2], ih[2];
    meao_oracle_upsample_constants(o, lo_level, il, ih, &c.NoiseFilterStrength, &c.StepSize, &c.kBlurTolerance, &c.kUpsampleTolerance);
    run_striped(ups_stripe, &c, (c.hih + 17) / 16, threads);
}

/* AmbientOcclusion.cs:511-531 */
void meao_oracle_run(MeaoOracle *o, const float *depth, int threads)
{
    meao_oracle_downsample(o, depth, threads);
    if (o->single_scale) {
        /* BASELINE.json configs[0] "single-scale AO" (SURVEY.md 8d item 1): DS1 + REN level 1 + the final-style UPS only -- three of
         * the ten dispatches of AO.cs:511-531: :513 (Downsample1; Downsample2 also runs, its outputs are unused), :519 (Render
         * TiledDepth1 -> Occlusion1) and :531's dispatch (Upsample kernel "main", LinearDepth as HiResDB) fed Occlusion1 as LoResAO1. */
        meao_oracle_render(o, 1, threads);
        meao_oracle_upsample(o, 1, threads);
        return;
    }
    for (int k = 1; k <= 4; k++) meao_oracle_render(o, k, threads);
    for (int k = 1; k <= 4; k++)                                          /* not in AO.cs: upstream's per-level high-quality pass */
        if ((o->high_quality_mask >> (k - 1)) & 1) meao_oracle_render_wide(o, k, threads);
    for (int lo = 4; lo >= 1; lo--) meao_oracle_upsample(o, lo, threads);
}

/* ------------------------------------------------------------------------------------------
 * Debug views: PushDebugBlitCommands (AmbientOcclusion.cs:787-820) + Blit.shader passes 3 / 4
 * ---------------------------------------------------------------------------------------- */
void meao_oracle_debug_view(const MeaoOracle *o, int debug_id, uint8_t *out)
{
    int sw, sh, slices;
    const float *src = meao_oracle_get_buffer(o, debug_id, &sw, &sh, &slices);
    if (!src) return;
    const int W = o->W, H = o->H;
    for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
        float v;
        if (slices == 16) {
            /* Blit.shader:150-152: uv4 = uv * 4; slice = floor(uv4.x) + floor(uv4.y) * 4; sample at frac(uv4).
             * uv = (2x+1)/(2W): 4*uv = (8x+4)/(2W) -> integer part q, fraction r/(2W); texel = floor(frac * sw) */
            long long nx = 8LL * x + 4, ny = 8LL * y + 4;
            int qx = (int)(nx / (2LL * W)), qy = (int)(ny / (2LL * H));
            long long rx = nx - 2LL * W * qx, ry = ny - 2LL * H * qy;
            int tx = (int)(rx * sw / (2LL * W)), ty = (int)(ry * sh / (2LL * H));
            v = src[((size_t)(qx + 4 * qy) * sh + ty) * sw + tx];
        } else {
            /* cmd.Blit(rt, _result) (AO.cs:817): point-sampled stretch, texel = floor(uv * size) at the pixel centre */
            int tx = (int)((2LL * x + 1) * sw / (2LL * W)), ty = (int)((2LL * y + 1) * sh / (2LL * H));
            v = src[(size_t)ty * sw + tx];
        }
        out[(size_t)y * W + x] = meao_oracle_unorm8_code(v);               /* R8 render target store (AO.cs:475) */
    }
}

/* ------------------------------------------------------------------------------------------
 * Composite (Blit.shader passes 1 and 2; AmbientOcclusion.cs:822-839)
 * ---------------------------------------------------------------------------------------- */
static void scale_px(void *rgba, int is_half, size_t i, float f, int rgb, int alpha)
{
    for (int ch = 0; ch < 4; ch++) {
        if (!((ch < 3) ? rgb : alpha)) continue;                    /* blend factor 1: destination unchanged */
        if (is_half) {
            uint16_t *p = (uint16_t *)rgba + i * 4 + ch;
            *p = meao_oracle_f32_to_f16_bits(meao_oracle_f16_bits_to_f32(*p) * f);
        } else {
            uint8_t *p = (uint8_t *)rgba + i * 4 + ch;
            *p = meao_oracle_unorm8_code(((float)*p * (1.0f / 255.0f)) * f);
        }
    }
}

/* pass 2: return tex2D(_AOTexture, uv).r  with  Blend Zero SrcAlpha  (Blit.shader:86,96-99) */
void meao_oracle_composite_framebuffer(const uint8_t *ao_codes, void *rgba, int is_half, size_t npix)
{
    for (size_t i = 0; i < npix; i++) {
        float t = (float)ao_codes[i] * (1.0f / 255.0f);
        scale_px(rgba, is_half, i, t, 1, 1);
    }
}

/* pass 1: ao = 1 - tex.r; gbuffer0 = (0,0,0,ao), gbuffer3 = (ao,ao,ao,0) with
 * Blend Zero OneMinusSrcColor, Zero OneMinusSrcAlpha  (Blit.shader:68,83-89) */
void meao_oracle_composite_gbuffer(const uint8_t *ao_codes, uint8_t *gbuffer0_rgba8, void *gbuffer3_rgba, int g3_is_half, size_t npix)
{
    for (size_t i = 0; i < npix; i++) {
        float t = (float)ao_codes[i] * (1.0f / 255.0f);
        volatile float ao = 1.0f - t;                               /* Blit.shader:83 */
        float f = 1.0f - ao;                                        /* OneMinusSrc* */
        scale_px(gbuffer0_rgba8, 0, i, f, 0, 1);
        scale_px(gbuffer3_rgba, g3_is_half, i, f, 1, 0);
    }
}

/* pass 3: return tex2D(_AOTexture, uv).r to all four channels, no blend state (Blit.shader:116-134; AO.cs:826-829) */
void meao_oracle_composite_debug(const uint8_t *view_codes, void *rgba, int is_half, size_t npix)
{
    for (size_t i = 0; i < npix; i++) {
        float r = (float)view_codes[i] * (1.0f / 255.0f);
        for (int ch = 0; ch < 4; ch++) {
            if (is_half) ((uint16_t *)rgba)[i * 4 + ch] = meao_oracle_f32_to_f16_bits(r);
            else ((uint8_t *)rgba)[i * 4 + ch] = meao_oracle_unorm8_code(r);
        }
    }
}
