/* Pure-C client of include/meao.h: proves the drop-in boundary needs nothing but a C compiler and libmeao.so.
 *   smoke plan                       -> planning-only context (no GPU): constants, geometry, loud failure of compute calls
 *   smoke render W H depth.f32 ao.u8 intensity -> full frame through meao_render_host, compared with the expected AO bytes
 *   smoke bands W H depth.f32 ao.u8 intensity nbands ndevices -> the SAME frame as nbands row bands (band i on device i % ndevices),
 *                                   connected through meao_band_export / meao_band_connect and stepped with meao_band_step_host:
 *                                   multi-GPU from one single-threaded C process, no NCCL, no Python
 * exit code 0 = ok. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "meao.h"

static int fail(const char *what, MeaoCtx *c) { fprintf(stderr, "FAIL %s: %s\n", what, meao_last_error(c)); return 1; }

static void *slurp(const char *path, size_t bytes)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    void *p = malloc(bytes);
    size_t n = fread(p, 1, bytes, f);
    fclose(f);
    if (n != bytes) { free(p); return NULL; }
    return p;
}

int main(int argc, char **argv)
{
    if (argc < 2 || meao_abi_version() != MEAO_ABI_VERSION) return 2;
    MeaoCtx *c = NULL;
    if (!strcmp(argv[1], "plan")) {
        MeaoDeviceCfg cfg = {-1, MEAO_FLAG_NONE};
        if (meao_create(&cfg, &c)) return fail("create(plan)", NULL);
        MeaoCamera cam = {0.3f, 100.0f, 1.0264f, 1};
        if (meao_set_camera(c, &cam) || meao_resize(c, 3840, 2160) < 0) return fail("plan setup", c);
        float rc[28], uc[8];
        if (meao_render_constants(c, 1, rc) || meao_upsample_constants(c, 1, uc)) return fail("constants", c);
        float sum = 0; for (int i = 12; i < 24; i++) sum += rc[i];
        if (sum < 0.999999f || sum > 1.000001f) { fprintf(stderr, "weights do not sum to 1: %g\n", sum); return 1; }
        if (meao_algorithmic_bytes(c, 0) != 131613600LL) return fail("algorithmic bytes", c);
        /* ABI v2: the undispatched shader variants are plan inputs like the parameters */
        MeaoVariants v = {0, 1, 12, 0}, back;
        float rw[28], re[28];
        if (meao_kernels_per_frame(c) != 9 || meao_set_variants(c, &v) != 1 || meao_set_variants(c, &v) != 0) return fail("set_variants", c);
        if (meao_get_variants(c, &back) || back.high_quality_mask != 12 || back.sample_exhaustively != 1) return fail("get_variants", c);
        if (meao_kernels_per_frame(c) != 11 || meao_render_constants(c, 1, re) || meao_render_constants_wide(c, 1, rw)) return fail("variant constants", c);
        if (!(re[12] > 0.0f) || rc[12] != 0.0f) { fprintf(stderr, "exhaustive mode must keep weight slot 0 (AO.cs:711)\n"); return 1; }
        if (rw[24] != 1.0f / 1920.0f || re[24] != 1.0f / 480.0f) { fprintf(stderr, "gInvSliceDimension of the tiled / non-tiled source\n"); return 1; }
        v.high_quality_mask = 16;
        if (meao_set_variants(c, &v) != MEAO_ERR_INVALID) { fprintf(stderr, "mask 16 must be refused\n"); return 1; }
        float d = 0; uint8_t o = 0;
        if (meao_render_host(c, &d, MEAO_DEPTH_RAW_F32, &o) != MEAO_ERR_CUDA) { fprintf(stderr, "compute on a plan-only context must fail\n"); return 1; }
        meao_destroy(c);
        printf("plan ok\n");
        return 0;
    }
    if (!strcmp(argv[1], "render") && argc >= 7) {
        const int W = atoi(argv[2]), H = atoi(argv[3]);
        const size_t n = (size_t)W * H;
        float *depth = (float *)slurp(argv[4], n * 4);
        uint8_t *want = (uint8_t *)slurp(argv[5], n);
        if (!depth || !want) { fprintf(stderr, "cannot read inputs\n"); return 2; }
        MeaoDeviceCfg cfg = {0, MEAO_FLAG_NONE};
        if (meao_create(&cfg, &c)) return fail("create", NULL);
        MeaoParams p; meao_default_params(&p); p.intensity = (float)atof(argv[6]);
        /* projection of a 60 degree camera: tanHalfFovH = aspect * tan(30 deg) */
        MeaoCamera cam = {0.3f, 100.0f, (float)((double)W / H * 0.57735026918962576), 1};
        if (meao_set_params(c, &p) < 0 || meao_set_camera(c, &cam) || meao_resize(c, W, H) < 0) return fail("setup", c);
        uint8_t *got = (uint8_t *)malloc(n);
        if (meao_render_host(c, depth, MEAO_DEPTH_RAW_F32, got)) return fail("render_host", c);
        size_t bad = 0; for (size_t i = 0; i < n; i++) bad += got[i] != want[i];
        printf("render %dx%d: %zu mismatching pixels, %lld kernel launches\n", W, H, bad, (long long)meao_launch_count(c));
        meao_destroy(c);
        return bad ? 1 : 0;
    }
    if (!strcmp(argv[1], "bands") && argc >= 9) {
        const int W = atoi(argv[2]), H = atoi(argv[3]), nb = atoi(argv[7]), ndev = atoi(argv[8]);
        const size_t n = (size_t)W * H;
        float *depth = (float *)slurp(argv[4], n * 4);
        uint8_t *want = (uint8_t *)slurp(argv[5], n);
        if (!depth || !want || nb < 2 || nb > 16 || ndev < 1) { fprintf(stderr, "bad arguments\n"); return 2; }
        MeaoCtx *b[16]; MeaoPeerHandle h[16]; int cut[17];
        const int blocks = (H + 15) / 16;
        for (int i = 0; i < nb; i++) { cut[i] = 16 * ((blocks * i) / nb); if (cut[i] > H) cut[i] = H; }
        cut[nb] = H;
        MeaoParams p; meao_default_params(&p); p.intensity = (float)atof(argv[6]);
        MeaoCamera cam = {0.3f, 100.0f, (float)((double)W / H * 0.57735026918962576), 1};
        for (int i = 0; i < nb; i++) {
            MeaoDeviceCfg cfg = {i % ndev, MEAO_FLAG_NONE};
            if (meao_create(&cfg, &b[i])) return fail("create", NULL);
            if (meao_set_params(b[i], &p) < 0 || meao_set_camera(b[i], &cam) || meao_resize(b[i], W, H) < 0) return fail("setup", b[i]);
            if (meao_set_row_band(b[i], cut[i], cut[i + 1], i > 0 ? cut[i - 1] : -1, i + 1 < nb ? cut[i + 2] : -1)) return fail("set_row_band", b[i]);
            if (meao_band_export(b[i], &h[i])) return fail("band_export", b[i]);
        }
        for (int i = 0; i < nb; i++) {
            if (i > 0 && meao_band_connect(b[i], 0, &h[i - 1])) return fail("band_connect up", b[i]);
            if (i + 1 < nb && meao_band_connect(b[i], 1, &h[i + 1])) return fail("band_connect down", b[i]);
        }
        /* PINNED host memory: with pageable buffers the copies would block this (only) host thread inside band 0's call, band 1
         * would never be enqueued, and band 0's exchange kernel would wait for it until the time-out */
        float *pdepth = (float *)meao_host_alloc(n * 4);
        uint8_t *got = (uint8_t *)meao_host_alloc(n);
        if (!pdepth || !got) { fprintf(stderr, "meao_host_alloc failed\n"); return 1; }
        memcpy(pdepth, depth, n * 4);
        free(depth); depth = pdepth;
        size_t bad = 0;
        for (int rep = 0; rep < 3; rep++) {              /* replays of the captured graphs through the same epoch flags */
            memset(got, 0, n);
            for (int i = 0; i < nb; i++)                 /* enqueue every band, THEN wait: the exchange kernels handshake on the devices */
                if (meao_band_step_host(b[i], depth + (size_t)cut[i] * W, MEAO_DEPTH_RAW_F32, got + (size_t)cut[i] * W)) return fail("band_step_host", b[i]);
            for (int i = 0; i < nb; i++) if (meao_host_wait(b[i], 0)) return fail("host_wait", b[i]);
            for (int i = 0; i < nb; i++) { int32_t st[4]; if (meao_band_status(b[i], st) || st[1] != 0) { fprintf(stderr, "band %d: exchange error %d\n", i, st[1]); return 1; } }
            bad = 0; for (size_t k = 0; k < n; k++) bad += got[k] != want[k];
            if (bad) break;
        }
        printf("bands %dx%d, %d bands on %d device(s): %zu mismatching pixels\n", W, H, nb, ndev, bad);
        for (int i = 0; i < nb; i++) meao_destroy(b[i]);
        meao_host_free(pdepth); meao_host_free(got);
        return bad ? 1 : 0;
    }
    return 2;
}
