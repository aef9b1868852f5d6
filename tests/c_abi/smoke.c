/* Pure-C client of include/meao.h: proves the drop-in boundary needs nothing but a C compiler and libmeao.so.
 *   smoke plan                       -> planning-only context (no GPU): constants, geometry, loud failure of compute calls
 *   smoke render W H depth.f32 ao.u8 intensity -> full frame through meao_render_host, compared with the expected AO bytes
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
    return 2;
}
