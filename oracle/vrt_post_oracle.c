/*
 * vrt_post_oracle.c — CPU restatement of the reference's image-space step AFTER the traversal path
 * (SURVEY.md §8f row N4): tiled-framebuffer blit, temporal reprojection, SVGF variance estimation,
 * à-trous passes and the tone-mapped present.  TEST INFRASTRUCTURE ONLY: nothing in the product
 * path may link or call this file (tests/, __graft_entry__.smoke() and bench.py's CPU leg only).
 *
 * PARITY UNPINNED: the reference implements this step as GLSL compute/fragment shaders and there
 * is no GL device in this image, so no output of the reference itself exists to pin against, and
 * GLSL leaves the rounding of exp/pow/sqrt, FMA contraction and f16 stores implementation-defined.
 * This file therefore DEFINES the canonical arithmetic (below) and follows the shaders statement
 * by statement, one texture per reference texture, in the reference's dispatch order:
 *
 *   src/VoxelRT/Shaders/CopyTiledFramebuffer.comp:10-37   blit of Framebuffer::Tile -> G-buffer textures
 *   src/VoxelRT/Shaders/GBuffer.glsl:15-24                normal packing, bounds test
 *   src/VoxelRT/Shaders/Denoise/Reproject.comp:9-108      temporal accumulation
 *   src/VoxelRT/Shaders/Denoise/Filter.comp:17-146        variance estimate (pass -1), à-trous (pass >= 0)
 *   src/VoxelRT/Shaders/GBufferBlit.frag:8-46             ACES + gamma present, debug channels
 *   src/VoxelRT/GBuffer.h:31-130                          texture set, SetCamera swaps, pass/buffer rotation
 *
 * Canonical arithmetic: IEEE binary32, round-to-nearest-even, every operation evaluated separately in
 * the order written (no contraction: built with -ffp-contract=off; FMA only where fmaf() is spelled);
 * vec dot / mat*vec accumulate left to right; mix(a,b,t) = a*(1-t) + b*t (GLSL definition);
 * max(a,b) = a<b ? b : a, min(a,b) = b<a ? b : a (NaN keeps the first operand); rgba16f / rg16f image
 * stores round to nearest even; unorm8 stores are (int)(clamp(c,0,1)*255 + 0.5); exp() and log() are
 * the polynomial forms post_exp / post_log below (relative error < 3e-7, bit-reproducible on any IEEE
 * machine); pow(x,128) is seven squarings; pow(c,0.45) = post_exp(0.45 * post_log(c)), 0 for c below the
 * smallest normal.  The two per-tap quotients of the filters (w_luma = |dl| / lumaPhi, w_depth = |dd| / (length + 0.001),
 * Filter.comp:51,57,117,123) are evaluated as a * (1/b) with a correctly rounded reciprocal — what GPU GLSL compilers emit
 * for `/` (the spec allows 2.5 ulp), and the divisor is per pixel / per tap, so the reciprocal is shared by all taps.
 * The luminance dot product and the running sums of the two filters (sum += value * weight, Filter.comp:60-62,128-129, and
 * the 3x3 variance prefilter :79) are FMAs — the contraction GPU compilers apply to GLSL `a += b * c`; nothing else is fused.
 * Out-of-range image loads return 0 (robust buffer access).
 *
 * One documented deviation: Reproject.comp reads u_HistoryLenTex at NEIGHBOUR positions (:78) while
 * other invocations of the same dispatch store to it (:100,106) — a data race whose outcome depends
 * on scheduling.  Here (and in the CUDA path) every such read sees the value from BEFORE the dispatch.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PO_API __attribute__((visibility("default")))

/* ---- scalar helpers ------------------------------------------------------------------------- */
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float fmax_g(float a, float b) { return a < b ? b : a; }
static inline float fmin_g(float a, float b) { return b < a ? b : a; }

/* f32 -> f16 RNE, f16 -> f32 (same conversion the traversal oracle uses for the irradiance stores) */
static uint16_t f32_to_f16(float f) {
    uint32_t x = f2u(f), sign = (x >> 16) & 0x8000u, ax = x & 0x7FFFFFFFu;
    if (ax > 0x7F800000u) return 0x7E00u; /* NaN: one canonical quiet NaN */
    if (ax >= 0x47800000u) return (uint16_t)(sign | 0x7C00u); /* >= 65536 -> inf (65520 rounds up below) */
    if (ax >= 0x38800000u) {
        uint32_t m = ax - 0x38000000u; /* rebias */
        uint32_t h = m >> 13, rem = m & 0x1FFFu;
        if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) h++;
        return (uint16_t)(sign | h); /* carry into the exponent (and into inf) is correct */
    }
    if (ax < 0x33000000u) return (uint16_t)sign; /* < 2^-25 -> 0 */
    uint32_t e = ax >> 23, man = (ax & 0x7FFFFFu) | 0x800000u;
    uint32_t shift = 126 - e; /* 14..24 */
    uint32_t h = man >> shift, rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) h++;
    return (uint16_t)(sign | h);
}
static float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 31, m = h & 0x3FFu;
    if (e == 0) {
        if (m == 0) return u2f(sign);
        float v = (float)m * 5.9604644775390625e-08f; /* 2^-24, exact */
        return sign ? -v : v;
    }
    if (e == 31) return u2f(sign | 0x7F800000u | (m << 13));
    return u2f(sign | ((e + 112) << 23) | (m << 13));
}

/* exp(x): n = rint(x*log2 e); r = x - n*ln2 (two-step, FMA); degree-6 Taylor in Horner/FMA form; scale by 2^n.
 * Flushes to 0 below -87 (the result would be subnormal), saturates the argument at 88, propagates NaN. */
PO_API float post_exp(float x) {
    if (x != x) return x;
    if (!(x > -87.0f)) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    float n = rintf(x * 1.44269502f);
    float r = fmaf(n, -0.693145752f, x);
    r = fmaf(n, -1.42860677e-06f, r);
    float p = 1.38888892e-03f;
    p = fmaf(p, r, 8.33333377e-03f);
    p = fmaf(p, r, 4.16666679e-02f);
    p = fmaf(p, r, 1.66666672e-01f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    return p * u2f((uint32_t)((int)n + 127) << 23);
}
/* log(x) for normal positive x: x = m*2^e with m in (sqrt(.5), sqrt 2]; s = (m-1)/(m+1);
 * log m = 2 s (1 + s^2/3 + s^4/5 + s^6/7 + s^8/9). */
PO_API float post_log(float x) {
    uint32_t ix = f2u(x);
    int e = (int)(ix >> 23) - 127;
    float m = u2f((ix & 0x7FFFFFu) | 0x3F800000u);
    if (m > 1.41421354f) { m = m * 0.5f; e += 1; }
    float s = (m - 1.0f) / (m + 1.0f);
    float s2 = s * s;
    float p = 0.111111112f;
    p = fmaf(p, s2, 0.142857149f);
    p = fmaf(p, s2, 0.2f);
    p = fmaf(p, s2, 0.333333343f);
    p = fmaf(p, s2, 1.0f);
    float lm = (2.0f * s) * p;
    return fmaf((float)e, 0.693147182f, lm);
}
static float pow045(float c) { /* pow(color, vec3(0.45)), GBufferBlit.frag:29 */
    if (!(c >= 1.17549435e-38f)) return 0.0f;
    return post_exp(0.45f * post_log(c));
}
static float pow128(float x) { /* pow(x, 128), Filter.comp:54,120 */
    for (int i = 0; i < 7; i++) x = x * x;
    return x;
}

/* ---- texture set (GBuffer.h:12-16,33-47) ---------------------------------------------------- */
typedef struct { uint16_t c[4]; } Half4;
typedef struct { uint16_t c[2]; } Half2;
typedef struct PostOracle {
    int w, h;
    uint32_t *albedo, *prev_albedo;         /* rgba8, a = packed normal */
    Half4 *irr, *prev_irr, *temp_irr;       /* rgba16f                  */
    float *depth, *prev_depth;              /* r32f                     */
    Half2 *moments, *prev_moments;          /* rg16f                    */
    uint8_t *hist, *hist_snapshot;          /* r8ui (+ pre-dispatch copy, see header) */
    float cur_proj[16], hist_proj[16], cur_inv[16], hist_inv[16];
    double cur_pos[3], hist_pos[3];
    uint32_t frame_no;
    int have_camera;
} PostOracle;

PO_API PostOracle* post_oracle_create(int w, int h) {
    PostOracle* o = (PostOracle*)calloc(1, sizeof(PostOracle));
    size_t n = (size_t)w * h;
    o->w = w; o->h = h;
    o->albedo = calloc(n, 4); o->prev_albedo = calloc(n, 4);
    o->irr = calloc(n, 8); o->prev_irr = calloc(n, 8); o->temp_irr = calloc(n, 8);
    o->depth = calloc(n, 4); o->prev_depth = calloc(n, 4);
    o->moments = calloc(n, 4); o->prev_moments = calloc(n, 4);
    o->hist = calloc(n, 1); o->hist_snapshot = calloc(n, 1);
    return o;
}
PO_API void post_oracle_destroy(PostOracle* o) {
    if (!o) return;
    free(o->albedo); free(o->prev_albedo); free(o->irr); free(o->prev_irr); free(o->temp_irr);
    free(o->depth); free(o->prev_depth); free(o->moments); free(o->prev_moments); free(o->hist); free(o->hist_snapshot);
    free(o);
}
#define SWAP(T, a, b) do { T _t = a; a = b; b = _t; } while (0)

static inline int in_bounds(const PostOracle* o, int x, int y) { /* gbufferCheckBounds, GBuffer.glsl:22-24 */
    return (uint32_t)x < (uint32_t)o->w && (uint32_t)y < (uint32_t)o->h;
}
static inline void unpack_normal(uint32_t albedo_normal, float n[3]) { /* unpackGNormal, GBuffer.glsl:15-17 */
    uint32_t a = albedo_normal >> 24;
    n[0] = (float)(a & 3u) - 1.0f; n[1] = (float)((a >> 2) & 3u) - 1.0f; n[2] = (float)((a >> 4) & 3u) - 1.0f;
}
static inline float dot3(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline float luminance(const float c[3]) { return fmaf(c[2], 0.114f, fmaf(c[1], 0.587f, c[0] * 0.299f)); }
static inline void load_h4(const Half4* t, size_t i, float v[4]) { for (int k = 0; k < 4; k++) v[k] = f16_to_f32(t[i].c[k]); }
static inline void store_h4(Half4* t, size_t i, const float v[4]) { for (int k = 0; k < 4; k++) t[i].c[k] = f32_to_f16(v[k]); }
static inline void mat_vec(const float m[16], float x, float y, float z, float w, float r[4]) {
    for (int i = 0; i < 4; i++) r[i] = m[i] * x + m[4 + i] * y + m[8 + i] * z + m[12 + i] * w;
}

/* GBuffer::SetCamera (GBuffer.h:31-60).  inv_proj = GetInverseProjScreenMat(proj, viewSize), computed by the
 * caller (the same floats go to the CUDA path).  The very first call has no history: the reference leaves
 * HistoryProj/HistoryPos default-constructed; here they start equal to the current ones, textures zeroed. */
PO_API void post_oracle_set_camera(PostOracle* o, const float proj[16], const float inv_proj[16], const double pos[3]) {
    if (!o->have_camera) {
        memcpy(o->cur_proj, proj, 64); memcpy(o->cur_inv, inv_proj, 64); memcpy(o->cur_pos, pos, 24);
        o->have_camera = 1;
    }
    memcpy(o->hist_proj, o->cur_proj, 64); memcpy(o->hist_inv, o->cur_inv, 64); memcpy(o->hist_pos, o->cur_pos, 24);
    memcpy(o->cur_proj, proj, 64); memcpy(o->cur_inv, inv_proj, 64); memcpy(o->cur_pos, pos, 24);
    SWAP(uint32_t*, o->albedo, o->prev_albedo);
    SWAP(float*, o->depth, o->prev_depth);
    SWAP(Half2*, o->moments, o->prev_moments);
    o->frame_no++;
}

/* CopyTiledFramebuffer.comp:10-37 with TileShiftX = TileShiftY = 2 (16-lane packets) */
static void blit_tiles(PostOracle* o, const uint32_t* tiles) {
    const uint32_t stride = (uint32_t)o->w >> 2, field = 16, tile_words = 64;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < o->h; y++)
        for (int x = 0; x < o->w; x++) {
            uint32_t off = (((uint32_t)x >> 2) + ((uint32_t)y >> 2) * stride) * tile_words + ((uint32_t)x & 3) + (((uint32_t)y & 3) << 2);
            uint32_t a = tiles[off];
            float depth = u2f(tiles[off + field]);
            uint32_t rg = tiles[off + 2 * field], bx = tiles[off + 3 * field];
            uint32_t rgb = a & 0xFFFFFFu;
            if (depth < 0.0f) rgb = 0xFFFFFFu;                    /* :31 */
            uint32_t code = (a >> 24) & 0x3Fu;                    /* (n+1) fields at bits 24,26,28 -> packGNormal */
            size_t i = (size_t)y * o->w + x;
            o->albedo[i] = rgb | (code << 24);
            o->irr[i].c[0] = (uint16_t)rg; o->irr[i].c[1] = (uint16_t)(rg >> 16);
            o->irr[i].c[2] = (uint16_t)bx; o->irr[i].c[3] = 0;    /* :34 vec4(rg, b, 0) */
            o->depth[i] = depth;
        }
}

/* getWorldPos, Reproject.comp:13-16 */
static void world_pos(const float inv[16], int sx, int sy, float depth, float out[3]) {
    float v[4];
    mat_vec(inv, (float)sx, (float)sy, depth, 1.0f, v);
    float s = 16.0f / v[3];
    out[0] = v[0] * s; out[1] = v[1] * s; out[2] = v[2] * s;
}

/* isValidSample, Reproject.comp:26-41 */
static int valid_sample(const PostOracle* o, int sx, int sy, const float cw[3], const float cn[3], const float delta[3]) {
    if (!in_bounds(o, sx, sy)) return 0; /* texelFetch outside the texture: zeros -> depth 0 -> rejected */
    size_t i = (size_t)sy * o->w + sx;
    float sn[3];
    unpack_normal(o->prev_albedo[i], sn);
    if (dot3(cn, sn) < 0.5f) return 0;
    float sd = o->prev_depth[i];
    if (sd <= 0.0f) return 0;
    float sw[3], d[3];
    world_pos(o->hist_inv, sx, sy, sd, sw);
    for (int k = 0; k < 3; k++) d[k] = cw[k] - sw[k] + delta[k];
    float pd = fabsf(dot3(d, cn));
    if (pd > 6.0f) return 0;
    return 1;
}

/* reproject(), Reproject.comp:43-103 */
static int reproject_px(PostOracle* o, int x, int y, const float delta[3], int reset) {
    size_t i = (size_t)y * o->w + x;
    float depth = o->depth[i];
    if (depth <= 0.0f) return 0;
    float wp[3];
    world_pos(o->cur_inv, x, y, depth, wp);
    /* getHistoryPos :17-23 */
    float ndc[4];
    mat_vec(o->hist_proj, wp[0] + delta[0], wp[1] + delta[1], wp[2] + delta[2], 1.0f, ndc);
    float px = (ndc[0] / ndc[3] * 0.5f + 0.5f) * (float)o->w - 0.5f;
    float py = (ndc[1] / ndc[3] * 0.5f + 0.5f) * (float)o->h - 0.5f;
    if (!(px > -1.0e9f && px < 1.0e9f && py > -1.0e9f && py < 1.0e9f)) return 0; /* ivec2() of NaN / huge: undefined in GLSL */
    int pxi = (int)px, pyi = (int)py; /* truncation */
    float fx = px - floorf(px), fy = py - floorf(py);
    if (!in_bounds(o, pxi, pyi)) return 0;
    float cn[3];
    unpack_normal(o->albedo[i], cn);
    float wsum = 0.0f, pirr[3] = {0, 0, 0}, pmom[2] = {0, 0};
    uint32_t hl = o->hist_snapshot[i];
    for (int s = 0; s < 4; s++) {
        int sx = pxi + (s & 1), sy = pyi + (s >> 1);
        if (!valid_sample(o, sx, sy, wp, cn, delta)) continue;
        float w = ((s & 1) ? fx : 1.0f - fx) * ((s >> 1) ? fy : 1.0f - fy);
        size_t j = (size_t)sy * o->w + sx;
        float v[4];
        load_h4(o->prev_irr, j, v);
        for (int k = 0; k < 3; k++) pirr[k] = pirr[k] + v[k] * w;
        pmom[0] = pmom[0] + f16_to_f32(o->prev_moments[j].c[0]) * w;
        pmom[1] = pmom[1] + f16_to_f32(o->prev_moments[j].c[1]) * w;
        wsum = wsum + w;
        uint32_t nh = (uint32_t)o->hist_snapshot[j] + 1u;
        if (nh < hl) hl = nh;
    }
    if (wsum < 0.001f) return 0;
    for (int k = 0; k < 3; k++) pirr[k] = pirr[k] / wsum;
    pmom[0] = pmom[0] / wsum; pmom[1] = pmom[1] / wsum;
    if (reset && hl > 6) hl = 6;
    float blend = 1.0f / (float)(hl + 1u);
    float cur[4], nirr[4];
    load_h4(o->irr, i, cur);
    for (int k = 0; k < 3; k++) nirr[k] = pirr[k] * (1.0f - blend) + cur[k] * blend;
    float luma = luminance(nirr);
    float mb = fmax_g(0.5f, blend);
    float nm0 = pmom[0] * (1.0f - mb) + luma * mb;
    float nm1 = pmom[1] * (1.0f - mb) + (luma * luma) * mb;
    o->moments[i].c[0] = f32_to_f16(nm0); o->moments[i].c[1] = f32_to_f16(nm1);
    nirr[3] = fmax_g(0.0f, nm1 - nm0 * nm0);
    store_h4(o->irr, i, nirr);
    hl = hl + 1u < 64u ? hl + 1u : 64u;
    o->hist[i] = (uint8_t)hl;
    return 1;
}
static void reproject_pass(PostOracle* o, int reset) {
    float delta[3];
    for (int k = 0; k < 3; k++) delta[k] = (float)(o->cur_pos[k] - o->hist_pos[k]); /* GBuffer.h:82 */
    memcpy(o->hist_snapshot, o->hist, (size_t)o->w * o->h);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < o->h; y++)
        for (int x = 0; x < o->w; x++)
            if (!reproject_px(o, x, y, delta, reset)) { /* main(), Reproject.comp:104-108 */
                size_t i = (size_t)y * o->w + x;
                o->hist[i] = 0; o->moments[i].c[0] = 0; o->moments[i].c[1] = 0;
            }
}

/* varianceEstim, Filter.comp:17-68: in u_IrradianceTex, out u_TempIrradianceTex; note that the centre
 * luminance is read from the TEMP texture (:26), i.e. whatever the previous frame left there. */
static void variance_pass(PostOracle* o) {
#pragma omp parallel for schedule(static) /* every pixel reads o->irr and its OWN temp texel only */
    for (int y = 0; y < o->h; y++)
        for (int x = 0; x < o->w; x++) {
            size_t i = (size_t)y * o->w + x;
            uint32_t hl = o->hist[i];
            float cd = o->depth[i];
            if (hl > 4 || cd < 0.0f) { o->temp_irr[i] = o->irr[i]; continue; }
            float ct[4], cn[3];
            load_h4(o->temp_irr, i, ct);
            float cl = luminance(ct);
            unpack_normal(o->albedo[i], cn);
            const float luma_phi = 10.0f;
            float si[3] = {0, 0, 0}, sm[2] = {0, 0}, wsum = 0.0f;
            for (int ky = -3; ky <= 3; ky++)
                for (int kx = -3; kx <= 3; kx++) {
                    int sx = x + kx, sy = y + ky;
                    if (!in_bounds(o, sx, sy)) continue;
                    size_t j = (size_t)sy * o->w + sx;
                    float v[4], n[3];
                    load_h4(o->irr, j, v);
                    float l = luminance(v);
                    float w_luma = fabsf(l - cl) * (1.0f / luma_phi);
                    unpack_normal(o->albedo[j], n);
                    float w_normal = pow128(fmin_g(fmax_g(dot3(n, cn), 0.001f), 1.0f));
                    float w_depth = fabsf(cd - o->depth[j]) * (1.0f / (sqrtf((float)kx * (float)kx + (float)ky * (float)ky) + 0.001f));
                    float w = post_exp(-(w_luma + w_depth)) * w_normal;
                    for (int k = 0; k < 3; k++) si[k] = fmaf(v[k], w, si[k]);
                    sm[0] = fmaf(l, w, sm[0]); sm[1] = fmaf(l * l, w, sm[1]);
                    wsum = wsum + w;
                }
            wsum = fmax_g(wsum, 0.001f);
            float out[4];
            for (int k = 0; k < 3; k++) out[k] = si[k] / wsum;
            sm[0] = sm[0] / wsum; sm[1] = sm[1] / wsum;
            float var = fmax_g(0.0f, sm[1] - sm[0] * sm[0]);
            var = var * ((4.0f - (float)hl) * 3.0f);
            out[3] = var;
            store_h4(o->temp_irr, i, out);
        }
}

/* svgfAtrous + getFilteredVariance, Filter.comp:70-135 */
static void atrous_pass(PostOracle* o, const Half4* in, Half4* out, int pass_no) {
    static const float kvar[2][2] = {{0.25f, 0.125f}, {0.125f, 0.0625f}};
    static const float kern[3] = {0.375f, 0.25f, 0.0625f};
#pragma omp parallel for schedule(static)
    for (int y = 0; y < o->h; y++)
        for (int x = 0; x < o->w; x++) {
            size_t i = (size_t)y * o->w + x;
            float cd = o->depth[i];
            float ci[4];
            load_h4(in, i, ci);
            if (cd < 0.0f) { out[i] = in[i]; continue; }
            float cv = 0.0f;
            for (int ky = -1; ky <= 1; ky++)
                for (int kx = -1; kx <= 1; kx++) {
                    float v = in_bounds(o, x + kx, y + ky) ? f16_to_f32(in[(size_t)(y + ky) * o->w + (x + kx)].c[3]) : 0.0f;
                    cv = fmaf(v, kvar[abs(kx)][abs(ky)], cv);
                }
            float cn[3];
            unpack_normal(o->albedo[i], cn);
            float cl = luminance(ci);
            float luma_phi = sqrtf(fmax_g(0.0001f, cv)) * 4.0f;
            float sum[4] = {ci[0], ci[1], ci[2], ci[3]}, wsum = 1.0f;
            for (int ky = -2; ky <= 2; ky++)
                for (int kx = -2; kx <= 2; kx++) {
                    if (kx == 0 && ky == 0) continue;
                    int ox = kx * (1 << pass_no), oy = ky * (1 << pass_no);
                    int sx = x + ox, sy = y + oy;
                    if (!in_bounds(o, sx, sy)) continue;
                    size_t j = (size_t)sy * o->w + sx;
                    float v[4], n[3];
                    load_h4(in, j, v);
                    float w_luma = fabsf(luminance(v) - cl) * (1.0f / luma_phi);
                    unpack_normal(o->albedo[j], n);
                    float w_normal = pow128(fmin_g(fmax_g(dot3(n, cn), 0.001f), 1.0f));
                    float w_depth = fabsf(cd - o->depth[j]) * (1.0f / (sqrtf((float)ox * (float)ox + (float)oy * (float)oy) + 0.001f));
                    float w = kern[abs(kx)] * kern[abs(ky)];
                    w = w * (post_exp(-(w_luma + w_depth)) * w_normal);
                    for (int k = 0; k < 3; k++) sum[k] = fmaf(v[k], w, sum[k]);
                    sum[3] = fmaf(v[3], w * w, sum[3]);
                    wsum = wsum + w;
                }
            if (wsum < 0.001f) wsum = 0.001f;
            for (int k = 0; k < 3; k++) sum[k] = sum[k] / wsum;
            sum[3] = sum[3] / (wsum * wsum);
            store_h4(out, i, sum);
        }
}

static void aces(const float in[3], float out[3]) { /* aces_approx, GBufferBlit.frag:8-16 */
    for (int k = 0; k < 3; k++) {
        float v = in[k] * 0.6f;
        float r = (v * (2.51f * v + 0.03f)) / (v * (2.43f * v + 0.59f) + 0.14f);
        out[k] = r < 0.0f ? 0.0f : (r > 1.0f ? 1.0f : r);
    }
}
static uint32_t unorm8(float c) {
    if (!(c > 0.0f)) return 0;
    if (c > 1.0f) c = 1.0f;
    return (uint32_t)(c * 255.0f + 0.5f);
}
/* GBufferBlit.frag:18-46 */
static void present_pass(const PostOracle* o, int debug_channel, uint32_t* rgba) {
    const int64_t n = (int64_t)o->w * o->h;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        uint32_t a = o->albedo[i];
        float alb[3] = {(float)(a & 255u) / 255.0f, (float)((a >> 8) & 255u) / 255.0f, (float)((a >> 16) & 255u) / 255.0f};
        float irr[4], c[3];
        load_h4(o->irr, i, irr);
        for (int k = 0; k < 3; k++) c[k] = (alb[k] * irr[k]) * 0.48f;
        aces(c, c);
        for (int k = 0; k < 3; k++) c[k] = pow045(c[k]);
        if (debug_channel == 1) {
            memcpy(c, alb, 12);
        } else if (debug_channel == 2) {
            aces(irr, c);
        } else if (debug_channel == 3) {
            float nn[3];
            unpack_normal(a, nn);
            for (int k = 0; k < 3; k++) c[k] = nn[k] * 0.5f + 0.5f;
        } else if (debug_channel == 4) {
            float it = irr[3];
            if (it < 64.0f) {
                c[0] = c[1] = c[2] = it / 64.0f;
            } else {
                float t = (it - 64.0f) / 128.0f;
                c[0] = 1.0f * (1.0f - t) + 1.0f * t;
                c[1] = c[2] = 1.0f * (1.0f - t) + 0.0f * t;
            }
        } else if (debug_channel == 5) {
            c[0] = c[1] = c[2] = sqrtf(irr[3]) * 3.0f;
        }
        rgba[i] = unorm8(c[0]) | unorm8(c[1]) << 8 | unorm8(c[2]) << 16 | 0xFF000000u;
    }
}

/* One frame: CpuRenderer::RenderFrame's tail (CpuRenderer.cpp:466-473) = blit + GBuffer::DenoiseAndPresent
 * (GBuffer.h:86-130).  `tiles` = w*h*16 bytes in Framebuffer::Tile layout (4x4 tiles).  Call
 * post_oracle_set_camera first (RenderFrame does, :423). */
PO_API void post_oracle_frame(PostOracle* o, const uint32_t* tiles, int reset_history, int num_passes, int debug_channel,
                              uint32_t* out_rgba8) {
    blit_tiles(o, tiles);
    if (debug_channel != 4) {
        reproject_pass(o, reset_history);
        if (num_passes > 0) {
            variance_pass(o);
            for (int i = 0; i < num_passes; i++) {
                Half4* in = i == 1 ? o->prev_irr : (i % 2 == 0 ? o->temp_irr : o->irr);
                Half4* out = i % 2 == 0 ? o->irr : o->temp_irr;
                atrous_pass(o, in, out, i);
                if (i == 0) SWAP(Half4*, o->prev_irr, o->irr);
            }
            if (num_passes % 2 != 0) SWAP(Half4*, o->temp_irr, o->irr);
        }
    }
    present_pass(o, debug_channel, out_rgba8);
    if (num_passes == 0) SWAP(Half4*, o->prev_irr, o->irr);
}

/* Inspection: copy one texture out.  which: 0 irradiance (w*h*4 f16), 1 prev irradiance, 2 temp irradiance,
 * 3 moments (w*h*2 f16), 4 history length (w*h u8), 5 depth (f32), 6 albedo+normal (u32). */
PO_API int post_oracle_read(const PostOracle* o, int which, void* out) {
    size_t n = (size_t)o->w * o->h;
    switch (which) {
    case 0: memcpy(out, o->irr, n * 8); return 0;
    case 1: memcpy(out, o->prev_irr, n * 8); return 0;
    case 2: memcpy(out, o->temp_irr, n * 8); return 0;
    case 3: memcpy(out, o->moments, n * 4); return 0;
    case 4: memcpy(out, o->hist, n); return 0;
    case 5: memcpy(out, o->depth, n * 4); return 0;
    case 6: memcpy(out, o->albedo, n * 4); return 0;
    }
    return -1;
}
PO_API uint16_t post_f32_to_f16(float f) { return f32_to_f16(f); }
PO_API float post_f16_to_f32(uint16_t h) { return f16_to_f32(h); }
