/* Surface voxeliser for the bundled-model scenes (INPUT generation for BASELINE config 3; not the
 * product path, not the oracle).
 *
 * Follows the recipe of the reference's VoxelMap::VoxelizeModel (src/VoxelRT/Voxelize.cpp:5-147):
 * conservative triangle/voxel overlap after Schwarz & Seidel 2010 ("Fast parallel surface and solid
 * voxelization on GPUs", the test of §4.1: plane slab between the two critical points + three 2-D
 * edge-function projections), texture colour taken at the barycentric projection of the voxel's
 * corner onto the triangle (Voxelize.cpp:60-75,133-141), the bilinear level-0 tap the reference's sampler
 * takes there, alpha test, nearest palette entry written as the voxel id.  Written from the paper's
 * formulation; the layout (row rejection by the yz projection before walking x, sparse 8^3 brick pool,
 * the colour -> palette index memo) is our own.  The float operations follow the reference's order one by
 * one so that tests/test_ref_pin.py can demand the SAME voxels from the reference's own code.
 */
#include <immintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t bx, by, bz;     /* extent in bricks */
    int32_t* table;         /* brick coordinate -> pool slot, -1 = absent */
    uint8_t* pool;          /* slot * 512 voxel ids, x | z<<3 | y<<6 (VoxelMap.h:100-102) */
    int32_t* coords;        /* slot * 3 brick coordinates */
    int64_t n, cap;
    int64_t voxels_set;
} VoxGrid;

VoxGrid* vox_create(int32_t bx, int32_t by, int32_t bz) {
    VoxGrid* g = (VoxGrid*)calloc(1, sizeof(VoxGrid));
    if (!g) return NULL;
    g->bx = bx, g->by = by, g->bz = bz;
    size_t nb = (size_t)bx * by * bz;
    g->table = (int32_t*)malloc(nb * sizeof(int32_t));
    if (!g->table) {
        free(g);
        return NULL;
    }
    memset(g->table, 0xFF, nb * sizeof(int32_t));
    g->cap = 1 << 14;
    g->pool = (uint8_t*)calloc((size_t)g->cap, 512);
    g->coords = (int32_t*)malloc((size_t)g->cap * 3 * sizeof(int32_t));
    return g;
}

void vox_destroy(VoxGrid* g) {
    if (!g) return;
    free(g->table);
    free(g->pool);
    free(g->coords);
    free(g);
}

int64_t vox_brick_count(const VoxGrid* g) { return g->n; }
int64_t vox_voxels_set(const VoxGrid* g) { return g->voxels_set; }
void vox_export(const VoxGrid* g, uint8_t* bricks, int32_t* coords) {
    memcpy(bricks, g->pool, (size_t)g->n * 512);
    memcpy(coords, g->coords, (size_t)g->n * 3 * sizeof(int32_t));
}

static void vox_set(VoxGrid* g, int x, int y, int z, uint8_t id) {
    if (x < 0 || y < 0 || z < 0) return;
    int bx = x >> 3, by = y >> 3, bz = z >> 3;
    if (bx >= g->bx || by >= g->by || bz >= g->bz) return;
    size_t ti = (size_t)bx + (size_t)bz * g->bx + (size_t)by * g->bx * g->bz;
    int32_t slot = g->table[ti];
    if (slot < 0) {
        if (g->n == g->cap) {
            int64_t nc = g->cap * 2;
            uint8_t* np = (uint8_t*)realloc(g->pool, (size_t)nc * 512);
            int32_t* nq = (int32_t*)realloc(g->coords, (size_t)nc * 3 * sizeof(int32_t));
            if (!np || !nq) abort();
            memset(np + (size_t)g->cap * 512, 0, (size_t)(nc - g->cap) * 512);
            g->pool = np, g->coords = nq, g->cap = nc;
        }
        slot = (int32_t)g->n++;
        g->table[ti] = slot;
        g->coords[3 * slot] = bx, g->coords[3 * slot + 1] = by, g->coords[3 * slot + 2] = bz;
    }
    g->pool[(size_t)slot * 512 + (x & 7) + ((z & 7) << 3) + ((y & 7) << 6)] = id;
    g->voxels_set++;
}

/* ---- arithmetic of the reference's voxeliser, operation by operation (this file is compiled with -ffp-contract=off; the reference harness
 * build it is pinned against, oracle/_ref/libref_cpu_strict.so, too) ---- */
typedef struct {
    float x, y;
} v2;
static inline float gmax(float a, float b) { return (a < b) ? b : a; } /* glm::max */
static inline float gmin(float a, float b) { return (b < a) ? b : a; } /* glm::min */
static inline float dot3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static inline void cross3(float r[3], const float x[3], const float y[3]) { /* glm::cross */
    r[0] = x[1] * y[2] - y[1] * x[2];
    r[1] = x[2] * y[0] - y[2] * x[0];
    r[2] = x[0] * y[1] - y[0] * x[1];
}
static inline int32_t cvt_rn(float x) { return _mm_cvtss_si32(_mm_set_ss(x)); }  /* simd::round2i: nearest-even, NaN / overflow -> INT_MIN */
static inline int32_t cvt_tr(float x) { return _mm_cvttss_si32(_mm_set_ss(x)); } /* (ivec3) cast */

/* one 2-D projection: three edge normals + offsets, n.p + d >= 0 <=> the unit square at p overlaps (Voxelize.cpp:18-28) */
typedef struct {
    v2 n[3];
    float d[3];
} EdgeSet;
static void edge_set(EdgeSet* E, float nx0, float ny0, float nx1, float ny1, float nx2, float ny2, const float* pa, const float* pb, float flip) {
    const float nx[3] = {nx0, nx1, nx2}, ny[3] = {ny0, ny1, ny2};
    for (int i = 0; i < 3; i++) {
        E->n[i].x = nx[i] * flip;
        E->n[i].y = ny[i] * flip;
        E->d[i] = (-(E->n[i].x * pa[i] + E->n[i].y * pb[i]) + gmax(0.0f, E->n[i].x)) + gmax(0.0f, E->n[i].y);
    }
}
static inline int edge_pass(const EdgeSet* E, float px, float py) {
    return (E->n[0].x * px + E->n[0].y * py) + E->d[0] >= 0.0f && (E->n[1].x * px + E->n[1].y * py) + E->d[1] >= 0.0f &&
           (E->n[2].x * px + E->n[2].y * py) + E->d[2] >= 0.0f;
}

/* _mm_mulhrs_epi16 on one 16-bit lane, and simd::lerp16 (SIMD_AVX512.h:152) */
static inline int32_t lerp16(int32_t a, int32_t b, int32_t t) {
    int16_t d = (int16_t)(b - a);
    int16_t m = (int16_t)((((int32_t)d * (int32_t)(int16_t)t >> 14) + 1) >> 1);
    return (int32_t)(int16_t)(a + m);
}
/* Texture2D<RGBA8u>::Sample<{Repeat, Mag Linear, Min Nearest, mips}> as VoxelizeModel's voxel pass reaches it (Voxelize.cpp:139): all 16
 * lanes carry the same (u, v), so the UV derivatives are zero, the computed level is <= 0 whatever the `2` argument says (Texture.h:504-509)
 * and the sample is the MAG filter's bilinear tap of level 0 (Texture.h:580-615): 8 fractional bits, half-texel shift of 127/256, the
 * +1 neighbours clamped at the right / bottom edge, two channels at a time through 15-bit mulhrs lerps. */
static uint32_t sample_mag(const uint32_t* img, int32_t w, int32_t h, float u, float v) {
    const float su = u * (float)(w << 8), sv = v * (float)(h << 8);
    int32_t ix = cvt_rn(su) & ((w << 8) - 1), iy = cvt_rn(sv) & ((h << 8) - 1);
    ix = ix - 127 > 0 ? ix - 127 : 0;
    iy = iy - 127 > 0 ? iy - 127 : 0;
    const int32_t tx = ix >> 8, ty = iy >> 8;
    const int32_t fx = (ix & 255) << 7, fy = (iy & 255) << 7;
    const size_t i00 = (size_t)tx + (size_t)ty * (size_t)w;
    const size_t i10 = i00 + (tx + 1 < w ? 1 : 0);
    const size_t row = ty + 1 < h ? (size_t)w : 0;
    const uint32_t c00 = img[i00], c10 = img[i10], c01 = img[i00 + row], c11 = img[i10 + row];
    uint32_t out = 0;
    for (int ch = 0; ch < 4; ch++) {
        const int sh = ch * 8;
        const int32_t r1 = lerp16((c00 >> sh) & 255, (c10 >> sh) & 255, fx);
        const int32_t r2 = lerp16((c01 >> sh) & 255, (c11 >> sh) & 255, fx);
        out |= ((uint32_t)lerp16(r1, r2, fy) & 0xFFFFu) << sh; /* a 16-bit lane; bits above 7 would spill into the next channel like `rb | ga << 8` does */
    }
    return out;
}

/* PaletteBuilder::FindIndex (PaletteBuilder.h:66-128): smallest Manhattan distance, first entry wins */
static uint8_t find_index(const uint8_t* pal, int32_t n_pal, uint32_t color) {
    const int32_t r = color & 255, g = (color >> 8) & 255, b = (color >> 16) & 255;
    int32_t best = 32767, idx = 0;
    for (int32_t i = 0; i < n_pal; i++) {
        const int32_t d = abs(r - pal[3 * i]) + abs(g - pal[3 * i + 1]) + abs(b - pal[3 * i + 2]);
        if (d < best) best = d, idx = i;
    }
    return (uint8_t)idx;
}

/* VoxelMap::VoxelizeModel's triangle loop (Voxelize.cpp:117-146) over T triangles already placed in voxel space.
 * tris: T x 3 x 3; uvs: T x 3 x 2; tex: per-triangle texture id (or -1: `untextured_id` is written, the reference would sample an
 * uninitialised 4x4 placeholder there, Scene.cpp:82-85); tex_rgba[t]: tex_w[t] x tex_h[t] RGBA8 texels of level 0, R in the low byte;
 * pal: n_pal x 3 palette colours.  Triangles are taken in order and later writes win, like the reference's Set() calls; a voxel whose
 * colour maps to palette entry 0 is written as id 0, i.e. emptied, like the reference does (entry 0 is the empty voxel, VoxelMap.h:9-21).
 * Zero-area triangles are not skipped either: their NaN normal passes the slab test (Voxelize.cpp:41 compares `> 0`).
 * Returns the number of zero-area triangles seen. */
int64_t vox_triangles(VoxGrid* g, int64_t T, const float* tris, const float* uvs, const int32_t* tex, const uint32_t* const* tex_rgba,
                      const int32_t* tex_w, const int32_t* tex_h, const uint8_t* pal, int32_t n_pal, uint8_t untextured_id) {
    int64_t degenerate = 0;
    uint8_t* lut = (uint8_t*)malloc(1u << 24); /* colour -> palette index memo, 255 = not looked up yet (n_pal <= 240) */
    if (!lut) return -1;
    memset(lut, 0xFF, 1u << 24);
    for (int64_t t = 0; t < T; t++) {
        const float* v0 = tris + 9 * t;
        const float* v1 = v0 + 3;
        const float* v2p = v0 + 6;
        const float e[3][3] = {{v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]},
                               {v2p[0] - v1[0], v2p[1] - v1[1], v2p[2] - v1[2]},
                               {v0[0] - v2p[0], v0[1] - v2p[1], v0[2] - v2p[2]}};
        float n[3];
        cross3(n, e[0], e[1]);
        const float nn = dot3(n, n);
        if (!(nn > 0.0f) || !isfinite(nn)) degenerate++;
        const float inv = 1.0f / sqrtf(nn); /* glm::normalize */
        const float nu[3] = {n[0] * inv, n[1] * inv, n[2] * inv};
        /* critical point c = max(sign(n), 0), slab offsets d1 = n.(c - v0), d2 = n.((1 - c) - v0) */
        const float c[3] = {0.0f < nu[0] ? 1.0f : 0.0f, 0.0f < nu[1] ? 1.0f : 0.0f, 0.0f < nu[2] ? 1.0f : 0.0f};
        const float cv[3] = {c[0] - v0[0], c[1] - v0[1], c[2] - v0[2]};
        const float dv[3] = {(1.0f - c[0]) - v0[0], (1.0f - c[1]) - v0[1], (1.0f - c[2]) - v0[2]};
        const float d1 = dot3(nu, cv), d2 = dot3(nu, dv);
        const float vx[3] = {v0[0], v1[0], v2p[0]}, vy[3] = {v0[1], v1[1], v2p[1]}, vz[3] = {v0[2], v1[2], v2p[2]};
        EdgeSet Exy, Ezx, Eyz;
        edge_set(&Exy, -e[0][1], e[0][0], -e[1][1], e[1][0], -e[2][1], e[2][0], vx, vy, nu[2] < 0.0f ? -1.0f : 1.0f);
        edge_set(&Ezx, -e[0][0], e[0][2], -e[1][0], e[1][2], -e[2][0], e[2][2], vz, vx, nu[1] < 0.0f ? -1.0f : 1.0f);
        edge_set(&Eyz, -e[0][2], e[0][1], -e[1][2], e[1][1], -e[2][2], e[2][1], vy, vz, nu[0] < 0.0f ? -1.0f : 1.0f);

        int lo[3], hi[3];
        for (int a = 0; a < 3; a++) {
            lo[a] = cvt_tr(gmin(gmin(v0[a], v1[a]), v2p[a])); /* (ivec3) conversion truncates, Voxelize.cpp:30-31 */
            hi[a] = cvt_tr(gmax(gmax(v0[a], v1[a]), v2p[a]));
        }
        /* barycentric projection set-up (Voxelize.cpp:60-75): u = v1 - v0, v = v2 - v0, n' = u x v */
        const float bu[3] = {e[0][0], e[0][1], e[0][2]};
        const float bv[3] = {v2p[0] - v0[0], v2p[1] - v0[1], v2p[2] - v0[2]};
        float bn[3];
        cross3(bn, bu, bv);
        const float bnn = dot3(bn, bn);
        const float* uv = uvs + 6 * t;
        const int tid = tex[t];
        const uint32_t* timg = tid >= 0 ? tex_rgba[tid] : NULL;
        const int32_t tw = tid >= 0 ? tex_w[tid] : 1, th = tid >= 0 ? tex_h[tid] : 1;

        for (int y = lo[1]; y <= hi[1]; y++) {
            for (int z = lo[2]; z <= hi[2]; z++) {
                if (!edge_pass(&Eyz, (float)y, (float)z)) continue; /* the row cannot overlap for any x */
                for (int x = lo[0]; x <= hi[0]; x++) {
                    const float p[3] = {(float)x, (float)y, (float)z};
                    const float ndp = dot3(nu, p);
                    if ((ndp + d1) * (ndp + d2) > 0.0f) continue;
                    if (!edge_pass(&Exy, p[0], p[1]) || !edge_pass(&Ezx, p[2], p[0])) continue;
                    uint8_t id = untextured_id;
                    if (timg) {
                        const float w[3] = {p[0] - v0[0], p[1] - v0[1], p[2] - v0[2]};
                        float uxw[3], wxv[3];
                        cross3(uxw, bu, w);
                        cross3(wxv, w, bv);
                        const float gamma = dot3(uxw, bn) / bnn;
                        const float beta = dot3(wxv, bn) / bnn;
                        const float alpha = 1.0f - gamma - beta;
                        const float tu = (uv[0] * alpha + uv[2] * beta) + uv[4] * gamma;
                        const float tv = (uv[1] * alpha + uv[3] * beta) + uv[5] * gamma;
                        const uint32_t color = sample_mag(timg, tw, th, tu, tv);
                        if (color < 0x80000000u) continue; /* alpha test, Voxelize.cpp:140 */
                        id = lut[color & 0xFFFFFFu];
                        if (id == 255) id = lut[color & 0xFFFFFFu] = find_index(pal, n_pal, color);
                    }
                    vox_set(g, x, y, z, id);
                }
            }
        }
    }
    free(lut);
    return degenerate;
}

/* Texture2D::GenerateMip (Texture.h:683-704): each level is the 2x2 average of the one below, computed in fp32 on [0,1] values
 * ((c00 + c10 + c01 + c11) * 0.25, unpacked with * (1/255) and packed back with round-to-nearest-even of * 255).
 * src: w x h RGBA8, dst: (w/2) x (h/2). */
void vox_mip_half(const uint32_t* src, int32_t w, int32_t h, uint32_t* dst) {
    const float scale = 1.0f / 255;
    const int32_t w2 = w >> 1, h2 = h >> 1;
    for (int32_t y = 0; y < h2; y++)
        for (int32_t x = 0; x < w2; x++) {
            const uint32_t c00 = src[(size_t)(2 * y) * w + 2 * x], c10 = src[(size_t)(2 * y) * w + 2 * x + 1];
            const uint32_t c01 = src[(size_t)(2 * y + 1) * w + 2 * x], c11 = src[(size_t)(2 * y + 1) * w + 2 * x + 1];
            uint32_t out = 0;
            for (int ch = 0; ch < 4; ch++) {
                const int sh = ch * 8;
                const float a = (float)((c00 >> sh) & 255) * scale, b = (float)((c10 >> sh) & 255) * scale;
                const float c = (float)((c01 >> sh) & 255) * scale, d = (float)((c11 >> sh) & 255) * scale;
                const float avg = (((a + b) + c) + d) * 0.25f;
                int32_t q = cvt_rn(avg * 255.0f);
                q = q < 0 ? 0 : (q > 255 ? 255 : q); /* packs / packus saturation */
                out |= (uint32_t)q << sh;
            }
            dst[(size_t)y * w2 + x] = out;
        }
}

/* Radiance RGBE (.hdr) scanline decoder for scenes/convert_assets.py (new-style RLE, the only form the bundled skyboxes use).
 * data: the bytes after the resolution line; out: h * w * 4 bytes (R, G, B, E).  Returns 0, or -1 on malformed input. */
int rgbe_decode(const uint8_t* data, int64_t n, int32_t w, int32_t h, uint8_t* out) {
    int64_t p = 0;
    for (int32_t y = 0; y < h; y++) {
        uint8_t* row = out + (size_t)y * w * 4;
        if (p + 4 > n) return -1;
        if (data[p] != 2 || data[p + 1] != 2 || ((data[p + 2] << 8) | data[p + 3]) != w) return -1; /* flat / old RLE not needed */
        p += 4;
        for (int c = 0; c < 4; c++) {
            int32_t x = 0;
            while (x < w) {
                if (p >= n) return -1;
                int count = data[p++];
                if (count > 128) { /* run */
                    count -= 128;
                    if (p >= n || x + count > w) return -1;
                    uint8_t v = data[p++];
                    for (int i = 0; i < count; i++) row[(size_t)(x++) * 4 + c] = v;
                } else { /* literal */
                    if (count == 0 || p + count > n || x + count > w) return -1;
                    for (int i = 0; i < count; i++) row[(size_t)(x++) * 4 + c] = data[p++];
                }
            }
        }
    }
    return 0;
}
