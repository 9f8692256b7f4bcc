/* Surface voxeliser for the bundled-model scenes (INPUT generation for BASELINE config 3; not the
 * product path, not the oracle).
 *
 * Follows the recipe of the reference's VoxelMap::VoxelizeModel (src/VoxelRT/Voxelize.cpp:5-147):
 * conservative triangle/voxel overlap after Schwarz & Seidel 2010 ("Fast parallel surface and solid
 * voxelization on GPUs", the test of §4.1: plane slab between the two critical points + three 2-D
 * edge-function projections), texture colour taken at the barycentric projection of the voxel's
 * corner onto the triangle (Voxelize.cpp:60-75,133-141), nearest texel of a pre-reduced mip, alpha
 * test, palette index written as the voxel id.  Written from the paper's formulation; the layout
 * (row rejection by the yz projection before walking x, sparse 8^3 brick pool) is our own.
 *
 * Textures arrive already quantised: one byte per texel = palette index, 255 = transparent.
 */
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

typedef struct {
    float x, y;
} v2;
static float maxf(float a, float b) { return a > b ? a : b; }

/* one 2-D projection: three inward edge normals + offsets so that n.p + d >= 0 <=> the unit square at p overlaps */
typedef struct {
    v2 n[3];
    float d[3];
} EdgeSet;
static void edge_set(EdgeSet* E, const float a[3][2], float flip) {
    for (int i = 0; i < 3; i++) {
        const float* p = a[i];
        const float* q = a[(i + 1) % 3];
        float ex = q[0] - p[0], ey = q[1] - p[1];
        E->n[i].x = -ey * flip;
        E->n[i].y = ex * flip;
        E->d[i] = -(E->n[i].x * p[0] + E->n[i].y * p[1]) + maxf(0.0f, E->n[i].x) + maxf(0.0f, E->n[i].y);
    }
}
static int edge_pass(const EdgeSet* E, float px, float py) {
    return E->n[0].x * px + E->n[0].y * py + E->d[0] >= 0.0f && E->n[1].x * px + E->n[1].y * py + E->d[1] >= 0.0f &&
           E->n[2].x * px + E->n[2].y * py + E->d[2] >= 0.0f;
}

/* tris: T x 3 x 3 voxel-space positions; uvs: T x 3 x 2; tex: per-triangle texture id (or -1);
 * tex_idx[t]: tex_w[t] x tex_h[t] palette indices (255 = transparent), repeat addressing. */
int64_t vox_triangles(VoxGrid* g, int64_t T, const float* tris, const float* uvs, const int32_t* tex, const uint8_t* const* tex_idx,
                      const int32_t* tex_w, const int32_t* tex_h, uint8_t untextured_id) {
    int64_t degenerate = 0;
    for (int64_t t = 0; t < T; t++) {
        const float* v0 = tris + 9 * t;
        const float* v1 = v0 + 3;
        const float* v2p = v0 + 6;
        float e0[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]};
        float e1[3] = {v2p[0] - v1[0], v2p[1] - v1[1], v2p[2] - v1[2]};
        float n[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
        float nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
        if (!(nn > 0.0f) || !isfinite(nn)) {
            degenerate++;
            continue;
        }
        float inv = 1.0f / sqrtf(nn);
        float nu[3] = {n[0] * inv, n[1] * inv, n[2] * inv};
        /* critical point c = (n > 0), slab offsets d1 = n.(c - v0), d2 = n.((1-c) - v0) */
        float c[3] = {nu[0] > 0.0f ? 1.0f : 0.0f, nu[1] > 0.0f ? 1.0f : 0.0f, nu[2] > 0.0f ? 1.0f : 0.0f};
        float d1 = nu[0] * (c[0] - v0[0]) + nu[1] * (c[1] - v0[1]) + nu[2] * (c[2] - v0[2]);
        float d2 = nu[0] * ((1.0f - c[0]) - v0[0]) + nu[1] * ((1.0f - c[1]) - v0[1]) + nu[2] * ((1.0f - c[2]) - v0[2]);
        const float* V[3] = {v0, v1, v2p};
        float axy[3][2], azx[3][2], ayz[3][2];
        for (int i = 0; i < 3; i++) {
            axy[i][0] = V[i][0], axy[i][1] = V[i][1];
            azx[i][0] = V[i][2], azx[i][1] = V[i][0];
            ayz[i][0] = V[i][1], ayz[i][1] = V[i][2];
        }
        EdgeSet Exy, Ezx, Eyz;
        edge_set(&Exy, axy, nu[2] < 0.0f ? -1.0f : 1.0f);
        edge_set(&Ezx, azx, nu[1] < 0.0f ? -1.0f : 1.0f);
        edge_set(&Eyz, ayz, nu[0] < 0.0f ? -1.0f : 1.0f);

        int lo[3], hi[3];
        for (int a = 0; a < 3; a++) {
            float mn = fminf(fminf(v0[a], v1[a]), v2p[a]), mx = fmaxf(fmaxf(v0[a], v1[a]), v2p[a]);
            lo[a] = (int)mn; /* (ivec3) conversion truncates, Voxelize.cpp:32-33 */
            hi[a] = (int)mx;
        }
        /* barycentric projection set-up (Voxelize.cpp:60-75): u = v1 - v0, v = v2 - v0, n' = u x v */
        float bu[3] = {e0[0], e0[1], e0[2]};
        float bv[3] = {v2p[0] - v0[0], v2p[1] - v0[1], v2p[2] - v0[2]};
        float bn[3] = {bu[1] * bv[2] - bu[2] * bv[1], bu[2] * bv[0] - bu[0] * bv[2], bu[0] * bv[1] - bu[1] * bv[0]};
        float bnn = bn[0] * bn[0] + bn[1] * bn[1] + bn[2] * bn[2];
        const float* uv = uvs + 6 * t;
        int tid = tex[t];
        const uint8_t* timg = tid >= 0 ? tex_idx[tid] : NULL;
        int tw = tid >= 0 ? tex_w[tid] : 1, th = tid >= 0 ? tex_h[tid] : 1;

        for (int y = lo[1]; y <= hi[1]; y++) {
            for (int z = lo[2]; z <= hi[2]; z++) {
                if (!edge_pass(&Eyz, (float)y, (float)z)) continue; /* the row cannot overlap for any x */
                for (int x = lo[0]; x <= hi[0]; x++) {
                    float px = (float)x, py = (float)y, pz = (float)z;
                    float ndp = nu[0] * px + nu[1] * py + nu[2] * pz;
                    if ((ndp + d1) * (ndp + d2) > 0.0f) continue;
                    if (!edge_pass(&Exy, px, py) || !edge_pass(&Ezx, pz, px)) continue;
                    uint8_t id = untextured_id;
                    if (timg) {
                        float w[3] = {px - v0[0], py - v0[1], pz - v0[2]};
                        float uxw[3] = {bu[1] * w[2] - bu[2] * w[1], bu[2] * w[0] - bu[0] * w[2], bu[0] * w[1] - bu[1] * w[0]};
                        float wxv[3] = {w[1] * bv[2] - w[2] * bv[1], w[2] * bv[0] - w[0] * bv[2], w[0] * bv[1] - w[1] * bv[0]};
                        float gamma = (uxw[0] * bn[0] + uxw[1] * bn[1] + uxw[2] * bn[2]) / bnn;
                        float beta = (wxv[0] * bn[0] + wxv[1] * bn[1] + wxv[2] * bn[2]) / bnn;
                        float alpha = 1.0f - gamma - beta;
                        float tu = uv[0] * alpha + uv[2] * beta + uv[4] * gamma;
                        float tv = uv[1] * alpha + uv[3] * beta + uv[5] * gamma;
                        tu -= floorf(tu); /* repeat */
                        tv -= floorf(tv);
                        int ix = (int)(tu * (float)tw), iy = (int)(tv * (float)th);
                        if (ix >= tw) ix = tw - 1;
                        if (iy >= th) iy = th - 1;
                        if (ix < 0) ix = 0;
                        if (iy < 0) iy = 0;
                        id = timg[(size_t)iy * tw + ix];
                        if (id == 255) continue; /* alpha test */
                    }
                    if (id == 0) continue; /* id 0 is the empty voxel (VoxelMap.h:9-21) */
                    vox_set(g, x, y, z, id);
                }
            }
        }
    }
    return degenerate;
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
