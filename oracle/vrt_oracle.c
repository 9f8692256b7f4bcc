/*
 * vrt_oracle.c — CPU ORACLE (test infrastructure, never linked into the product).
 *
 * Scalar lane-wise restatement of the reference CPU renderer.  Every function cites the
 * reference lines it follows (paths relative to /root/reference/src).  Arithmetic recipe
 * ("canonical arithmetic"): IEEE-754 binary32, round-to-nearest, FMA exactly
 * where written as fmaf(), x86 min/cvt semantics spelled out.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math -mfma  (oracle/Makefile)
 * so the compiler neither fuses nor splits anything.
 */
#include "vrt_oracle.h"
#include "x86_approx14_tables.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* Storage: FlatVoxelStorage (VoxelRT/CpuRenderer.cpp:20-31) with the dense 2 GiB buffer       */
/* replaced by per-sector blocks allocated on first use.  Addressing semantics are the dense   */
/* ones: an unallocated brick reads as zeros, indices wrap by masking.                         */
/* ------------------------------------------------------------------------------------------ */
struct OrcMap {
    uint32_t shift_xz, shift_y; /* ViewSectorIndexer<ShiftXZ,ShiftY> (CpuRenderer.cpp:14-16) */
    uint32_t n_sectors;
    uint64_t* sector_masks;  /* SectorMasks[]                                   */
    uint8_t** sector_voxels; /* StorageBuffer slice of the sector, 64*512 B     */
    uint64_t** sector_cells; /* OccupancyStorage slice of the sector, 64*8 u64  */
    uint64_t palette[256];
    uint8_t* blue_noise; /* 128 x 8192 x (R,G) */
    VrtSkyDesc sky;
    uint32_t* sky_texels;
};

static inline uint32_t sector_index(const OrcMap* m, int32_t sx, int32_t sy, int32_t sz) {
    /* LinearIndexer3D::GetIndex (VoxelRT/VoxelMap.h:94-97): x | z << sXZ | y << 2 sXZ, masked */
    uint32_t mxz = (1u << m->shift_xz) - 1, my = (1u << m->shift_y) - 1;
    return ((uint32_t)sx & mxz) | (((uint32_t)sz & mxz) << m->shift_xz) | (((uint32_t)sy & my) << (2 * m->shift_xz));
}

OrcMap* orc_map_create(uint32_t sectors_xz_log2, uint32_t sectors_y_log2) {
    OrcMap* m = (OrcMap*)calloc(1, sizeof(OrcMap));
    m->shift_xz = sectors_xz_log2;
    m->shift_y = sectors_y_log2;
    m->n_sectors = 1u << (2 * sectors_xz_log2 + sectors_y_log2);
    m->sector_masks = (uint64_t*)calloc(m->n_sectors, sizeof(uint64_t));
    m->sector_voxels = (uint8_t**)calloc(m->n_sectors, sizeof(uint8_t*));
    m->sector_cells = (uint64_t**)calloc(m->n_sectors, sizeof(uint64_t*));
    return m;
}

void orc_map_destroy(OrcMap* m) {
    if (!m) return;
    for (uint32_t i = 0; i < m->n_sectors; i++) {
        free(m->sector_voxels[i]);
        free(m->sector_cells[i]);
    }
    free(m->sector_masks);
    free(m->sector_voxels);
    free(m->sector_cells);
    free(m->blue_noise);
    free(m->sky_texels);
    free(m);
}

void orc_map_set_palette(OrcMap* m, const uint64_t palette[256]) { memcpy(m->palette, palette, sizeof(m->palette)); }

/* FlatVoxelStorage::UpdateOccupancy, VoxelRT/CpuRenderer.cpp:63-83.
 * cell (cx,cy,cz) -> BrickMaskIndexer index cx | cz<<1 | cy<<2 (CpuRenderer.cpp:18,80);
 * bit vx + 4 vz + 16 vy (CpuRenderer.cpp:78); voxel index x | z<<3 | y<<6 (VoxelMap.h:102). */
void orc_build_occupancy(const uint8_t brick[512], uint64_t cells[8]) {
    for (uint32_t cy = 0; cy < 8; cy += 4)
        for (uint32_t cz = 0; cz < 8; cz += 4)
            for (uint32_t cx = 0; cx < 8; cx += 4) {
                uint64_t mask = 0;
                for (uint32_t vy = 0; vy < 4; vy++)
                    for (uint32_t vz = 0; vz < 4; vz++)
                        for (uint32_t vx = 0; vx < 4; vx++) {
                            uint32_t idx = (cx + vx) | ((cz + vz) << 3) | ((cy + vy) << 6);
                            uint64_t occupied = brick[idx] != 0;
                            mask |= occupied << (vx + vz * 4 + vy * 16);
                        }
                cells[(cx >> 2) | ((cz >> 2) << 1) | ((cy >> 2) << 2)] = mask;
            }
}

/* FlatVoxelStorage::SyncBuffers, VoxelRT/CpuRenderer.cpp:38-60 */
int orc_map_sync(OrcMap* m, uint32_t n, const VrtDirtySector* recs) {
    for (uint32_t r = 0; r < n; r++) {
        const VrtDirtySector* d = &recs[r];
        /* ViewSectorIndexer::CheckInBounds (VoxelMap.h:73-76) */
        if (((uint32_t)(d->sx | d->sz) >> m->shift_xz) != 0 || ((uint32_t)d->sy >> m->shift_y) != 0) continue;
        uint32_t si = sector_index(m, d->sx, d->sy, d->sz);
        if (d->flags & VRT_SECTOR_REMOVED) { /* CpuRenderer.cpp:43-46 */
            m->sector_masks[si] = 0;
            /* deviation (documented): the reference leaves stale voxel bytes behind; they are
             * only observable through the wrapped material fetch of out-of-grid rays (Q4). */
            if (m->sector_voxels[si]) memset(m->sector_voxels[si], 0, 64 * 512);
            if (m->sector_cells[si]) memset(m->sector_cells[si], 0, 64 * 8 * sizeof(uint64_t));
            continue;
        }
        uint64_t old_mask = m->sector_masks[si];
        m->sector_masks[si] = d->alloc_mask; /* CpuRenderer.cpp:49-50 */
        if (!m->sector_voxels[si]) {
            m->sector_voxels[si] = (uint8_t*)calloc(64, 512);
            m->sector_cells[si] = (uint64_t*)calloc(64 * 8, sizeof(uint64_t));
        }
        /* bricks that left the allocation mask: zero them (same deviation as above) */
        uint64_t gone = old_mask & ~d->alloc_mask;
        for (; gone; gone &= gone - 1) {
            uint32_t b = (uint32_t)__builtin_ctzll(gone);
            memset(m->sector_voxels[si] + b * 512, 0, 512);
            memset(m->sector_cells[si] + b * 8, 0, 64);
        }
        const uint8_t* src = d->bricks;
        uint64_t todo = d->dirty_mask & d->alloc_mask; /* CpuRenderer.cpp:52 */
        for (; todo; todo &= todo - 1) {
            uint32_t b = (uint32_t)__builtin_ctzll(todo);
            memcpy(m->sector_voxels[si] + b * 512, src, 512);            /* :56 */
            orc_build_occupancy(src, m->sector_cells[si] + b * 8);       /* :57 */
            src += 512;
        }
    }
    return 0;
}

int orc_map_read_sector(const OrcMap* m, int32_t sx, int32_t sy, int32_t sz, uint64_t* alloc_mask, uint8_t* bricks,
                        uint64_t* cells) {
    if (((uint32_t)(sx | sz) >> m->shift_xz) != 0 || ((uint32_t)sy >> m->shift_y) != 0) return -1;
    uint32_t si = sector_index(m, sx, sy, sz);
    if (alloc_mask) *alloc_mask = m->sector_masks[si];
    if (bricks) {
        if (m->sector_voxels[si]) memcpy(bricks, m->sector_voxels[si], 64 * 512);
        else memset(bricks, 0, 64 * 512);
    }
    if (cells) {
        if (m->sector_cells[si]) memcpy(cells, m->sector_cells[si], 64 * 8 * sizeof(uint64_t));
        else memset(cells, 0, 64 * 8 * sizeof(uint64_t));
    }
    return 0;
}

void orc_set_blue_noise(OrcMap* m, const uint8_t* rg, size_t bytes) {
    free(m->blue_noise);
    m->blue_noise = (uint8_t*)malloc(VRT_BLUE_NOISE_BYTES);
    memset(m->blue_noise, 0, VRT_BLUE_NOISE_BYTES);
    memcpy(m->blue_noise, rg, bytes < VRT_BLUE_NOISE_BYTES ? bytes : VRT_BLUE_NOISE_BYTES);
}

void orc_set_sky(OrcMap* m, const VrtSkyDesc* desc, const uint32_t* texels) {
    free(m->sky_texels);
    m->sky = *desc;
    m->sky_texels = (uint32_t*)malloc(desc->texel_count * sizeof(uint32_t));
    memcpy(m->sky_texels, texels, desc->texel_count * sizeof(uint32_t));
}

/* ------------------------------------------------------------------------------------------ */
/* x86 semantics the reference inherits from its intrinsics                                    */
/* ------------------------------------------------------------------------------------------ */
/* _mm512_min_ps(a,b): a < b ? a : b — returns b when either is NaN (LibGlimpsw/SwRast/SIMD_AVX512.h:123) */
static inline float x86_min(float a, float b) { return a < b ? a : b; }
static inline float x86_max(float a, float b) { return a > b ? a : b; }
/* _mm512_cvt_roundps_epi32(x, TO_NEG_INF) (SIMD_AVX512.h:110): NaN / out of range -> 0x80000000 */
static inline int32_t x86_floor2i(float x) {
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return INT32_MIN;
    return (int32_t)floorf(x);
}
/* _mm512_cvtps_epi32 (SIMD_AVX512.h:108): round half even, indefinite on NaN / overflow */
static inline int32_t x86_round2i(float x) {
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return INT32_MIN;
    return (int32_t)nearbyintf(x); /* default rounding mode = RNE */
}
static inline int32_t x86_trunc2i(float x) {
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return INT32_MIN;
    return (int32_t)x;
}
static inline float x86_fract(float x); /* below */
static inline uint32_t f2u(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
static inline float u2f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
/* simd::fract = VREDUCEPS imm8=TO_NEG_INF (SIMD_AVX512.h:116): x - floor(x) with the result itself
 * rounded toward -inf (pinned against oracle/_ref: a negative x whose 1 + x is inexact comes out
 * one ulp below the round-to-nearest value).  x - floor(x) is exact in binary64. */
static inline float x86_fract(float x) {
    if (!(x == x)) return x;     /* NaN propagates */
    if (isinf(x)) return 0.0f;   /* VREDUCEPS(+-inf) = +0 */
    double e = (double)x - floor((double)x);
    if (e == 0.0) return -0.0f; /* an exact zero difference is -0 under round-toward-negative */
    float r = (float)e;
    if ((double)r > e) r = nextafterf(r, -INFINITY);
    return r;
}
/* approx_rsqrt = _mm512_rsqrt14_ps, approx_rcp = _mm512_rcp14_ps (SIMD_AVX512.h:136-138), bit-exact: both instructions are
 * architecturally defined; measured on the instruction, the result is a 64-segment piecewise-linear function of the top
 * mantissa bits in integer arithmetic (x86_approx14_tables.h), exact for powers of two (four).  orc_x86_rcp14 / orc_x86_rsqrt14
 * are compared with the instructions on ALL 2^32 inputs by tests/test_x86_approx14.py (needs an AVX-512 host). */
float orc_x86_rsqrt14(float x) {
    uint32_t u = f2u(x), sign = u & 0x80000000u, a = u & 0x7FFFFFFFu;
    if (a > 0x7F800000u) return u2f(u | 0x00400000u); /* NaN -> quiet, payload kept */
    if (a == 0) return u2f(sign | 0x7F800000u);       /* +-0 -> +-inf */
    if (sign) return u2f(0xFFC00000u);                /* negative -> real indefinite */
    if (a == 0x7F800000u) return 0.0f;
    int e = (int)(a >> 23);
    uint32_t mant = a & 0x7FFFFFu;
    if (e == 0) { /* denormal argument */
        int sh = __builtin_clz(mant) - 8;
        mant = (mant << sh) & 0x7FFFFFu;
        e = 1 - sh;
    }
    int E = e - 127, odd = E & 1, k = (E - odd) / 2; /* x = (1.m * 2^odd) * 4^k */
    uint32_t rb = 0x3F800000u;
    if (mant != 0 || odd) {
        const uint32_t* c = orc_rsqrt14_coef[((uint32_t)odd << 5) | (mant >> 18)];
        rb = ((c[0] - c[1] * ((mant >> 8) & 0x3FFu)) >> 9) << 7;
    }
    return u2f(rb - ((uint32_t)k << 23));
}
float orc_x86_rcp14(float x) {
    uint32_t u = f2u(x), sign = u & 0x80000000u, a = u & 0x7FFFFFFFu;
    if (a > 0x7F800000u) return u2f(u | 0x00400000u);
    if (a == 0x7F800000u) return u2f(sign);
    if (a == 0) return u2f(sign | 0x7F800000u);
    int e = (int)(a >> 23);
    uint32_t mant = a & 0x7FFFFFu;
    if (e == 0) {
        int sh = __builtin_clz(mant) - 8;
        mant = (mant << sh) & 0x7FFFFFu;
        e = 1 - sh;
    }
    uint32_t rb = 0x3F800000u;
    if (mant != 0) {
        const uint32_t* c = orc_rcp14_coef[mant >> 17];
        rb = ((c[0] - c[1] * ((mant >> 7) & 0x3FFu)) >> 9) << 7;
    }
    int re = (int)(rb >> 23) - (e - 127);
    uint32_t rm = rb & 0x7FFFFFu;
    if (re >= 255) return u2f(sign | 0x7F800000u);
    if (re <= 0) { /* denormal result, truncated */
        int sh = 1 - re;
        return u2f(sh > 24 ? sign : (sign | ((rm | 0x800000u) >> sh)));
    }
    return u2f(sign | ((uint32_t)re << 23) | rm);
}
static inline float approx_rsqrt(float x) { return orc_x86_rsqrt14(x); }
static inline float approx_rcp(float x) { return orc_x86_rcp14(x); }

/* f16 <-> f32, IEEE RNE (Texture.h:101-116 uses vcvtps2ph/vcvtph2ps); NaN canonicalised */
static uint16_t f32_to_f16(float f) {
    uint32_t x = f2u(f);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7FFFFFFFu;
    if (ax > 0x7F800000u) return (uint16_t)(sign | 0x7E00u | ((x >> 13) & 0x1FFu)); /* NaN: vcvtps2ph keeps sign + upper payload, quiets */
    if (ax >= 0x47800000u) {                                /* >= 65536 -> inf unless rounds below */
        return (uint16_t)(sign | 0x7C00u);
    }
    if (ax >= 0x38800000u) { /* normal half */
        uint32_t mant = ax & 0x7FFFFFu;
        uint32_t exp = (ax >> 23) - 112;
        uint32_t h = (exp << 10) | (mant >> 13);
        uint32_t rem = mant & 0x1FFFu;
        if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) h++;
        return (uint16_t)(sign | h); /* carry into exponent (up to inf) is correct */
    }
    if (ax < 0x33000000u) return (uint16_t)sign; /* < 2^-25 -> 0 */
    /* subnormal half */
    uint32_t mant = (ax & 0x7FFFFFu) | 0x800000u;
    uint32_t shift = 126 - (ax >> 23); /* 14..24 */
    uint32_t h = mant >> shift;
    uint32_t rem = mant & ((1u << shift) - 1);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) h++;
    return (uint16_t)(sign | h);
}
static float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 31, mant = h & 0x3FFu;
    if (exp == 31) return u2f(sign | 0x7F800000u | (mant << 13));
    if (exp == 0) {
        if (mant == 0) return u2f(sign);
        float v = (float)mant * (1.0f / 16777216.0f); /* mant * 2^-24, exact */
        return sign ? -v : v;
    }
    return u2f(sign | ((exp + 112) << 23) | (mant << 13));
}

/* Material::GetEncoded, VoxelRT/VoxelMap.h:27-41 */
uint64_t orc_encode_material(uint8_t r, uint8_t g, uint8_t b, uint8_t fuzz, float emission) {
    uint64_t packed = 0;
    packed |= (uint64_t)(r >> 3) << 11;
    packed |= (uint64_t)(g >> 2) << 5;
    packed |= (uint64_t)(b >> 3) << 0;
    packed |= (uint64_t)f32_to_f16(emission) << 16; /* packHalf2x16(vec2(0, Emission)) */
    packed |= (uint64_t)fuzz << 32;
    return packed;
}

/* ------------------------------------------------------------------------------------------ */
/* Traversal                                                                                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t iters, sector_fetches, cell_fetches;
    uint64_t lod[6]; /* steps taken per cell size 1,2,4,8,16,32 */
} CastCounters;

/* GetInboundMask, VoxelRT/CpuRenderer.cpp:114-117 */
static inline int inbound(const OrcMap* m, int32_t x, int32_t y, int32_t z) {
    return (uint32_t)(x | z) < (1u << (m->shift_xz + 5)) && (uint32_t)y < (1u << (m->shift_y + 5));
}

/* GetVoxelMaterial, VoxelRT/CpuRenderer.cpp:120-132 — every index is masked, so any position
 * reads some in-buffer voxel (the "wrapped address" of quirk Q4). */
static inline uint32_t voxel_material(const OrcMap* m, int32_t x, int32_t y, int32_t z) {
    uint32_t si = sector_index(m, x >> 5, y >> 5, z >> 5);
    uint32_t bi = ((uint32_t)(x >> 3) & 3) | (((uint32_t)(z >> 3) & 3) << 2) | (((uint32_t)(y >> 3) & 3) << 4);
    uint32_t vi = ((uint32_t)x & 7) | (((uint32_t)z & 7) << 3) | (((uint32_t)y & 7) << 6);
    uint8_t id = m->sector_voxels[si] ? m->sector_voxels[si][bi * 512 + vi] : 0;
    return (uint32_t)m->palette[id];
}

/* GetStepPos, VoxelRT/CpuRenderer.cpp:135-171.  Returns the hit flag; p is advanced to the far
 * corner of the empty cell it is in (unchanged on hit). */
static inline int step_pos(const OrcMap* m, int32_t p[3], const float d[3], CastCounters* c) {
    uint32_t si = sector_index(m, p[0] >> 5, p[1] >> 5, p[2] >> 5); /* :136 */
    uint64_t mask = m->sector_masks[si];                            /* :138-139 */
    c->sector_fetches++;
    uint32_t idx = ((uint32_t)(p[0] >> 3) & 3) | (((uint32_t)(p[2] >> 3) & 3) << 2) | (((uint32_t)(p[1] >> 3) & 3) << 4); /* :141 */
    int level0 = (int)((mask >> idx) & 1); /* :142-143 */
    int lod = 3;                           /* :144 */
    if (level0) {                          /* :146-158 */
        uint32_t cell = ((uint32_t)(p[0] >> 2) & 1) | (((uint32_t)(p[2] >> 2) & 1) << 1) | (((uint32_t)(p[1] >> 2) & 1) << 2);
        mask = m->sector_cells[si] ? m->sector_cells[si][idx * 8 + cell] : 0; /* :147-148,153-154 */
        c->cell_fetches++;
        idx = ((uint32_t)p[0] & 3) | (((uint32_t)p[2] & 3) << 2) | (((uint32_t)p[1] & 3) << 4); /* :150 */
        lod = 0;                                                                               /* :151 */
        level0 = (int)((mask >> idx) & 1);                                                     /* :156-157 */
    }
    uint32_t half = idx < 32 ? (uint32_t)mask : (uint32_t)(mask >> 32);
    int level4 = mask == 0;                                          /* :160 */
    int level2 = ((half >> (idx & 0xA)) & 0x00330033u) == 0;         /* :161 */
    lod += level4 ? 2 : (level2 ? 1 : 0);                            /* :162 */
    if (!level0) c->lod[lod]++;
    int32_t cm = (1 << lod) - 1;                                     /* :164 */
    for (int a = 0; a < 3; a++) p[a] = d[a] < 0 ? (p[a] & ~cm) : (p[a] | cm); /* :166-168 */
    return level0;
}

/* RayCast, VoxelRT/CpuRenderer.cpp:172-224, one lane. */
static void cast_ray_ex(const OrcMap* m, const float o[3], const float d[3], const int32_t wo[3], uint32_t max_iters,
                        VrtHit* out, CastCounters* cnt, uint32_t* trips_out);
static void cast_ray(const OrcMap* m, const float o[3], const float d[3], const int32_t wo[3], uint32_t max_iters,
                     VrtHit* out, CastCounters* cnt) {
    cast_ray_ex(m, o, d, wo, max_iters, out, cnt, NULL);
}
/* trips_out: loop trips started before the lane stopped (1 = stopped in the first trip); max_iters for a capped lane */
static void cast_ray_ex(const OrcMap* m, const float o[3], const float d[3], const int32_t wo[3], uint32_t max_iters,
                        VrtHit* out, CastCounters* cnt, uint32_t* trips_out) {
    float inv[3], ts[3], sd[3] = {0, 0, 0}, cur[3];
    int32_t p[3] = {0, 0, 0};
    for (int a = 0; a < 3; a++) {
        inv[a] = 1.0f / d[a];                               /* :173 */
        ts[a] = ((d[a] < 0 ? 0.0f : 1.0f) - o[a]) * inv[a]; /* :175-179 */
        cur[a] = o[a];                                      /* :181 */
    }
    int hit = 0, inb = 0, active = 1;
    uint32_t it = 0;
    for (; it < max_iters; it++) { /* :185 */
        for (int a = 0; a < 3; a++) p[a] = (int32_t)((uint32_t)wo[a] + (uint32_t)x86_floor2i(cur[a])); /* :186 */
        cnt->iters++;
        inb = inbound(m, p[0], p[1], p[2]); /* :188 */
        if (!inb) {                         /* :189 */
            active = 0;
            break;
        }
        hit = step_pos(m, p, d, cnt); /* :190 */
        if (hit) {                    /* :192-193 */
            active = 0;
            break;
        }
        for (int a = 0; a < 3; a++) /* :195-198, contracted to FMA by the reference's compilers */
            sd[a] = fmaf((float)(int32_t)((uint32_t)p[a] - (uint32_t)wo[a]), inv[a], ts[a]);
        float tmin = x86_min(x86_min(sd[0], sd[1]), sd[2]) + 0.001f; /* :200 */
        for (int a = 0; a < 3; a++) cur[a] = fmaf(tmin, d[a], o[a]);  /* :201 */
    }
    float hd = x86_min(x86_min(sd[0], sd[1]), sd[2]); /* :204 */
    int mx = sd[0] == hd, my = sd[1] == hd;           /* :205-206 */
    int mz = !mx && !my;                              /* :207 */
    /* :214-216  csel(side, (dir & -0.0f) ^ -1.0f, 0): sign BIT of dir set -> +1 else -1 */
    int nx = mx ? ((f2u(d[0]) >> 31) ? 1 : -1) : 0;
    int ny = my ? ((f2u(d[1]) >> 31) ? 1 : -1) : 0;
    int nz = mz ? ((f2u(d[2]) >> 31) ? 1 : -1) : 0;
    float fu = mx ? cur[1] : cur[0], fv = mz ? cur[1] : cur[2]; /* :218-221 */
    out->vx = p[0];
    out->vy = p[1];
    out->vz = p[2];
    /* :210  material fetched for every non-active lane; lanes still active (cap) read 0 */
    out->material = active ? 0u : voxel_material(m, p[0], p[1], p[2]);
    out->dist = hd;
    out->px = cur[0];
    out->py = cur[1];
    out->pz = cur[2];
    out->u = x86_fract(fu); /* fract = _mm512_reduce_ps(x, TO_NEG_INF) (SIMD_AVX512.h:116) */
    out->v = x86_fract(fv);
    uint32_t iters_done = it < max_iters ? it + 1 : max_iters;
    if (trips_out) *trips_out = iters_done;
    if (iters_done > 0xFFFF) iters_done = 0xFFFF;
    out->flags = (uint32_t)((nx + 1) | ((ny + 1) << 2) | ((nz + 1) << 4)) | ((!active && inb) ? VRT_HIT_HIT : 0) | /* :222 */
                 (inb ? VRT_HIT_INBOUND : 0) | (active ? VRT_HIT_CAPPED : 0) | (iters_done << VRT_HIT_ITERS_SHIFT);
    out->_pad = 0;
}

static void stats_add(OrcStats* s, const CastCounters* c, const VrtHit* h) {
    s->rays++;
    s->iters += c->iters;
    s->sector_fetches += c->sector_fetches;
    s->cell_fetches += c->cell_fetches;
    for (int i = 0; i < 6; i++) s->lod_hist[i] += c->lod[i];
    if (h->flags & VRT_HIT_HIT) s->hits++;
    if (h->flags & VRT_HIT_CAPPED) s->capped++;
    uint32_t it = h->flags >> VRT_HIT_ITERS_SHIFT;
    int b = it < 4 ? 0 : it < 8 ? 1 : it < 16 ? 2 : it < 32 ? 3 : it < 64 ? 4 : it < 128 ? 5 : it < 256 ? 6 : 7;
    s->iter_hist[b]++;
}
static void stats_merge(OrcStats* dst, const OrcStats* src) {
    dst->rays += src->rays;
    dst->iters += src->iters;
    dst->sector_fetches += src->sector_fetches;
    dst->cell_fetches += src->cell_fetches;
    dst->hits += src->hits;
    dst->capped += src->capped;
    for (int i = 0; i < 8; i++) dst->iter_hist[i] += src->iter_hist[i];
    for (int i = 0; i < 6; i++) dst->lod_hist[i] += src->lod_hist[i];
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_trace(const OrcMap* m, uint64_t n, const float* origin3, const float* dir3, const int32_t wo[3],
               uint32_t max_iters, VrtHit* out, OrcStats* stats, int threads) {
    if (max_iters == 0) max_iters = VRT_MAX_ITERS_DEFAULT;
    OrcStats total;
    memset(&total, 0, sizeof(total));
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
#endif
    {
        OrcStats local;
        memset(&local, 0, sizeof(local));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4096)
#endif
        for (int64_t i = 0; i < (int64_t)n; i++) {
            CastCounters c = {0, 0, 0, {0}};
            cast_ray(m, origin3 + 3 * i, dir3 + 3 * i, wo, max_iters, &out[i], &c);
            stats_add(&local, &c, &out[i]);
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        stats_merge(&total, &local);
    }
    if (stats) *stats = total;
}

/* ------------------------------------------------------------------------------------------ */
/* Hit query: VoxelMap::RayCast + GetStepLevel, VoxelRT/VoxelMap.cpp:125-170 (fp64)            */
/* ------------------------------------------------------------------------------------------ */
static inline double glm_min(double x, double y) { return (y < x) ? y : x; } /* glm::min */

/* GetStepLevel, VoxelMap.cpp:125-139.  The reference looks the sector up in the unbounded
 * hash; the resident view stands in for it (sectors outside the view = absent). */
static int step_level(const OrcMap* m, int32_t x, int32_t y, int32_t z) {
    int32_t sx = x >> 5, sy = y >> 5, sz = z >> 5;
    if (((uint32_t)(sx | sz) >> m->shift_xz) != 0 || ((uint32_t)sy >> m->shift_y) != 0) return 5;
    uint32_t si = sector_index(m, sx, sy, sz);
    uint64_t mask = m->sector_masks[si];
    if (mask == 0 && !m->sector_voxels[si]) return 5; /* sector not in the map            */
    /* NOTE: a sector present with an all-zero mask also steps 32 here; the reference would
     * step 8 brick by brick through it.  vrt_sync never reports such a sector as present
     * (the adapter erases empty sectors like RegionDispatchSIMD does, VoxelMap.h:254-262). */
    if (mask == 0) return 5;
    uint32_t bi = ((uint32_t)(x >> 3) & 3) | (((uint32_t)(z >> 3) & 3) << 2) | (((uint32_t)(y >> 3) & 3) << 4);
    if (!((mask >> bi) & 1)) return 3;
    uint32_t vi = ((uint32_t)x & 7) | (((uint32_t)z & 7) << 3) | (((uint32_t)y & 7) << 6);
    return m->sector_voxels[si][bi * 512 + vi] == 0 ? 0 : -1;
}

static void hit_query_one(const OrcMap* m, const double o[3], const double d[3], uint32_t max_iters, VrtHitD* out) {
    double inv[3], ts[3];
    int32_t pos[3];
    for (int a = 0; a < 3; a++) {
        inv[a] = 1.0 / d[a];                                   /* :141 */
        ts[a] = ((d[a] < 0.0 ? 0.0 : 1.0) - o[a]) * inv[a];    /* :142 step(0,dir) */
        pos[a] = (int32_t)floor(o[a]);                         /* :143 */
    }
    memset(out, 0, sizeof(*out));
    out->dist = -1.0;
    for (uint32_t i = 0; i < max_iters; i++) {
        double sd[3], hp[3];
        for (int a = 0; a < 3; a++) sd[a] = fma((double)pos[a], inv[a], ts[a]); /* :146 */
        double tmin = glm_min(glm_min(sd[0], sd[1]), sd[2]) + 0.0001;           /* :147 */
        for (int a = 0; a < 3; a++) {
            hp[a] = fma(tmin, d[a], o[a]);                                      /* :148 */
            double f = floor(hp[a]);
            pos[a] = (f >= -2147483648.0 && f < 2147483648.0) ? (int32_t)f : INT32_MIN; /* :149 */
        }
        int k = step_level(m, pos[0], pos[1], pos[2]); /* :151 */
        if (k < 0) {                                   /* :153-162 */
            int smx = tmin >= sd[0], smy = tmin >= sd[1], smz = tmin >= sd[2];
            out->dist = tmin;
            out->nx = smx ? (float)-((d[0] > 0.0) - (d[0] < 0.0)) : 0.0f;
            out->ny = smy ? (float)-((d[1] > 0.0) - (d[1] < 0.0)) : 0.0f;
            out->nz = smz ? (float)-((d[2] > 0.0) - (d[2] < 0.0)) : 0.0f;
            float fu = smx ? (float)hp[1] : (float)hp[0];
            float fv = smz ? (float)hp[1] : (float)hp[2];
            out->u = fu - floorf(fu);
            out->v = fv - floorf(fv);
            out->vx = pos[0];
            out->vy = pos[1];
            out->vz = pos[2];
            out->iters = i + 1;
            return;
        }
        int32_t mk = (1 << k) - 1; /* :164-167 */
        for (int a = 0; a < 3; a++) pos[a] = d[a] < 0.0 ? (pos[a] & ~mk) : (pos[a] | mk);
    }
    out->iters = max_iters;
}

void orc_hit_query(const OrcMap* m, uint64_t n, const double* origin3, const double* dir3, uint32_t max_iters,
                   VrtHitD* out, int threads) {
    if (max_iters == 0) max_iters = 1024;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 256)
#endif
    for (int64_t i = 0; i < (int64_t)n; i++) hit_query_one(m, origin3 + 3 * i, dir3 + 3 * i, max_iters, &out[i]);
}

/* ------------------------------------------------------------------------------------------ */
/* Ray generation and shading                                                                  */
/* ------------------------------------------------------------------------------------------ */
/* simd::TransformVector, LibGlimpsw/SwRast/SIMD.h:207-214; m is column-major (glm::mat4) */
static inline void transform_vec4(const float m[16], const float v[4], float r[4]) {
    for (int row = 0; row < 4; row++)
        r[row] = fmaf(m[0 + row], v[0], fmaf(m[4 + row], v[1], fmaf(m[8 + row], v[2], m[12 + row] * v[3])));
}
/* simd::normalize, SIMD.h:109-115: a * approx_rsqrt(dot(a, a)) */
static inline void normalize3(float a[3]) {
    float len = approx_rsqrt(fmaf(a[0], a[0], fmaf(a[1], a[1], a[2] * a[2])));
    a[0] *= len;
    a[1] *= len;
    a[2] *= len;
}

/* GetPrimaryRay (CpuRenderer.cpp:226-233) with u = x + 0.5, v = y + 0.5 (:327,330) and
 * origin += OriginFrac (:334).  The half-pixel offset is applied a second time inside
 * inv_proj (GBuffer.h:137) — quirk Q9, kept. */
void orc_primary_ray(const VrtFrame* f, uint32_t x, uint32_t y, float origin[3], float dir[3]) {
    float uv[4] = {(float)(int32_t)x + 0.5f, (float)(int32_t)y + 0.5f, 0.0f, 1.0f};
    float nearp[4], farp[4];
    transform_vec4(f->inv_proj, uv, nearp);                             /* :227 */
    for (int a = 0; a < 4; a++) farp[a] = nearp[a] + f->inv_proj[8 + a]; /* :228 */
    float rn = 1.0f / nearp[3], rf = 1.0f / farp[3];
    for (int a = 0; a < 3; a++) {
        origin[a] = nearp[a] * rn; /* :229 */
        dir[a] = farp[a] * rf;     /* :230 */
    }
    normalize3(dir); /* :232 */
    for (int a = 0; a < 3; a++) origin[a] = origin[a] + f->origin_frac[a];
}

/* simd::sincos_2pi (AVX-512 branch), SIMD.h:175-190 */
static inline void sincos_2pi(float x, float* s, float* c) {
    float t = x + 0.25f;
    float xr = t - nearbyintf(t); /* _mm512_reduce_ps(t, NEAREST) */
    float x1 = fabsf(xr) - 0.25f;
    float x2 = x1 * x1;
    *s = x1 * fmaf(x2, -36.26749369f, 6.23786927f);
    float cc = fmaf(x2, fmaf(x2, 57.34151006f, -19.56474772f), 0.99940322f);
    *c = u2f(f2u(cc) | (f2u(xr) & 0x80000000u));
}

/* SampleDirection, CpuRenderer.cpp:273-291.  approx_sqrt(v) = rsqrt14(v) * v (SIMD_AVX512.h:136);
 * v = 0 gives inf * 0 = NaN (quirk Q7). */
void orc_sample_direction(float sx, float sy, float out[3]) {
    float y = fmaf(sy, 2.0f, -1.0f); /* :284 (exact either way) */
    float x, z;
    sincos_2pi(sx, &x, &z);          /* :287 */
    float v = fmaf(-y, y, 1.0f);     /* :288, 1 - y*y contracted */
    float s = approx_rsqrt(v) * v;
    out[0] = x * s;
    out[1] = y;
    out[2] = z * s;
}

/* dir = normalize(hit.Normal + SampleDirection(bn)), CpuRenderer.cpp:392.  SampleDirection is inlined into RenderRow, and the reference's
 * compilers contract `Normal.x + x * sy` / `Normal.z + z * sy` into FMAs (the same -ffp-contract=fast that fuses sideDist and
 * currPos in RayCast; found by a 1-ulp direction difference that turned a stalled ray of the reference into a hit).  The y
 * component has no product: plain add. */
static inline void bounce_direction(const float nrm[3], float sx, float sy, float dir[3]) {
    float y = fmaf(sy, 2.0f, -1.0f);
    float x, z;
    sincos_2pi(sx, &x, &z);
    float v = fmaf(-y, y, 1.0f);
    float s = approx_rsqrt(v) * v;
    dir[0] = fmaf(x, s, nrm[0]);
    dir[1] = nrm[1] + y;
    dir[2] = fmaf(z, s, nrm[2]);
    normalize3(dir);
}

/* VBlueNoise::Sample, CpuRenderer.cpp:254-270, for the 16-lane (4x4 tile) build: the R2 offset
 * is applied to the tile origin and snapped to the tile grid; the lane keeps its in-tile
 * position (quirk Q6). */
void orc_blue_noise_sample(const OrcMap* m, uint32_t x, uint32_t y, uint32_t frame_no, uint32_t sample_idx, float out[2]) {
    float fi = (float)sample_idx;
    float ox = fi * 0.75487766624669276005f + 0.5f, oy = fi * 0.56984029099805326591f + 0.5f; /* :258 */
    ox = ox - floorf(ox);
    oy = oy - floorf(oy);
    uint32_t px = ((x & ~3u) + (uint32_t)(ox * 128.0f)) & 127u; /* :259 */
    uint32_t py = ((y & ~3u) + (uint32_t)(oy * 128.0f)) & 127u;
    py += (frame_no & 63u) * 128u;                              /* :260 */
    uint32_t tx = (px & ~3u) + (x & 3u), ty = (py & ~3u) + (y & 3u); /* :262-263 + ctor :242-250 */
    const uint8_t* t = m->blue_noise + ((size_t)ty * 128 + tx) * 2;
    out[0] = (float)t[0] * (float)(1.0 / 255); /* :269 */
    out[1] = (float)t[1] * (float)(1.0 / 255);
}

/* R11G11B10f, LibGlimpsw/SwRast/Texture.h:150-205 */
static inline float unpack_f11(uint32_t x) { return u2f(((x << 17) & 0x0FFE0000u) + 0x38000000u); }
static inline float unpack_f10(uint32_t x) { return u2f(((x << 18) & 0x0FFC0000u) + 0x38000000u); }
uint32_t orc_pack_r11g11b10f(float r, float g, float b) {
    r = x86_min(x86_max(r, 1.0f / (1 << 15)), 130048.0f);
    g = x86_min(x86_max(g, 1.0f / (1 << 15)), 130048.0f);
    b = x86_min(x86_max(b, 1.0f / (1 << 15)), 129024.0f);
    uint32_t pr = (((uint32_t)((int32_t)f2u(r) >> 17)) & 0x3FFFu) - 0x1C00u;
    uint32_t pg = (((uint32_t)((int32_t)f2u(g) >> 17)) & 0x3FFFu) - 0x1C00u;
    uint32_t pb = (((uint32_t)((int32_t)f2u(b) >> 18)) & 0x1FFFu) - 0x0E00u;
    return (pr << 21) | (pg << 10) | pb;
}

/* texutil::ProjectCubemap (Texture.h:264-288) + Texture2D::Sample<Nearest,…,IsCube>
 * (Texture.h:487-545) at an integer mip + the x3 of CpuRenderer.cpp:355-356. */
void orc_sky_sample(const OrcMap* m, const float dir[3], uint32_t mip, float out[3]) {
    if (!m->sky_texels) {
        out[0] = out[1] = out[2] = 0.0f;
        return;
    }
    /* VFloat operator> is _mm512_cmp_ps_mask(a, b, _MM_CMPINT_GT) (SIMD_AVX512.h:83): the INTEGER predicate 6 read as a float predicate is
     * _CMP_NLE_US, "not less-or-equal" — TRUE when either operand is NaN.  A NaN direction (quirk Q7) therefore selects face 4 / 5. */
    float w = dir[0];
    int wy = !(fabsf(dir[1]) <= fabsf(w));
    w = wy ? dir[1] : w;
    int wz = !(fabsf(dir[2]) <= fabsf(w));
    w = wz ? dir[2] : w;
    int wx = wy | wz;
    wy &= !wz;
    uint32_t face = wz ? 4u : (wy ? 2u : 0u);
    face += f2u(w) >> 31;
    w = approx_rcp(fabsf(w)) * 0.5f; /* Texture.h:285 */
    float u = fmaf(wx ? dir[0] : dir[2], w, 0.5f);
    float v = fmaf(wy ? dir[2] : dir[1], w, 0.5f);

    const VrtSkyDesc* s = &m->sky;
    int32_t mask_lerp = (int32_t)(s->face_size << 8) - 1;
    float scale_lerp = (float)(mask_lerp + 1);
    int32_t ix = x86_round2i(u * scale_lerp), iy = x86_round2i(v * scale_lerp);
    ix = ix < 0 ? 0 : (ix > mask_lerp ? mask_lerp : ix); /* cube sample clamps (Texture.h:493-495) */
    iy = iy < 0 ? 0 : (iy > mask_lerp ? mask_lerp : iy);
    uint32_t mlev = mip < s->mip_levels ? mip : s->mip_levels - 1; /* :506 */
    uint32_t row_shift = (uint32_t)__builtin_ctz(s->face_size);
    uint32_t stride = row_shift - mlev;
    uint32_t off = (face << s->layer_shift) + s->mip_offset[mlev];
    ix = (ix >> mlev) >> 8;
    iy = (iy >> mlev) >> 8;
    uint32_t texel = m->sky_texels[off + (uint32_t)ix + ((uint32_t)iy << stride)];
    out[0] = unpack_f11(texel >> 21) * 3.0f;
    out[1] = unpack_f11(texel >> 10) * 3.0f;
    out[2] = unpack_f10(texel) * 3.0f;
}

/* RGBA8u::Pack, Texture.h:41-62: round2i(v*255), packs (i32->i16 signed sat), packus (->u8) */
static inline uint32_t pack_unorm8(float v) {
    int32_t i = x86_round2i(v * 255.0f);
    if (i < -32768) i = -32768;
    if (i > 32767) i = 32767;
    return (uint32_t)(i < 0 ? 0 : (i > 255 ? 255 : i));
}

uint32_t orc_pack_unorm8x4(float r, float g, float b, float a) {
    return pack_unorm8(r) | (pack_unorm8(g) << 8) | (pack_unorm8(b) << 16) | (pack_unorm8(a) << 24);
}
uint32_t orc_pack_half2(float x, float y) { return (uint32_t)f32_to_f16(x) | ((uint32_t)f32_to_f16(y) << 16); }
void orc_sincos_2pi(float x, float* s, float* c) { sincos_2pi(x, s, c); }

/* RenderRow body for one pixel, CpuRenderer.cpp:326-402, LANE-WISE (a lane whose mask bit is off does nothing further): only the
 * debugging aid orc_debug_pixel uses it; frames are rendered by render_tile below, which carries the packet semantics. */
/* optional capture of every ray a pixel casts (debugging aid for the parity tests) */
typedef struct {
    float* rays; /* per bounce: origin xyz, dir xyz */
    VrtHit* hits;
    uint32_t n;
} PixelTrace;

static void render_pixel_ex(const OrcMap* m, const VrtFrame* f, uint32_t x, uint32_t y, uint32_t* o_albedo, float* o_depth,
                            uint32_t* o_rg, uint32_t* o_bx, VrtHit* aux, OrcStats* stats, PixelTrace* pt);
uint32_t orc_debug_pixel(const OrcMap* m, const VrtFrame* f, uint32_t x, uint32_t y, float* rays6, VrtHit* hits, uint32_t out4[4]) {
    PixelTrace pt = {rays6, hits, 0};
    float dep;
    render_pixel_ex(m, f, x, y, &out4[0], &dep, &out4[2], &out4[3], NULL, NULL, &pt);
    memcpy(&out4[1], &dep, 4);
    return pt.n;
}
static void render_pixel_ex(const OrcMap* m, const VrtFrame* f, uint32_t x, uint32_t y, uint32_t* o_albedo, float* o_depth,
                            uint32_t* o_rg, uint32_t* o_bx, VrtHit* aux, OrcStats* stats, PixelTrace* pt) {
    float origin[3], dir[3];
    orc_primary_ray(f, x, y, origin, dir); /* :327-334 */
    uint32_t max_iters = f->max_iters ? f->max_iters : VRT_MAX_ITERS_DEFAULT;

    uint32_t albedo = 0;
    float depth = 0.0f;
    float irr[3] = {0, 0, 0}, thr[3] = {1, 1, 1}; /* :338-339 */

    for (uint32_t i = 0; i <= f->bounces; i++) { /* :342 */
        VrtHit hit;
        CastCounters c = {0, 0, 0, {0}};
        cast_ray(m, origin, dir, f->world_origin, max_iters, &hit, &c); /* :343 */
        if (pt) {
            memcpy(pt->rays + 6 * pt->n, origin, 12);
            memcpy(pt->rays + 6 * pt->n + 3, dir, 12);
            pt->hits[pt->n++] = hit;
        }
        if (stats) stats_add(stats, &c, &hit);
        if (i == 0 && aux) *aux = hit;

        uint32_t md = hit.material;
        /* VHitResult::GetColor / GetEmissionStrength, :97-107 */
        float col[3] = {(float)((md >> 11) & 31) * (1.0f / 31), (float)((md >> 5) & 63) * (1.0f / 63),
                        (float)(md & 31) * (1.0f / 31)};
        for (int a = 0; a < 3; a++) col[a] = col[a] * col[a];
        float emission = f16_to_f32((uint16_t)(md >> 16));

        int is_hit = (hit.flags & VRT_HIT_HIT) != 0;
        if (!is_hit) { /* :348-369 */
            float sky[3];
            orc_sky_sample(m, dir, i == 0 ? 1 : 3, sky);
            if (i == 0) {
                irr[0] = sky[0];
                irr[1] = sky[1];
                irr[2] = sky[2];
            } else {
                col[0] = sky[0];
                col[1] = sky[1];
                col[2] = sky[2];
                emission = 1.0f;
            }
        }
        int nx = (int)(hit.flags & 3) - 1, ny = (int)((hit.flags >> 2) & 3) - 1, nz = (int)((hit.flags >> 4) & 3) - 1;
        if (i == 0) { /* :370-382 */
            albedo = pack_unorm8(col[0]) | (pack_unorm8(col[1]) << 8) | (pack_unorm8(col[2]) << 16) | (pack_unorm8(0.0f) << 24);
            albedo |= (uint32_t)(nx + 1) << 24;
            albedo |= (uint32_t)(ny + 1) << 26;
            albedo |= (uint32_t)(nz + 1) << 28;
            float pp[4] = {hit.px / 16.0f, hit.py / 16.0f, hit.pz / 16.0f, 1.0f}, pr[4];
            transform_vec4(f->proj, pp, pr);
            depth = !is_hit ? -1.0f : pr[2] / pr[3];
            if (f->bounces == 0) {
                irr[0] = irr[1] = irr[2] = 1.0f;
                break;
            }
        } else {
            for (int a = 0; a < 3; a++) thr[a] = thr[a] * col[a]; /* :384 */
        }
        for (int a = 0; a < 3; a++) irr[a] = fmaf(thr[a], emission, irr[a]); /* :386 */
        if (!is_hit) break;                                                /* :387 mask &= hit.Mask */

        float nrm[3] = {(float)nx, (float)ny, (float)nz};
        origin[0] = fmaf(nrm[0], 0.01f, hit.px); /* :389 */
        origin[1] = fmaf(nrm[1], 0.01f, hit.py);
        origin[2] = fmaf(nrm[2], 0.01f, hit.pz);
        float bn[2];
        orc_blue_noise_sample(m, x, y, f->frame_no, i, bn); /* :391 */
        bounce_direction(nrm, bn[0], bn[1], dir);           /* :392 */
    }
    *o_albedo = albedo;
    *o_depth = depth;
    *o_rg = (uint32_t)f32_to_f16(irr[0]) | ((uint32_t)f32_to_f16(irr[1]) << 16); /* :398 */
    uint32_t hz = f32_to_f16(irr[2]);
    *o_bx = hz | (hz << 16); /* :399 Pack({z}) -> VFloat2(v){x=y=v} (SIMD.h:32) */
}

/* ------------------------------------------------------------------------------------------ */
/* The 16-lane PACKET semantics of RayCast / RenderRow.                                          */
/* The reference traces 4x4-pixel tiles as one SIMD packet (AVX-512: simd::VectorWidth = 16,      */
/* TileWidth = TileHeight = 4), and a few statements of RayCast and RenderRow are not masked, so   */
/* what a lane returns depends on its 15 neighbours (DESIGN.md §3, quirks Q2 / Q10 and the          */
/* "zombie" lanes).  orc_render reproduces all of it; orc_trace stays lane-wise (one ray per packet).*/
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    float pos[3];      /* VHitResult::Pos      */
    uint32_t material; /* MaterialData (low 32 bits of the palette entry) */
    int n[3];          /* Normal               */
    int hit;           /* Mask                 */
} PacketLane;

/* RayCast (CpuRenderer.cpp:172-224) for one packet.  `active` = the activeMask argument.  Per lane, the loop of an ACTIVE lane is the
 * lane-wise cast_ray; the coupling is:
 *  - the loop runs until no lane is active, and `currPos = origin + tmin * dir` (:200-201) is not masked: a lane that stopped in
 *    the FIRST trip, or was never active, still has sideDist = 0, so once the packet goes on past its first trip its currPos
 *    becomes origin + 0.001 * dir, and voxelPos / inboundMask follow (:186-188) (Q10);
 *  - when some lane is still active after the last trip (iteration cap), the loop ends right after `voxelPos -= worldOrigin` (:195),
 *    so GetVoxelMaterial (:210) reads voxel (voxelPos - worldOrigin) for every stopped lane of that packet (Q2);
 *  - MaterialData is gathered for every lane that is not active at the end, never-active lanes included (:210), and
 *    Mask = ~activeMask & inboundMask (:222) can be set for a never-active lane. */
static void cast_packet(const OrcMap* m, float o[16][3], float d[16][3], uint32_t active, const int32_t wo[3], uint32_t max_iters,
                        PacketLane out[16], VrtHit lane_hits[16], OrcStats* stats) {
    uint32_t trips[16];
    int capped[16];
    int pk_multi = 0, pk_capped = 0;
    for (int l = 0; l < 16; l++) {
        trips[l] = 0;
        capped[l] = 0;
        if (!((active >> l) & 1)) continue;
        CastCounters c = {0, 0, 0, {0}};
        cast_ray_ex(m, o[l], d[l], wo, max_iters, &lane_hits[l], &c, &trips[l]);
        if (stats) stats_add(stats, &c, &lane_hits[l]);
        capped[l] = (lane_hits[l].flags & VRT_HIT_CAPPED) != 0;
        if (capped[l]) pk_capped = 1;
        if (capped[l] || trips[l] >= 2) pk_multi = 1; /* still active after the first trip */
    }
    for (int l = 0; l < 16; l++) {
        const int was_active = (int)((active >> l) & 1);
        const int first = !was_active || (trips[l] == 1 && !capped[l]);
        PacketLane* r = &out[l];
        int32_t p[3];
        int inb;
        if (first) {
            float sd0 = 0.0f;
            float tmin = x86_min(x86_min(sd0, sd0), sd0) + 0.001f; /* :200 with sideDist = 0 */
            for (int a = 0; a < 3; a++) {
                r->pos[a] = pk_multi ? fmaf(tmin, d[l][a], o[l][a]) : o[l][a]; /* :201, :181 */
                p[a] = (int32_t)((uint32_t)wo[a] + (uint32_t)x86_floor2i(r->pos[a])); /* :186 */
            }
            inb = inbound(m, p[0], p[1], p[2]); /* :188 */
            /* :204-216 with sideDist = 0: X and Y both match */
            r->n[0] = (f2u(d[l][0]) >> 31) ? 1 : -1;
            r->n[1] = (f2u(d[l][1]) >> 31) ? 1 : -1;
            r->n[2] = 0;
        } else {
            const VrtHit* h = &lane_hits[l];
            r->pos[0] = h->px, r->pos[1] = h->py, r->pos[2] = h->pz;
            p[0] = h->vx, p[1] = h->vy, p[2] = h->vz;
            inb = (h->flags & VRT_HIT_INBOUND) != 0;
            r->n[0] = (int)(h->flags & 3) - 1, r->n[1] = (int)((h->flags >> 2) & 3) - 1, r->n[2] = (int)((h->flags >> 4) & 3) - 1;
        }
        if (capped[l]) { /* still active: not gathered (:210), Mask clear (:222) */
            r->material = 0;
            r->hit = 0;
        } else {
            if (pk_capped)
                for (int a = 0; a < 3; a++) p[a] = (int32_t)((uint32_t)p[a] - (uint32_t)wo[a]); /* :195 without the matching += */
            r->material = voxel_material(m, p[0], p[1], p[2]);
            r->hit = inb;
        }
    }
}

/* RenderRow body for one 4x4 tile = one packet, CpuRenderer.cpp:332-400.  Lane l = pixel (x0 + l % 4, y0 + l / 4) (SIMD.h TileOffsets).
 * Everything after RayCast is unmasked in the reference except the sky lookup (:348-369) and the depth of missed primary rays (:377):
 * lanes whose path has ended keep accumulating `throughput * emission` of whatever voxel their stale position maps to and keep
 * producing new origins / directions for as long as some lane of the packet is alive. */
static void render_tile(const OrcMap* m, const VrtFrame* f, uint32_t x0, uint32_t y0, uint32_t albedo[16], float depth[16], uint32_t irr_rg[16],
                        uint32_t irr_bx[16], VrtHit* aux /* frame-sized, row-major, or NULL */, OrcStats* stats) {
    float origin[16][3], dir[16][3], irr[16][3], thr[16][3];
    const uint32_t max_iters = f->max_iters ? f->max_iters : VRT_MAX_ITERS_DEFAULT;
    for (int l = 0; l < 16; l++) {
        orc_primary_ray(f, x0 + (uint32_t)(l & 3), y0 + (uint32_t)(l >> 2), origin[l], dir[l]); /* :327-334 */
        for (int a = 0; a < 3; a++) irr[l][a] = 0.0f, thr[l][a] = 1.0f;                          /* :338-339 */
        albedo[l] = 0;
        depth[l] = 0.0f;
    }
    uint32_t mask = 0xFFFFu;
    for (uint32_t i = 0; i <= f->bounces && mask; i++) { /* :342 */
        PacketLane hit[16];
        VrtHit lane_hits[16];
        cast_packet(m, origin, dir, mask, f->world_origin, max_iters, hit, lane_hits, stats); /* :343 */
        for (int l = 0; l < 16; l++) {
            const uint32_t x = x0 + (uint32_t)(l & 3), y = y0 + (uint32_t)(l >> 2);
            if (i == 0 && aux) aux[(size_t)y * f->width + x] = lane_hits[l]; /* the lane-wise RayCast record of the primary ray */
            const uint32_t md = hit[l].material;
            float col[3] = {(float)((md >> 11) & 31) * (1.0f / 31), (float)((md >> 5) & 63) * (1.0f / 63), (float)(md & 31) * (1.0f / 31)}; /* :97-104 */
            for (int a = 0; a < 3; a++) col[a] = col[a] * col[a];
            float emission = f16_to_f32((uint16_t)(md >> 16)); /* :105-107 */
            const int miss = (int)((mask >> l) & 1) && !hit[l].hit; /* :346 missMask = mask & ~hit.Mask */
            if (miss) {                                            /* :347-369 */
                float sky[3];
                orc_sky_sample(m, dir[l], i == 0 ? 1 : 3, sky);
                if (i == 0) irr[l][0] = sky[0], irr[l][1] = sky[1], irr[l][2] = sky[2];
                else col[0] = sky[0], col[1] = sky[1], col[2] = sky[2], emission = 1.0f;
            }
            if (i == 0) { /* :370-382 */
                albedo[l] = pack_unorm8(col[0]) | (pack_unorm8(col[1]) << 8) | (pack_unorm8(col[2]) << 16) | (pack_unorm8(0.0f) << 24);
                albedo[l] |= (uint32_t)(hit[l].n[0] + 1) << 24;
                albedo[l] |= (uint32_t)(hit[l].n[1] + 1) << 26;
                albedo[l] |= (uint32_t)(hit[l].n[2] + 1) << 28;
                float pp[4] = {hit[l].pos[0] / 16.0f, hit[l].pos[1] / 16.0f, hit[l].pos[2] / 16.0f, 1.0f}, pr[4];
                transform_vec4(f->proj, pp, pr);
                depth[l] = miss ? -1.0f : pr[2] / pr[3];
                if (f->bounces == 0) { /* :379-382 */
                    irr[l][0] = irr[l][1] = irr[l][2] = 1.0f;
                    continue;
                }
            } else {
                for (int a = 0; a < 3; a++) thr[l][a] = thr[l][a] * col[a]; /* :384 */
            }
            for (int a = 0; a < 3; a++) irr[l][a] = fmaf(thr[l][a], emission, irr[l][a]); /* :386 */
            const float nrm[3] = {(float)hit[l].n[0], (float)hit[l].n[1], (float)hit[l].n[2]};
            for (int a = 0; a < 3; a++) origin[l][a] = fmaf(nrm[a], 0.01f, hit[l].pos[a]); /* :389 */
            float bn[2];
            orc_blue_noise_sample(m, x, y, f->frame_no, i, bn); /* :391 */
            bounce_direction(nrm, bn[0], bn[1], dir[l]);        /* :392 */
        }
        if (f->bounces == 0) break;
        for (int l = 0; l < 16; l++)
            if (!hit[l].hit) mask &= ~(1u << l); /* :387 mask &= hit.Mask */
    }
    for (int l = 0; l < 16; l++) {
        irr_rg[l] = (uint32_t)f32_to_f16(irr[l][0]) | ((uint32_t)f32_to_f16(irr[l][1]) << 16); /* :398 */
        uint32_t hz = f32_to_f16(irr[l][2]);
        irr_bx[l] = hz | (hz << 16); /* :399 Pack({z}) -> VFloat2(v){x=y=v} (SIMD.h:32) */
    }
}

void orc_render(const OrcMap* m, const VrtFrame* f, void* out, VrtHit* aux_hits, OrcStats* stats, int threads,
                uint32_t row0, uint32_t row1) {
    uint32_t w = f->width, h = f->height;
    if (row1 > h) row1 = h;
    OrcStats total;
    memset(&total, 0, sizeof(total));
    uint32_t part_count = f->part_count ? f->part_count : 1;
    uint32_t tiles_x = (w + 31) / 32;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
#endif
    {
        OrcStats local;
        memset(&local, 0, sizeof(local));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int64_t ty = row0 / 4; ty < (int64_t)(row1 / 4); ty++) { /* :455-462 one tile row */
            for (uint32_t x0 = 0; x0 + 4 <= w; x0 += 4) {             /* :329 one packet */
                const uint32_t y0 = (uint32_t)ty * 4;
                if (part_count > 1) { /* a 4x4 tile never straddles two parts (bands are 8 rows, macro tiles 32x32) */
                    uint32_t t = (f->flags & VRT_FRAME_PART_ROWS) ? (y0 / VRT_BAND_ROWS) : (y0 / 32) * tiles_x + (x0 / 32);
                    if (t % part_count != f->part_index) continue;
                }
                uint32_t alb[16], rg[16], bx[16];
                float dep[16];
                render_tile(m, f, x0, y0, alb, dep, rg, bx, aux_hits, stats ? &local : NULL);
                for (uint32_t l = 0; l < 16; l++) {
                    const uint32_t x = x0 + (l & 3), y = y0 + (l >> 2);
                    if (f->flags & VRT_FRAME_LINEAR_OUTPUT) {
                        uint32_t* o = (uint32_t*)out;
                        size_t n = (size_t)w * h, p = (size_t)y * w + x;
                        o[p] = alb[l];
                        memcpy(&o[n + p], &dep[l], 4);
                        o[2 * n + p] = rg[l];
                        o[3 * n + p] = bx[l];
                    } else { /* Framebuffer::Tile, :299-309, tile = (y/4)*TileStride + x/4 (:459) */
                        VrtTile* t = (VrtTile*)out + ((size_t)(y / 4) * (w / 4) + x / 4);
                        t->albedo[l] = alb[l];
                        t->depth[l] = dep[l];
                        t->irr_rg[l] = rg[l];
                        t->irr_bx[l] = bx[l];
                    }
                }
            }
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        stats_merge(&total, &local);
    }
    if (stats) *stats = total;
}

/* ------------------------------------------------------------------------------------------ */
/* GLSL-renderer ray casts: rayCast / rayCastCoarse / getStepPos of the     */
/* reference's GPU renderer, VoxelRT/Shaders/VoxelTraversal.glsl:14-52,92-131,162-243, with      */
/* VoxelMap.glsl:28-76 addressing.  PARITY UNPINNED (GLSL needs a GL device).  Canonical          */
/* arithmetic: IEEE binary32 RN, one operation at a time, no FMA; min(a,b) = b<a ? b : a,         */
/* max(a,b) = a<b ? b : a; ivec3(floor(x)) of NaN / out-of-range values = INT_MIN (out of grid).  */
/* ------------------------------------------------------------------------------------------ */
static inline float glsl_min(float a, float b) { return b < a ? b : a; }
static inline float glsl_max(float a, float b) { return a < b ? b : a; }
static inline int32_t glsl_floor2i(float x) {
    float f = floorf(x);
    return (f >= -2147483648.0f && f < 2147483648.0f) ? (int32_t)f : (int32_t)0x80000000;
}

/* GenerateRayCellInteractionMaskLUT, VoxelRT/GpuRenderer.cpp:193-210: for every direction octant and origin cell of a
 * 4x4x4 mask, the cells a ray travelling in that octant can still reach. */
void orc_interaction_lut(uint64_t table[512]) {
    for (uint32_t oct = 0; oct < 8; oct++) {
        int dir[3] = {(int)(oct & 1) * 2 - 1, (int)((oct >> 1) & 1) * 2 - 1, (int)((oct >> 2) & 1) * 2 - 1}; /* x, y, z */
        for (uint32_t origin = 0; origin < 64; origin++) {
            int o[3] = {(int)(origin & 3), (int)((origin >> 4) & 3), (int)((origin >> 2) & 3)}; /* MaskIndexer: x | z<<2 | y<<4 */
            uint64_t mask = 0;
            for (uint32_t j = 0; j < 64; j++) {
                int p[3] = {o[0] + (int)(j & 3) * dir[0], o[1] + (int)((j >> 4) & 3) * dir[1], o[2] + (int)((j >> 2) & 3) * dir[2]};
                if ((unsigned)p[0] < 4 && (unsigned)p[1] < 4 && (unsigned)p[2] < 4) mask |= 1ull << (p[0] | p[2] << 2 | p[1] << 4);
            }
            table[origin + oct * 64] = mask;
        }
    }
}

/* SectorMasks[] of the GPU storage (GpuRenderer.cpp:134-142): bit = sector has any brick allocated */
static uint64_t sector_group_mask(const OrcMap* m, uint32_t gx, uint32_t gy, uint32_t gz) {
    uint64_t g = 0;
    for (uint32_t i = 0; i < 64; i++) {
        uint32_t si = sector_index(m, (int32_t)(gx * 4 + (i & 3)), (int32_t)(gy * 4 + ((i >> 4) & 3)), (int32_t)(gz * 4 + ((i >> 2) & 3)));
        if (m->sector_masks[si]) g |= 1ull << i;
    }
    return g;
}

/* getStepPos, VoxelTraversal.glsl:92-131.  Returns 1 when the ray may step on (p moved to the far corner of the empty
 * cell), 0 on a hit (coarse hits move p to an occupied voxel of the 4^3 cell, :127). */
static int glsl_step_pos(const OrcMap* m, int32_t p[3], const float d[3], int coarse, int aniso, const uint64_t* lut, CastCounters* c) {
    uint32_t si = sector_index(m, p[0] >> 5, p[1] >> 5, p[2] >> 5);
    uint64_t mask = m->sector_masks[si]; /* BrickMasks[] = allocation mask */
    c->sector_fetches++;
    uint32_t idx = ((uint32_t)(p[0] >> 3) & 3) | (((uint32_t)(p[2] >> 3) & 3) << 2) | (((uint32_t)(p[1] >> 3) & 3) << 4);
    int scale = 8;
    if ((mask >> idx) & 1) { /* isFineLod: brick allocated -> its 4^3 cell mask (:101-109) */
        uint32_t cell = ((uint32_t)(p[0] >> 2) & 1) | (((uint32_t)(p[2] >> 2) & 1) << 1) | (((uint32_t)(p[1] >> 2) & 1) << 2);
        mask = m->sector_cells[si] ? m->sector_cells[si][idx * 8 + cell] : 0;
        c->cell_fetches++;
        idx = ((uint32_t)p[0] & 3) | (((uint32_t)p[2] & 3) << 2) | (((uint32_t)p[1] & 3) << 4);
        scale = 1;
        if ((mask >> idx) & 1) return 0;
    } else if (mask == 0) { /* sector without bricks -> the 128^3 level (:110-114) */
        mask = sector_group_mask(m, (uint32_t)(p[0] >> 7), (uint32_t)(p[1] >> 7), (uint32_t)(p[2] >> 7));
        idx = ((uint32_t)(p[0] >> 5) & 3) | (((uint32_t)(p[2] >> 5) & 3) << 2) | (((uint32_t)(p[1] >> 5) & 3) << 4);
        scale = 32;
    }
    if (aniso) { /* :116-121 */
        uint32_t oct = (d[0] < 0 ? 0u : 1u) + (d[1] < 0 ? 0u : 2u) + (d[2] < 0 ? 0u : 4u);
        mask &= lut[idx + oct * 64];
    }
    uint32_t half = idx < 32 ? (uint32_t)mask : (uint32_t)(mask >> 32);
    int lod = (mask == 0 ? 4 : (((half >> (idx & 0xA)) & 0x00330033u) == 0 ? 2 : 1)) * scale; /* getIsotropicLod :22-32 */
    if (coarse && lod < 4) { /* findAnyOccupiedPos :40-52 */
        uint32_t cur = (uint32_t)mask, bit = 0;
        if (cur == 0) { bit = 32; cur = (uint32_t)(mask >> 32); }
        bit += (uint32_t)__builtin_ctz(cur);
        p[0] = (p[0] & ~3) | (int32_t)(bit & 3);
        p[1] = (p[1] & ~3) | (int32_t)((bit >> 4) & 3);
        p[2] = (p[2] & ~3) | (int32_t)((bit >> 2) & 3);
        return 0;
    }
    int32_t cm = lod - 1;
    for (int a = 0; a < 3; a++) p[a] = d[a] < 0 ? (p[a] & ~cm) : (p[a] | cm); /* alignToCellBoundaries :33-39 */
    return 1;
}

/* clipRayToAABB, VoxelTraversal.glsl:133-145 with the bounds rayCast passes (:173,207) */
static void glsl_clip(const OrcMap* m, const float o[3], const float d[3], const int32_t wo[3], float out[3]) {
    float grid[3] = {(float)(1u << (m->shift_xz + 5)), (float)(1u << (m->shift_y + 5)), (float)(1u << (m->shift_xz + 5))};
    float t1[3], t2[3];
    for (int a = 0; a < 3; a++) {
        float inv = 1.0f / d[a];
        float lo = (float)(int32_t)(1u - (uint32_t)wo[a]);        /* -u_WorldOrigin + 1 (integer arithmetic) */
        float hi = (grid[a] - (float)wo[a]) - 1.0f;               /* vec3(GRID) - u_WorldOrigin - 1          */
        float a1 = (lo - o[a]) * inv, a2 = (hi - o[a]) * inv;
        t1[a] = glsl_min(a1, a2);
        t2[a] = glsl_max(a1, a2);
    }
    float tmin = glsl_max(t1[0], glsl_max(t1[1], t1[2]));
    float tmax = glsl_min(t2[0], glsl_min(t2[1], t2[2]));
    int clip = tmin > 0.0f && tmin < tmax;
    for (int a = 0; a < 3; a++) out[a] = clip ? o[a] + d[a] * tmin : o[a];
}

static void glsl_cast(const OrcMap* m, const float o_in[3], const float d[3], const int32_t wo[3], uint32_t flags, const uint64_t* lut,
                      VrtHit* out, CastCounters* cnt) {
    const int coarse_mode = (flags & VRT_GLSL_COARSE) != 0, aniso = (flags & VRT_GLSL_ANISOTROPIC) != 0;
    const uint32_t cap = coarse_mode ? 96u : 256u; /* :176,211 */
    float o[3] = {o_in[0], o_in[1], o_in[2]}, start[3], inv[3], ts[3], sd[3] = {0, 0, 0}, cur[3] = {0, 0, 0};
    glsl_clip(m, o, d, wo, start);
    if (coarse_mode) memcpy(o, start, sizeof(o)); /* rayCastCoarse moves the origin itself (:207) */
    for (int a = 0; a < 3; a++) {
        inv[a] = 1.0f / d[a];
        ts[a] = ((d[a] < 0.0f ? 0.0f : 1.0f) - o[a]) * inv[a]; /* (step(0, dir) - origin) * invDir */
    }
    int32_t p[3];
    for (int a = 0; a < 3; a++) p[a] = (int32_t)((uint32_t)wo[a] + (uint32_t)glsl_floor2i(start[a]));
    int hit = 0, inb = 1;
    float tmin = 0.0f;
    uint32_t i = 0;
    for (; i < cap; i++) {
        cnt->iters++;
        for (int a = 0; a < 3; a++) sd[a] = ts[a] + (float)(int32_t)((uint32_t)p[a] - (uint32_t)wo[a]) * inv[a];
        tmin = glsl_min(glsl_min(sd[0], sd[1]), sd[2]);
        tmin = coarse_mode ? tmin + 0.001f : (tmin == tmin ? u2f(f2u(tmin) + 5u) : tmin); /* :215 / :179-180; a NaN stays a NaN whatever its payload */
        for (int a = 0; a < 3; a++) cur[a] = o[a] + tmin * d[a];
        for (int a = 0; a < 3; a++) p[a] = (int32_t)((uint32_t)wo[a] + (uint32_t)glsl_floor2i(cur[a]));
        inb = inbound(m, p[0], p[1], p[2]);
        if (!inb) break;
        if (!glsl_step_pos(m, p, d, coarse_mode && i > 30, aniso, lut, cnt)) { /* :218 coarse = i > 30 */
            hit = 1;
            break;
        }
    }
    int capped = i >= cap;
    int sm[3] = {tmin >= sd[0], tmin >= sd[1], tmin >= sd[2]}; /* :198,228 */
    int nrm[3];
    for (int a = 0; a < 3; a++) nrm[a] = sm[a] ? (d[a] > 0.0f ? -1 : (d[a] < 0.0f ? 1 : 0)) : 0; /* mix(0, -sign(dir), sideMask) */
    out->vx = p[0];
    out->vy = p[1];
    out->vz = p[2];
    out->material = hit ? voxel_material(m, p[0], p[1], p[2]) : 0u;
    out->dist = tmin;
    out->px = cur[0];
    out->py = cur[1];
    out->pz = cur[2];
    float fu = sm[0] ? cur[1] : cur[0], fv = sm[2] ? cur[1] : cur[2]; /* fract(mix(currPos.xz, currPos.yy, sideMask.xz)) */
    out->u = hit ? fu - floorf(fu) : 0.0f;
    out->v = hit ? fv - floorf(fv) : 0.0f;
    uint32_t iters_done = capped ? cap : i;
    out->flags = (hit ? (uint32_t)((nrm[0] + 1) | ((nrm[1] + 1) << 2) | ((nrm[2] + 1) << 4)) : 21u) | (hit ? VRT_HIT_HIT : 0) |
                 (inb ? VRT_HIT_INBOUND : 0) | (capped ? VRT_HIT_CAPPED : 0) | (iters_done << VRT_HIT_ITERS_SHIFT);
    out->_pad = 0;
}

void orc_trace_glsl(const OrcMap* m, uint64_t n, const float* origin3, const float* dir3, const int32_t wo[3], uint32_t flags,
                    VrtHit* out, OrcStats* stats, int threads) {
    uint64_t lut[512];
    orc_interaction_lut(lut);
    OrcStats total;
    memset(&total, 0, sizeof(total));
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel num_threads(threads)
#endif
    {
        OrcStats local;
        memset(&local, 0, sizeof(local));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4096)
#endif
        for (int64_t i = 0; i < (int64_t)n; i++) {
            CastCounters c = {0, 0, 0, {0}};
            glsl_cast(m, origin3 + 3 * i, dir3 + 3 * i, wo, flags, lut, &out[i], &c);
            stats_add(&local, &c, &out[i]);
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        stats_merge(&total, &local);
    }
    if (stats) *stats = total;
}

/* ---------------------------------------------------------------------------------------------- */
/* The GPU renderer's frame shader: main() of VoxelRT/Shaders/VoxelRender.comp:29-93 per pixel, with */
/* getPrimaryRay :20-25, getSkyColor :15-17, random_dir / blueNoise (RandomGen.glsl:27-48),          */
/* getMaterialColor / getMaterialEmission (VoxelMap.glsl:78-84), packGNormal (GBuffer.glsl:17-20).   */
/* PARITY UNPINNED (GLSL needs a GL device).  Canonical arithmetic: fp32, one operation at a time, no */
/* contraction (this file is compiled with -ffp-contract=off); mat * vec summed column by column;     */
/* normalize(v) = v * (1 / sqrt(dot(v, v))); sin / cos through the CPU renderer's sincos_2pi on the   */
/* blue-noise fraction; imageStore: rgba8 = rint(clamp(x, 0, 1) * 255), rgba16f = round to nearest    */
/* even.  Stand-ins: the sky is the map's cube (the CPU renderer's) at level 0 through ProjectCubemap */
/* + nearest, not GL's seamless bilinear fetch of the cube PanoramaToCube.comp builds; a HitInfo      */
/* field the shader leaves unassigned on a path keeps its previous value.                             */
/* ---------------------------------------------------------------------------------------------- */
typedef struct {
    float pos[3], nrm[3];
    uint32_t mat, iters;
} GlslHitInfo;

static int glsl_cast_info(const OrcMap* m, const float o[3], const float d[3], const int32_t wo[3], uint32_t flags, const uint64_t* lut, GlslHitInfo* H,
                          CastCounters* cnt) {
    VrtHit r;
    glsl_cast(m, o, d, wo, flags, lut, &r, cnt);
    if (!(r.flags & VRT_HIT_CAPPED)) H->iters = r.flags >> VRT_HIT_ITERS_SHIFT; /* :189,199 / :223,237; the final `return false` assigns nothing */
    if (!(r.flags & VRT_HIT_HIT)) return 0;
    H->mat = r.material;
    H->pos[0] = r.px, H->pos[1] = r.py, H->pos[2] = r.pz;
    for (int a = 0; a < 3; a++) H->nrm[a] = (float)((int)((r.flags >> (2 * a)) & 3u) - 1);
    return 1;
}
static void glsl_normalize(float v[3]) {
    float len2 = (v[0] * v[0] + v[1] * v[1]) + v[2] * v[2];
    float k = 1.0f / sqrtf(len2);
    v[0] *= k, v[1] *= k, v[2] *= k;
}
static void glsl_material_color(uint32_t md, float c[3]) {
    c[0] = (float)((md >> 11) & 31u) * (1.0f / 31.0f);
    c[1] = (float)((md >> 5) & 63u) * (1.0f / 63.0f);
    c[2] = (float)(md & 31u) * (1.0f / 31.0f);
    for (int a = 0; a < 3; a++) c[a] = c[a] * c[a];
}
static void glsl_sky(const OrcMap* m, const float d[3], float c[3]) {
    if (!m->sky_texels) {
        c[0] = c[1] = c[2] = 0.0f;
        return;
    }
    float t[3];
    orc_sky_sample(m, d, 0, t); /* texel * 3 (exact: 6 + 2 significant bits), so / 3 returns the texel */
    for (int a = 0; a < 3; a++) c[a] = glsl_min((t[a] / 3.0f) * 5.0f, 50000.0f);
}
static void glsl_random_dir(const OrcMap* m, uint32_t x, uint32_t y, uint32_t frame_no, uint32_t i, float r[3]) {
    float fi = (float)i, ox = fi * 0.75487766624669276005f, oy = fi * 0.56984029099805326591f;
    float sx = ox + 0.5f, sy0 = oy + 0.5f;
    sx -= floorf(sx);
    sy0 -= floorf(sy0);
    uint32_t px = (x + (uint32_t)(sx * 128.0f)) & 127u, py = ((y + (uint32_t)(sy0 * 128.0f)) & 127u) + (frame_no & 63u) * 128u;
    const uint8_t* t = m->blue_noise + ((size_t)py * 128 + px) * 2;
    float nx = ((float)t[0] + 0.5f) * (1.0f / 256.0f), ny = ((float)t[1] + 0.5f) * (1.0f / 256.0f);
    float yy = nx * 2.0f - 1.0f, s, c;
    sincos_2pi(ny, &s, &c);
    float sy = sqrtf(1.0f - yy * yy);
    r[0] = s * sy, r[1] = yy, r[2] = c * sy;
}
static uint32_t glsl_unorm8(float v) { return (uint32_t)(int32_t)nearbyintf(glsl_min(glsl_max(v, 0.0f), 1.0f) * 255.0f); }

static void glsl_frame_pixel(const OrcMap* m, const VrtFrame* f, const uint64_t* lut, uint32_t x, uint32_t y, uint32_t* o_albedo, float* o_depth,
                             uint32_t* o_rg, uint32_t* o_bx, CastCounters* cnt) {
    const int32_t* wo = f->world_origin;
    const float* iv = f->inv_proj;
    float pos[3], dir[3], nr[4], fr[4];
    const float fx = (float)(int32_t)x, fy = (float)(int32_t)y;
    for (int k = 0; k < 4; k++) { /* getPrimaryRay */
        nr[k] = ((iv[k] * fx + iv[4 + k] * fy) + iv[8 + k] * 0.0f) + iv[12 + k] * 1.0f;
        fr[k] = nr[k] + iv[8 + k];
    }
    const float in = 1.0f / nr[3], iff = 1.0f / fr[3];
    for (int a = 0; a < 3; a++) {
        pos[a] = nr[a] * in + f->origin_frac[a];
        dir[a] = fr[a] * iff;
    }
    glsl_normalize(dir);
    float albedo[3], irr[3], nrm[3] = {0, 0, 0}, depth = -1.0f;
    GlslHitInfo hit, sun_hit;
    memset(&hit, 0, sizeof(hit));
    const uint32_t fine = (f->flags & VRT_FRAME_GLSL_ANISOTROPIC) ? VRT_GLSL_ANISOTROPIC : 0u, coarse = fine | VRT_GLSL_COARSE;
    if (glsl_cast_info(m, pos, dir, wo, fine, lut, &hit, cnt)) {
        glsl_material_color(hit.mat, albedo);
        memcpy(nrm, hit.nrm, sizeof(nrm));
        {
            const float hx = hit.pos[0] * 0.0625f, hy = hit.pos[1] * 0.0625f, hz = hit.pos[2] * 0.0625f;
            const float* q = f->proj;
            const float pz = ((q[2] * hx + q[6] * hy) + q[10] * hz) + q[14] * 1.0f;
            const float pw = ((q[3] * hx + q[7] * hy) + q[11] * hz) + q[15] * 1.0f;
            depth = pz / pw;
        }
        const float em0 = f16_to_f32((uint16_t)(hit.mat >> 16));
        float thr[3] = {1.0f, 1.0f, 1.0f};
        for (int a = 0; a < 3; a++) irr[a] = f->bounces == 0 ? albedo[a] : albedo[a] * em0;
        float sun[3] = {0.3f, 0.9f, -0.28f};
        glsl_normalize(sun);
        const float sun_intensity = 5.0f;
        const float sun_col[3] = {1.2f * sun_intensity, 1.1f * sun_intensity, 1.0f * sun_intensity};
        sun_hit = hit;
        if (f->bounces != 0) {
            const float so[3] = {hit.pos[0] + hit.nrm[0] * 0.01f, hit.pos[1] + hit.nrm[1] * 0.01f, hit.pos[2] + hit.nrm[2] * 0.01f};
            if (!glsl_cast_info(m, so, sun, wo, coarse, lut, &sun_hit, cnt)) {
                for (int a = 0; a < 3; a++) irr[a] = irr[a] + sun_col[a];
            } else {
                for (int a = 0; a < 3; a++) thr[a] = thr[a] * 0.5f;
            }
        }
        for (uint32_t i = 0; i < f->bounces; i++) {
            float rnd[3];
            glsl_random_dir(m, x, y, f->frame_no, i, rnd);
            for (int a = 0; a < 3; a++) {
                pos[a] = hit.pos[a] + hit.nrm[a] * 0.01f;
                dir[a] = hit.nrm[a] + rnd[a];
            }
            glsl_normalize(dir);
            if (!glsl_cast_info(m, pos, dir, wo, coarse, lut, &hit, cnt)) {
                float sky[3];
                glsl_sky(m, dir, sky);
                for (int a = 0; a < 3; a++) irr[a] = irr[a] + thr[a] * sky[a];
                break;
            }
            float col[3];
            glsl_material_color(hit.mat, col);
            for (int a = 0; a < 3; a++) thr[a] = thr[a] * col[a];
            float emission = f16_to_f32((uint16_t)(hit.mat >> 16));
            if (i < 2u) {
                const float so[3] = {hit.pos[0] + hit.nrm[0] * 0.01f, hit.pos[1] + hit.nrm[1] * 0.01f, hit.pos[2] + hit.nrm[2] * 0.01f};
                if (!glsl_cast_info(m, so, sun, wo, coarse, lut, &sun_hit, cnt)) {
                    for (int a = 0; a < 3; a++) thr[a] = thr[a] * sun_col[a];
                    emission = emission + sun_intensity;
                }
            }
            for (int a = 0; a < 3; a++) irr[a] = irr[a] + thr[a] * emission;
        }
    } else {
        glsl_sky(m, dir, irr);
        albedo[0] = albedo[1] = albedo[2] = 1.0f;
    }
    uint32_t code = 0;
    for (int a = 0; a < 3; a++) code |= (uint32_t)(int32_t)glsl_min(glsl_max(nrm[a] + 1.0f, 0.0f), 3.0f) << (2 * a);
    *o_albedo = glsl_unorm8(albedo[0]) | (glsl_unorm8(albedo[1]) << 8) | (glsl_unorm8(albedo[2]) << 16) | (code << 24);
    *o_depth = depth;
    *o_rg = (uint32_t)f32_to_f16(irr[0]) | ((uint32_t)f32_to_f16(irr[1]) << 16);
    *o_bx = (uint32_t)f32_to_f16(irr[2]) | ((uint32_t)f32_to_f16((float)hit.iters) << 16);
}

/* A frame of the GPU renderer (GpuRenderer::RenderFrame's dispatch, GpuRenderer.cpp:257-268) into the 16 B/px tile framebuffer (or the 4
 * planes of VRT_FRAME_LINEAR_OUTPUT); honours part_index / part_count like orc_render.  stats->iters = trips of all casts. */
void orc_render_glsl(const OrcMap* m, const VrtFrame* f, void* out, OrcStats* stats, int threads) {
    const uint32_t w = f->width, h = f->height;
    uint64_t lut[512];
    orc_interaction_lut(lut);
    const uint32_t part_count = f->part_count ? f->part_count : 1, tiles_x = (w + 31) / 32;
    uint64_t iters = 0;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads) reduction(+ : iters)
#endif
    for (int64_t y = 0; y < (int64_t)h; y++) {
        CastCounters c = {0, 0, 0, {0}};
        for (uint32_t x = 0; x < w; x++) {
            if (part_count > 1) {
                uint32_t t = (f->flags & VRT_FRAME_PART_ROWS) ? ((uint32_t)y / VRT_BAND_ROWS) : ((uint32_t)y / 32) * tiles_x + (x / 32);
                if (t % part_count != f->part_index) continue;
            }
            uint32_t alb, rg, bx;
            float dep;
            glsl_frame_pixel(m, f, lut, x, (uint32_t)y, &alb, &dep, &rg, &bx, &c);
            if (f->flags & VRT_FRAME_LINEAR_OUTPUT) {
                uint32_t* o = (uint32_t*)out;
                size_t n = (size_t)w * h, p = (size_t)y * w + x;
                o[p] = alb;
                memcpy(&o[n + p], &dep, 4);
                o[2 * n + p] = rg;
                o[3 * n + p] = bx;
            } else {
                VrtTile* t = (VrtTile*)out + ((size_t)(y / 4) * (w / 4) + x / 4);
                const uint32_t l = (x & 3) | (((uint32_t)y & 3) << 2);
                t->albedo[l] = alb;
                t->depth[l] = dep;
                t->irr_rg[l] = rg;
                t->irr_bx[l] = bx;
            }
        }
        iters += c.iters;
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->iters = iters;
    }
}
