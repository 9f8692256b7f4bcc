// Ray casts with the semantics of the reference's GPU renderer: rayCast / rayCastCoarse /
// getStepPos of src/VoxelRT/Shaders/VoxelTraversal.glsl:14-52,92-131,133-243 over the resident brickmap of this library —
// the extra 128^3 level (one 64-bit mask per 4x4x4 group of sectors, VoxelTraversal.glsl:110-114), the +5-ulp bias, the
// clip of outside origins to the grid box, 256 / 96 iterations, coarse stops on occupied 4^3 cells after 30 iterations
// and the optional ray/cell interaction-mask LUT (GpuRenderer.cpp:193-210).  These change which cells a ray visits and
// therefore normals and iteration counts relative to the CPU renderer, so they live in their OWN entry point
// (vrt_trace_glsl) and never touch the bit-exact frame kernels.  Arithmetic: fp32 RN, one operation at a time, no FMA
// (the canonical form the parity oracle defines for this entry point; GLSL itself leaves contraction to the compiler).
#pragma once
#include "vrt_device.cuh"

namespace vrt {

struct GlslScene {
    const uint2* __restrict__ groups;  // SectorMasks[]: bit (x | z<<2 | y<<4) of group (sx>>2, sy>>2, sz>>2) = sector has bricks
    const uint2* __restrict__ lut;     // RayCellInteractionMaskLUT[64 * 8]
    uint32_t gxz;                      // log2 of the group grid extent along x and z
};

// one group per thread: OR the 64 sector headers of the group into a mask (GpuRenderer.cpp:134-142)
__global__ void k_build_groups(DevScene S, uint2* groups, uint32_t gxz, uint32_t n_groups) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint32_t gx = g & ((1u << gxz) - 1), gz = (g >> gxz) & ((1u << gxz) - 1), gy = g >> (2 * gxz);
    uint32_t lo = 0, hi = 0;
    for (uint32_t i = 0; i < 64; i++) {
        const uint4 h = ldg_hdr(S.hdr + hdr_index(S.sxp, S.sxzp, (int)(gx * 4 + (i & 3)), (int)(gy * 4 + ((i >> 4) & 3)), (int)(gz * 4 + ((i >> 2) & 3))));
        if (h.x | h.y) (i < 32 ? lo : hi) |= 1u << (i & 31);
    }
    groups[g] = make_uint2(lo, hi);
}

__device__ __forceinline__ float glsl_min(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float glsl_max(float a, float b) { return a < b ? b : a; }
__device__ __forceinline__ int glsl_floor2i(float x) {
    const float f = floorf(x);
    return (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : (int)0x80000000;
}

// What rayCast / rayCastCoarse leave behind (VoxelTraversal.glsl:162-243) in raw form: the voxel the walk stopped in, the entry distance
// and point, the three side distances the normal is taken from.
struct GlslCast {
    int p[3];
    float tmin, cur[3], sd[3], d[3];
    bool hit, inb, capped;
    uint32_t iters;  // trips made (the cap when capped)
};

__device__ inline void glsl_cast(const DevScene& S, const GlslScene& G, const int wo[3], const float o_in[3], const float d_in[3], uint32_t flags, GlslCast& C) {
    const bool coarse_mode = (flags & VRT_GLSL_COARSE) != 0, aniso = (flags & VRT_GLSL_ANISOTROPIC) != 0;
    const uint32_t cap = coarse_mode ? 96u : 256u;
    float o[3] = {o_in[0], o_in[1], o_in[2]};
    const float d[3] = {d_in[0], d_in[1], d_in[2]};
    float inv[3], ts[3], start[3];
    {  // clipRayToAABB(origin, dir, -wo + 1, grid - wo - 1), VoxelTraversal.glsl:133-145,173,207
        const float grid[3] = {(float)S.lim_xz, (float)S.lim_y, (float)S.lim_xz};
        float t1[3], t2[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            inv[a] = __fdiv_rn(1.0f, d[a]);
            const float lo = (float)(int)(1u - (uint32_t)wo[a]);
            const float hi = __fsub_rn(__fsub_rn(grid[a], (float)wo[a]), 1.0f);
            const float a1 = __fmul_rn(__fsub_rn(lo, o[a]), inv[a]), a2 = __fmul_rn(__fsub_rn(hi, o[a]), inv[a]);
            t1[a] = glsl_min(a1, a2);
            t2[a] = glsl_max(a1, a2);
        }
        const float tmin = glsl_max(t1[0], glsl_max(t1[1], t1[2]));
        const float tmax = glsl_min(t2[0], glsl_min(t2[1], t2[2]));
        const bool clip = tmin > 0.0f && tmin < tmax;
#pragma unroll
        for (int a = 0; a < 3; a++) start[a] = clip ? __fadd_rn(o[a], __fmul_rn(d[a], tmin)) : o[a];
    }
    if (coarse_mode) {
#pragma unroll
        for (int a = 0; a < 3; a++) o[a] = start[a];
    }
    int p[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        ts[a] = __fmul_rn(__fsub_rn(d[a] < 0.0f ? 0.0f : 1.0f, o[a]), inv[a]);
        p[a] = (int)((uint32_t)wo[a] + (uint32_t)glsl_floor2i(start[a]));
    }
    const uint32_t oct = (d[0] < 0.0f ? 0u : 1u) + (d[1] < 0.0f ? 0u : 2u) + (d[2] < 0.0f ? 0u : 4u);
    bool hit = false, inb = true;
    float tmin = 0.0f, sd[3] = {0.0f, 0.0f, 0.0f}, cur[3] = {0.0f, 0.0f, 0.0f};
    uint32_t i = 0;
    for (; i < cap; i++) {
#pragma unroll
        for (int a = 0; a < 3; a++) sd[a] = __fadd_rn(ts[a], __fmul_rn((float)(int)((uint32_t)p[a] - (uint32_t)wo[a]), inv[a]));
        tmin = glsl_min(glsl_min(sd[0], sd[1]), sd[2]);
        tmin = coarse_mode ? __fadd_rn(tmin, 0.001f) : (tmin == tmin ? __uint_as_float(__float_as_uint(tmin) + 5u) : tmin);  // a NaN stays a NaN whatever its payload
#pragma unroll
        for (int a = 0; a < 3; a++) {
            cur[a] = __fadd_rn(o[a], __fmul_rn(tmin, d[a]));
            p[a] = (int)((uint32_t)wo[a] + (uint32_t)glsl_floor2i(cur[a]));
        }
        inb = (uint32_t)(p[0] | p[2]) < S.lim_xz && (uint32_t)p[1] < S.lim_y;
        if (!inb) break;
        // getStepPos, VoxelTraversal.glsl:92-131
        const uint4 h = ldg_hdr(S.hdr + hdr_index(S.sxp, S.sxzp, p[0] >> 5, p[1] >> 5, p[2] >> 5));
        uint32_t lo = h.x, hi = h.y;
        uint32_t idx = ((uint32_t)(p[0] >> 3) & 3u) | (((uint32_t)(p[2] >> 3) & 3u) << 2) | (((uint32_t)(p[1] >> 3) & 3u) << 4);
        int scale = 8;
        if ((((idx & 32u) ? hi : lo) >> (idx & 31u)) & 1u) {
            const uint32_t cell = ((uint32_t)(p[0] >> 2) & 1u) | (((uint32_t)(p[2] >> 2) & 1u) << 1) | (((uint32_t)(p[1] >> 2) & 1u) << 2);
            const uint2 m = ldg_u2(S.cells + (size_t)brick_slot(h, idx) * 8u + cell);
            lo = m.x, hi = m.y;
            idx = ((uint32_t)p[0] & 3u) | (((uint32_t)p[2] & 3u) << 2) | (((uint32_t)p[1] & 3u) << 4);
            scale = 1;
            if ((((idx & 32u) ? hi : lo) >> (idx & 31u)) & 1u) {
                hit = true;
                break;
            }
        } else if ((lo | hi) == 0u) {
            const uint32_t g = (uint32_t)(p[0] >> 7) | ((uint32_t)(p[2] >> 7) << G.gxz) | ((uint32_t)(p[1] >> 7) << (2 * G.gxz));
            const uint2 m = ldg_u2(G.groups + g);
            lo = m.x, hi = m.y;
            idx = ((uint32_t)(p[0] >> 5) & 3u) | (((uint32_t)(p[2] >> 5) & 3u) << 2) | (((uint32_t)(p[1] >> 5) & 3u) << 4);
            scale = 32;
        }
        if (aniso) {
            const uint2 l = ldg_u2(G.lut + idx + oct * 64u);
            lo &= l.x, hi &= l.y;
        }
        const uint32_t half = (idx & 32u) ? hi : lo;
        const int lod = ((lo | hi) == 0u ? 4 : (((half >> (idx & 0xAu)) & 0x00330033u) == 0u ? 2 : 1)) * scale;
        if (coarse_mode && i > 30u && lod < 4) {  // findAnyOccupiedPos, :40-52
            const uint32_t bit = lo ? (uint32_t)(__ffs((int)lo) - 1) : 32u + (uint32_t)(__ffs((int)hi) - 1);
            p[0] = (p[0] & ~3) | (int)(bit & 3u);
            p[1] = (p[1] & ~3) | (int)((bit >> 4) & 3u);
            p[2] = (p[2] & ~3) | (int)((bit >> 2) & 3u);
            hit = true;
            break;
        }
        const int cm = lod - 1;
#pragma unroll
        for (int a = 0; a < 3; a++) p[a] = d[a] < 0.0f ? (p[a] & ~cm) : (p[a] | cm);
    }
    C.p[0] = p[0], C.p[1] = p[1], C.p[2] = p[2];
    C.tmin = tmin;
#pragma unroll
    for (int a = 0; a < 3; a++) C.cur[a] = cur[a], C.sd[a] = sd[a], C.d[a] = d[a];
    C.hit = hit, C.inb = inb, C.capped = i >= cap;
    C.iters = C.capped ? cap : i;
}
// hit.normal = mix(vec3(0), -sign(dir), greaterThanEqual(vec3(tmin), sideDist)) (VoxelTraversal.glsl:195-197), one axis
__device__ __forceinline__ int glsl_normal(const GlslCast& C, int a) { return (C.tmin >= C.sd[a]) ? (C.d[a] > 0.0f ? -1 : (C.d[a] < 0.0f ? 1 : 0)) : 0; }

__global__ void __launch_bounds__(128) k_trace_glsl(DevScene S, GlslScene G, int wx, int wy, int wz, const float* __restrict__ o3,
                                                    const float* __restrict__ d3, uint32_t flags, uint64_t n, VrtHit* __restrict__ out) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float o[3] = {o3[3 * r], o3[3 * r + 1], o3[3 * r + 2]};
    const float d[3] = {d3[3 * r], d3[3 * r + 1], d3[3 * r + 2]};
    const int wo[3] = {wx, wy, wz};
    GlslCast C;
    glsl_cast(S, G, wo, o, d, flags, C);
    const bool hit = C.hit, inb = C.inb, capped = C.capped;
    const int* p = C.p;
    const float tmin = C.tmin;
    const float* cur = C.cur;
    const float* sd = C.sd;
    const uint32_t cap = C.iters, i = C.iters;
    VrtHit H;
    H.vx = p[0], H.vy = p[1], H.vz = p[2];
    H.material = hit ? __ldg(&S.palette[voxel_palette_id(S, p[0], p[1], p[2])]).x : 0u;
    H.dist = tmin;
    H.px = cur[0], H.py = cur[1], H.pz = cur[2];
    uint32_t code = 21u;
    H.u = 0.0f, H.v = 0.0f;
    if (hit) {
        code = 0;
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const int nrm = glsl_normal(C, a);
            code |= (uint32_t)(nrm + 1) << (2 * a);
        }
        const float fu = (tmin >= sd[0]) ? cur[1] : cur[0], fv = (tmin >= sd[2]) ? cur[1] : cur[2];
        H.u = __fsub_rn(fu, floorf(fu));
        H.v = __fsub_rn(fv, floorf(fv));
    }
    H.flags = code | (hit ? VRT_HIT_HIT : 0u) | (inb ? VRT_HIT_INBOUND : 0u) | (capped ? VRT_HIT_CAPPED : 0u) | ((capped ? cap : i) << VRT_HIT_ITERS_SHIFT);
    H._pad = 0;
    out[r] = H;
}

}  // namespace vrt
