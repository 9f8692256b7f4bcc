// The frame shader of the reference's GPU renderer: src/VoxelRT/Shaders/VoxelRender.comp:18-93 for one pixel —
// getPrimaryRay, rayCast, the sun shadow ray, u_MaxBounces diffuse bounces through rayCastCoarse with a sun ray after each of the first two,
// sky on a miss — over the casts of vrt_glsl.cuh, behind vrt_render_glsl.  One thread per pixel (the shader has no cross-lane coupling);
// warp tiles and the 16 B/px tile framebuffer as in the CPU-renderer frame kernels, so the GBuffer step consumes either.
//
// PARITY UNPINNED (GLSL needs a GL device; its float expressions have no defined contraction or intrinsic precision).  The oracle
// (orc_render_glsl) and this file share one canonical arithmetic: fp32 round-to-nearest, one operation at a time, no FMA;
// mat * vec summed column by column left to right; normalize(v) = v * (1 / sqrt(dot(v, v))); sin / cos of the bounce direction = the CPU
// renderer's sincos_2pi polynomial (SIMD.h:175-190) on the blue-noise fraction; imageStore conversions: rgba8 = rint(clamp(x, 0, 1) * 255),
// rgba16f = round-to-nearest-even.  Two stand-ins, both stated in the header of vrt_render_glsl: the sky is the context's cube (the CPU
// renderer's, vrt_set_sky) read at its level 0 with ProjectCubemap + nearest instead of GL's seamless bilinear fetch of the cube that
// PanoramaToCube.comp builds; and an `out HitInfo` field the shader does not assign on a path keeps its previous value.
#pragma once
#include "vrt_glsl.cuh"
#include "vrt_kernels.cuh"

namespace vrt {

struct GlslHitInfo {  // HitInfo, VoxelTraversal.glsl:153-160 (uv is never read by the frame shader)
    float pos[3];
    float nrm[3];
    uint32_t mat;    // Material.Data.x
    uint32_t iters;
};

// rayCast / rayCastCoarse as the frame shader sees them: returns the bool, assigns what the shader assigns on that path
__device__ inline bool glsl_cast_info(const DevScene& S, const GlslScene& G, const int wo[3], const float o[3], const float d[3], uint32_t flags, GlslHitInfo& H) {
    GlslCast C;
    glsl_cast(S, G, wo, o, d, flags, C);
    if (!C.capped) H.iters = C.iters;  // :189,199 / :223,237; the fall-through `return false` after the loop assigns nothing
    if (!C.hit) return false;
    H.mat = __ldg(&S.palette[voxel_palette_id(S, C.p[0], C.p[1], C.p[2])]).x;
#pragma unroll
    for (int a = 0; a < 3; a++) H.pos[a] = C.cur[a], H.nrm[a] = (float)glsl_normal(C, a);
    return true;
}

__device__ __forceinline__ void glsl_normalize(float v[3]) {
    const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2]));
    const float k = __fdiv_rn(1.0f, __fsqrt_rn(len2));
    v[0] = __fmul_rn(v[0], k), v[1] = __fmul_rn(v[1], k), v[2] = __fmul_rn(v[2], k);
}
// getMaterialColor / getMaterialEmission, VoxelMap.glsl:78-84
__device__ __forceinline__ void glsl_material_color(uint32_t md, float c[3]) {
    c[0] = __fmul_rn((float)((md >> 11) & 31u), __fdiv_rn(1.0f, 31.0f));
    c[1] = __fmul_rn((float)((md >> 5) & 63u), __fdiv_rn(1.0f, 63.0f));
    c[2] = __fmul_rn((float)(md & 31u), __fdiv_rn(1.0f, 31.0f));
#pragma unroll
    for (int a = 0; a < 3; a++) c[a] = __fmul_rn(c[a], c[a]);
}
__device__ __forceinline__ float glsl_material_emission(uint32_t md) { return __half2float(__ushort_as_half((unsigned short)(md >> 16))); }
// getSkyColor, VoxelRender.comp:15-17 (stand-in sampler, see the header):
// the raw R11G11B10F texel of the context's cube at level 0 (ProjectCubemap + nearest), times 5, capped at 50000
__device__ __forceinline__ void glsl_sky(const FrameParams& F, const float d[3], float c[3]) {
    if (F.sky == nullptr) {
        c[0] = c[1] = c[2] = 0.0f;
        return;
    }
    float r3, g3, b3;
    sky_sample(F, d[0], d[1], d[2], 0u, r3, g3, b3);
    // sky_sample multiplies the unpacked texel by 3 (the CPU renderer's exposure); a R11G11B10F value times 3 is exact in fp32 (<= 6 + 2
    // significant bits), so dividing by 3 gives the texel back exactly
    c[0] = glsl_min(__fmul_rn(__fdiv_rn(r3, 3.0f), 5.0f), 50000.0f);
    c[1] = glsl_min(__fmul_rn(__fdiv_rn(g3, 3.0f), 5.0f), 50000.0f);
    c[2] = glsl_min(__fmul_rn(__fdiv_rn(b3, 3.0f), 5.0f), 50000.0f);
}
// random_dir, RandomGen.glsl:38-48 with blueNoise :27-36 (per pixel; the CPU renderer reads the same texture per 4x4 tile)
__device__ __forceinline__ void glsl_random_dir(const FrameParams& F, uint32_t x, uint32_t y, uint32_t i, float r[3]) {
    const uint32_t px = (x + F.bn_off[i][0]) & 127u, py = ((y + F.bn_off[i][1]) & 127u) + (F.frame_no & 63u) * 128u;
    const uint16_t t = __ldg(reinterpret_cast<const uint16_t*>(F.bn) + (size_t)py * 128u + px);
    const float nx = __fmul_rn(__fadd_rn((float)(t & 255u), 0.5f), 1.0f / 256.0f), ny = __fmul_rn(__fadd_rn((float)(t >> 8), 0.5f), 1.0f / 256.0f);
    const float yy = __fsub_rn(__fmul_rn(nx, 2.0f), 1.0f);
    float s, c;
    sincos_2pi(ny, s, c);  // a = n.y * 2 pi; sin(a), cos(a)
    const float sy = __fsqrt_rn(__fsub_rn(1.0f, __fmul_rn(yy, yy)));
    r[0] = __fmul_rn(s, sy), r[1] = yy, r[2] = __fmul_rn(c, sy);
}
__device__ __forceinline__ uint32_t glsl_unorm8(float v) { return (uint32_t)(int)rintf(__fmul_rn(glsl_min(glsl_max(v, 0.0f), 1.0f), 255.0f)); }

// main(), VoxelRender.comp:29-93
__device__ inline void glsl_frame_pixel(const DevScene& S, const GlslScene& G, const FrameParams& F, uint32_t cast_flags, uint32_t x, uint32_t y, PixelOut& P) {
    const int wo[3] = {F.W.wx, F.W.wy, F.W.wz};
    const float* m = F.inv_proj;
    float pos[3], dir[3];
    {  // getPrimaryRay, :20-25
        const float fx = (float)(int)x, fy = (float)(int)y;
        float nr[4], fr[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            nr[k] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[k], fx), __fmul_rn(m[4 + k], fy)), __fmul_rn(m[8 + k], 0.0f)), __fmul_rn(m[12 + k], 1.0f));
            fr[k] = __fadd_rn(nr[k], m[8 + k]);
        }
        const float in = __fdiv_rn(1.0f, nr[3]), iff = __fdiv_rn(1.0f, fr[3]);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            pos[a] = __fadd_rn(__fmul_rn(nr[a], in), F.frac[a]);
            dir[a] = __fmul_rn(fr[a], iff);
        }
        glsl_normalize(dir);
    }
    float albedo[3], irr[3], nrm[3] = {0.0f, 0.0f, 0.0f}, depth = -1.0f;
    GlslHitInfo hit;
    hit.iters = 0, hit.mat = 0;
    hit.pos[0] = hit.pos[1] = hit.pos[2] = 0.0f;
    hit.nrm[0] = hit.nrm[1] = hit.nrm[2] = 0.0f;
    const uint32_t fine = cast_flags & VRT_GLSL_ANISOTROPIC, coarse = fine | VRT_GLSL_COARSE;
    if (glsl_cast_info(S, G, wo, pos, dir, fine, hit)) {
        glsl_material_color(hit.mat, albedo);
        nrm[0] = hit.nrm[0], nrm[1] = hit.nrm[1], nrm[2] = hit.nrm[2];
        {  // :43-44, u_ProjMat * vec4(hit.pos / 16, 1)
            const float hx = __fmul_rn(hit.pos[0], 0.0625f), hy = __fmul_rn(hit.pos[1], 0.0625f), hz = __fmul_rn(hit.pos[2], 0.0625f);
            const float* q = F.proj;
            const float pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q[2], hx), __fmul_rn(q[6], hy)), __fmul_rn(q[10], hz)), __fmul_rn(q[14], 1.0f));
            const float pw = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q[3], hx), __fmul_rn(q[7], hy)), __fmul_rn(q[11], hz)), __fmul_rn(q[15], 1.0f));
            depth = __fdiv_rn(pz, pw);
        }
        const float em0 = glsl_material_emission(hit.mat);
        float thr[3] = {1.0f, 1.0f, 1.0f};
#pragma unroll
        for (int a = 0; a < 3; a++) irr[a] = F.bounces == 0 ? albedo[a] : __fmul_rn(albedo[a], em0);  // :46,49
        float sun[3] = {0.3f, 0.9f, -0.28f};  // :51-52
        glsl_normalize(sun);
        const float sun_intensity = 5.0f;
        const float sun_col[3] = {__fmul_rn(1.2f, sun_intensity), __fmul_rn(1.1f, sun_intensity), __fmul_rn(1.0f, sun_intensity)};
        GlslHitInfo sun_hit = hit;
        if (F.bounces != 0) {  // :55-61
            const float so[3] = {__fadd_rn(hit.pos[0], __fmul_rn(hit.nrm[0], 0.01f)), __fadd_rn(hit.pos[1], __fmul_rn(hit.nrm[1], 0.01f)),
                                 __fadd_rn(hit.pos[2], __fmul_rn(hit.nrm[2], 0.01f))};
            if (!glsl_cast_info(S, G, wo, so, sun, coarse, sun_hit)) {
#pragma unroll
                for (int a = 0; a < 3; a++) irr[a] = __fadd_rn(irr[a], sun_col[a]);
            } else {
#pragma unroll
                for (int a = 0; a < 3; a++) thr[a] = __fmul_rn(thr[a], 0.5f);
            }
        }
        for (uint32_t i = 0; i < F.bounces; i++) {  // :63-88
            float rnd[3];
            glsl_random_dir(F, x, y, i, rnd);
#pragma unroll
            for (int a = 0; a < 3; a++) {
                pos[a] = __fadd_rn(hit.pos[a], __fmul_rn(hit.nrm[a], 0.01f));
                dir[a] = __fadd_rn(hit.nrm[a], rnd[a]);
            }
            glsl_normalize(dir);
            if (!glsl_cast_info(S, G, wo, pos, dir, coarse, hit)) {
                float sky[3];
                glsl_sky(F, dir, sky);
#pragma unroll
                for (int a = 0; a < 3; a++) irr[a] = __fadd_rn(irr[a], __fmul_rn(thr[a], sky[a]));
                break;
            }
            float col[3];
            glsl_material_color(hit.mat, col);
#pragma unroll
            for (int a = 0; a < 3; a++) thr[a] = __fmul_rn(thr[a], col[a]);
            float emission = glsl_material_emission(hit.mat);
            if (i < 2u) {
                const float so[3] = {__fadd_rn(hit.pos[0], __fmul_rn(hit.nrm[0], 0.01f)), __fadd_rn(hit.pos[1], __fmul_rn(hit.nrm[1], 0.01f)),
                                     __fadd_rn(hit.pos[2], __fmul_rn(hit.nrm[2], 0.01f))};
                if (!glsl_cast_info(S, G, wo, so, sun, coarse, sun_hit)) {
#pragma unroll
                    for (int a = 0; a < 3; a++) thr[a] = __fmul_rn(thr[a], sun_col[a]);
                    emission = __fadd_rn(emission, sun_intensity);
                }
            }
#pragma unroll
            for (int a = 0; a < 3; a++) irr[a] = __fadd_rn(irr[a], __fmul_rn(thr[a], emission));
        }
    } else {  // :89-92
        glsl_sky(F, dir, irr);
        albedo[0] = albedo[1] = albedo[2] = 1.0f;
    }
    // imageStore x3, :93-95 with packGNormal (GBuffer.glsl:17-20)
    uint32_t code = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) code |= (uint32_t)(int)glsl_min(glsl_max(__fadd_rn(nrm[a], 1.0f), 0.0f), 3.0f) << (2 * a);
    P.albedo = glsl_unorm8(albedo[0]) | (glsl_unorm8(albedo[1]) << 8) | (glsl_unorm8(albedo[2]) << 16) | (code << 24);
    P.depth = depth;
    P.irr_rg = f2h_bits(irr[0]) | (f2h_bits(irr[1]) << 16);
    P.irr_bx = f2h_bits(irr[2]) | (f2h_bits((float)hit.iters) << 16);
}

template <bool ROWS>
__global__ void __launch_bounds__(VRT_RENDER_THREADS) k_render_glsl(const __grid_constant__ DevScene S, const __grid_constant__ GlslScene G,
                                                                    const __grid_constant__ FrameParams F, uint32_t cast_flags) {
    const uint32_t work = F.work_offset + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (work >= F.n_work) return;
    uint32_t x0, y0;
    if (!warp_tile_origin<ROWS>(F, work, x0, y0)) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t x = x0 + ((lane >> 4) << 2) + (lane & 3u), y = y0 + ((lane >> 2) & 3u);
    if (x >= F.width || y >= F.height) return;
    PixelOut P;
    glsl_frame_pixel(S, G, F, cast_flags, x, y, P);
    store_pixel(F, x, y, P);
}

}  // namespace vrt
