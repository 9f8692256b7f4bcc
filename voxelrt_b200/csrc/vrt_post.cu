// Image-space step after the traversal path (include/voxelrt_b200_post.h): the reference's
// GBuffer (src/VoxelRT/GBuffer.h) — CopyTiledFramebuffer.comp, Denoise/Reproject.comp, Denoise/Filter.comp,
// GBufferBlit.frag — as sm_100a kernels.  HBM/L2-bound stencil work, no tensor cores.
//
// Layout: where the reference keeps one texture per quantity (rgba16f irradiance, r32f depth, rgba8
// albedo+normal), every filter tap here is ONE 16-byte record {f16 r,g,b,variance; f32 depth; u32 albedo|normal<<24}:
// the three texels a tap needs arrive in a single LDG.128, and the "previous frame" set that Reproject samples
// (PrevIrradianceTex + PrevDepthTex + PrevAlbedoTex) is simply the record buffer the first à-trous pass wrote.  The
// blit is fused into the reprojection kernel (each thread unpacks its own pixel from the tile framebuffer), so a frame
// is 1 + 1 + N + 1 launches.  Arithmetic is spelled operation by operation (the library is built with -fmad=false) in
// the canonical order the parity oracle defines; exp/log are the same polynomial forms.
// (tests/native/emu_post.cpp compiles the device half of this file for the HOST behind a small shim, VRT_HOST_EMULATION, so
// that the CPU test tier checks these very kernels against the oracle without a GPU.)
#ifndef VRT_HOST_EMULATION
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#endif

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/voxelrt_b200_post.h"

namespace vrtpost {

struct Mat4 {
    float m[16];
};

// ---- arithmetic -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ float h2f(uint32_t bits16) { return __half2float(__ushort_as_half((unsigned short)bits16)); }
__device__ __forceinline__ uint32_t f2h(float f) {  // RNE, one canonical NaN
    if (f != f) return 0x7E00u;
    return (uint32_t)__half_as_ushort(__float2half_rn(f));
}
__device__ __forceinline__ float maxg(float a, float b) { return a < b ? b : a; }  // GLSL max: NaN keeps a
__device__ __forceinline__ float ming(float a, float b) { return b < a ? b : a; }

// exp(x) by range reduction + degree-6 Horner (see oracle header for the definition this must match bit for bit)
__device__ __forceinline__ float post_exp(float x) {
    if (x != x) return x;
    if (!(x > -87.0f)) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    const float n = rintf(__fmul_rn(x, 1.44269502f));
    float r = __fmaf_rn(n, -0.693145752f, x);
    r = __fmaf_rn(n, -1.42860677e-06f, r);
    float p = 1.38888892e-03f;
    p = __fmaf_rn(p, r, 8.33333377e-03f);
    p = __fmaf_rn(p, r, 4.16666679e-02f);
    p = __fmaf_rn(p, r, 1.66666672e-01f);
    p = __fmaf_rn(p, r, 0.5f);
    p = __fmaf_rn(p, r, 1.0f);
    p = __fmaf_rn(p, r, 1.0f);
    return __fmul_rn(p, __uint_as_float((uint32_t)((int)n + 127) << 23));
}
__device__ __forceinline__ float post_log(float x) {
    const uint32_t ix = __float_as_uint(x);
    int e = (int)(ix >> 23) - 127;
    float m = __uint_as_float((ix & 0x7FFFFFu) | 0x3F800000u);
    if (m > 1.41421354f) {
        m = __fmul_rn(m, 0.5f);
        e += 1;
    }
    const float s = __fdiv_rn(__fsub_rn(m, 1.0f), __fadd_rn(m, 1.0f));
    const float s2 = __fmul_rn(s, s);
    float p = 0.111111112f;
    p = __fmaf_rn(p, s2, 0.142857149f);
    p = __fmaf_rn(p, s2, 0.2f);
    p = __fmaf_rn(p, s2, 0.333333343f);
    p = __fmaf_rn(p, s2, 1.0f);
    const float lm = __fmul_rn(__fmul_rn(2.0f, s), p);
    return __fmaf_rn((float)e, 0.693147182f, lm);
}
__device__ __forceinline__ float pow045(float c) {
    if (!(c >= 1.17549435e-38f)) return 0.0f;
    return post_exp(__fmul_rn(0.45f, post_log(c)));
}
__device__ __forceinline__ float pow128(float x) {
#pragma unroll
    for (int i = 0; i < 7; i++) x = __fmul_rn(x, x);
    return x;
}
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}
// dot(color, vec3(0.299, 0.587, 0.114)) as a multiply and two FMAs (the contraction GPU GLSL compilers apply; oracle: same)
__device__ __forceinline__ float luminance(float r, float g, float b) { return __fmaf_rn(b, 0.114f, __fmaf_rn(g, 0.587f, __fmul_rn(r, 0.299f))); }
// GLSL mat4 * vec4, accumulated left to right
__device__ __forceinline__ float4 mat_vec(const Mat4& M, float x, float y, float z, float w) {
    float r[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
        r[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M.m[i], x), __fmul_rn(M.m[4 + i], y)), __fmul_rn(M.m[8 + i], z)),
                         __fmul_rn(M.m[12 + i], w));
    return make_float4(r[0], r[1], r[2], r[3]);
}
// getWorldPos, Reproject.comp:13-16
__device__ __forceinline__ float3 world_pos(const Mat4& inv, int sx, int sy, float depth) {
    const float4 v = mat_vec(inv, (float)sx, (float)sy, depth, 1.0f);
    const float s = __fdiv_rn(16.0f, v.w);
    return make_float3(__fmul_rn(v.x, s), __fmul_rn(v.y, s), __fmul_rn(v.z, s));
}
// unpackGNormal (GBuffer.glsl:15-17) of the alpha byte
__device__ __forceinline__ float3 unpack_normal(uint32_t albedo_normal) {
    const uint32_t a = albedo_normal >> 24;
    return make_float3((float)(a & 3u) - 1.0f, (float)((a >> 2) & 3u) - 1.0f, (float)((a >> 4) & 3u) - 1.0f);
}

// record accessors
struct Rec {
    float r, g, b, var, depth;
    uint32_t albedo;
};
__device__ __forceinline__ Rec unpack(const uint4 q) {
    Rec c;
    c.r = h2f(q.x & 0xFFFFu);
    c.g = h2f(q.x >> 16);
    c.b = h2f(q.y & 0xFFFFu);
    c.var = h2f(q.y >> 16);
    c.depth = __uint_as_float(q.z);
    c.albedo = q.w;
    return c;
}
__device__ __forceinline__ uint4 pack(float r, float g, float b, float var, float depth, uint32_t albedo) {
    return make_uint4(f2h(r) | (f2h(g) << 16), f2h(b) | (f2h(var) << 16), __float_as_uint(depth), albedo);
}

// CopyTiledFramebuffer.comp:10-37 for one pixel of the 4x4-tile framebuffer (VrtTile)
__device__ __forceinline__ uint4 load_tile_pixel(const uint32_t* __restrict__ tiles, int x, int y, int w) {
    const uint32_t off = (((uint32_t)x >> 2) + ((uint32_t)y >> 2) * ((uint32_t)w >> 2)) * 64u + ((uint32_t)x & 3u) + (((uint32_t)y & 3u) << 2);
    const uint32_t a = __ldg(tiles + off), d = __ldg(tiles + off + 16), rg = __ldg(tiles + off + 32), bx = __ldg(tiles + off + 48);
    uint32_t rgb = a & 0xFFFFFFu;
    if (__uint_as_float(d) < 0.0f) rgb = 0xFFFFFFu;  // :31
    return make_uint4(rg, bx & 0xFFFFu, d, rgb | (((a >> 24) & 0x3Fu) << 24));
}

// albedo + normal word only (present): two of the four tile fields
__device__ __forceinline__ uint32_t load_tile_albedo(const uint32_t* __restrict__ tiles, int x, int y, int w) {
    const uint32_t off = (((uint32_t)x >> 2) + ((uint32_t)y >> 2) * ((uint32_t)w >> 2)) * 64u + ((uint32_t)x & 3u) + (((uint32_t)y & 3u) << 2);
    const uint32_t a = __ldg(tiles + off), d = __ldg(tiles + off + 16);
    uint32_t rgb = a & 0xFFFFFFu;
    if (__uint_as_float(d) < 0.0f) rgb = 0xFFFFFFu;  // :31
    return rgb | (((a >> 24) & 0x3Fu) << 24);
}

struct FrameParams {
    Mat4 cur_inv, hist_proj, hist_inv;
    float delta[3];
    int w, h;
    int reset;
};

// ---- blit + Reproject.comp ------------------------------------------------------------------------------------------
// One thread per pixel, 32x8 blocks.  prev = the record buffer holding last frame's PrevIrradiance/Depth/Albedo.
__global__ void __launch_bounds__(256) k_reproject(const uint32_t* __restrict__ tiles, const uint4* __restrict__ prev,
                                                   const uint32_t* __restrict__ prev_moments, const uint8_t* __restrict__ hist_in,
                                                   uint4* __restrict__ out, uint32_t* __restrict__ moments, uint8_t* __restrict__ hist_out,
                                                   const __grid_constant__ FrameParams P) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= P.w || y >= P.h) return;
    const size_t i = (size_t)y * P.w + x;
    const uint4 q = load_tile_pixel(tiles, x, y, P.w);
    const Rec cur = unpack(q);
    bool ok = false;
    float nr = cur.r, ng = cur.g, nb = cur.b, nvar = 0.0f, nm0 = 0.0f, nm1 = 0.0f;
    uint32_t hl = 0;
    if (!(cur.depth <= 0.0f)) {  // :45 `if (depth <= 0) return false`
        const float3 wp = world_pos(P.cur_inv, x, y, cur.depth);
        const float4 ndc = mat_vec(P.hist_proj, __fadd_rn(wp.x, P.delta[0]), __fadd_rn(wp.y, P.delta[1]), __fadd_rn(wp.z, P.delta[2]), 1.0f);
        const float px = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fdiv_rn(ndc.x, ndc.w), 0.5f), 0.5f), (float)P.w), 0.5f);
        const float py = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fdiv_rn(ndc.y, ndc.w), 0.5f), 0.5f), (float)P.h), 0.5f);
        if (px > -1.0e9f && px < 1.0e9f && py > -1.0e9f && py < 1.0e9f) {
            const int pxi = (int)px, pyi = (int)py;  // truncation, like ivec2()
            const float fx = __fsub_rn(px, floorf(px)), fy = __fsub_rn(py, floorf(py));
            if ((uint32_t)pxi < (uint32_t)P.w && (uint32_t)pyi < (uint32_t)P.h) {
                const float3 cn = unpack_normal(cur.albedo);
                float wsum = 0.0f, ir = 0.0f, ig = 0.0f, ib = 0.0f, m0 = 0.0f, m1 = 0.0f;
                hl = hist_in[i];
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    const int sx = pxi + (s & 1), sy = pyi + (s >> 1);
                    if ((uint32_t)sx >= (uint32_t)P.w || (uint32_t)sy >= (uint32_t)P.h) continue;
                    const size_t j = (size_t)sy * P.w + sx;
                    const Rec sp = unpack(__ldg(prev + j));
                    const float3 sn = unpack_normal(sp.albedo);
                    if (dot3(cn.x, cn.y, cn.z, sn.x, sn.y, sn.z) < 0.5f) continue;
                    if (sp.depth <= 0.0f) continue;
                    const float3 sw = world_pos(P.hist_inv, sx, sy, sp.depth);
                    const float dx = __fadd_rn(__fsub_rn(wp.x, sw.x), P.delta[0]), dy = __fadd_rn(__fsub_rn(wp.y, sw.y), P.delta[1]),
                                dz = __fadd_rn(__fsub_rn(wp.z, sw.z), P.delta[2]);
                    if (fabsf(dot3(dx, dy, dz, cn.x, cn.y, cn.z)) > 6.0f) continue;
                    const float wgt = __fmul_rn((s & 1) ? fx : __fsub_rn(1.0f, fx), (s >> 1) ? fy : __fsub_rn(1.0f, fy));
                    ir = __fadd_rn(ir, __fmul_rn(sp.r, wgt));
                    ig = __fadd_rn(ig, __fmul_rn(sp.g, wgt));
                    ib = __fadd_rn(ib, __fmul_rn(sp.b, wgt));
                    const uint32_t pm = __ldg(prev_moments + j);
                    m0 = __fadd_rn(m0, __fmul_rn(h2f(pm & 0xFFFFu), wgt));
                    m1 = __fadd_rn(m1, __fmul_rn(h2f(pm >> 16), wgt));
                    wsum = __fadd_rn(wsum, wgt);
                    hl = min(hl, (uint32_t)hist_in[j] + 1u);
                }
                if (!(wsum < 0.001f)) {
                    ir = __fdiv_rn(ir, wsum);
                    ig = __fdiv_rn(ig, wsum);
                    ib = __fdiv_rn(ib, wsum);
                    m0 = __fdiv_rn(m0, wsum);
                    m1 = __fdiv_rn(m1, wsum);
                    if (P.reset && hl > 6u) hl = 6u;
                    const float blend = __fdiv_rn(1.0f, (float)(hl + 1u));
                    const float keep = __fsub_rn(1.0f, blend);
                    nr = __fadd_rn(__fmul_rn(ir, keep), __fmul_rn(cur.r, blend));
                    ng = __fadd_rn(__fmul_rn(ig, keep), __fmul_rn(cur.g, blend));
                    nb = __fadd_rn(__fmul_rn(ib, keep), __fmul_rn(cur.b, blend));
                    const float luma = luminance(nr, ng, nb);
                    const float mb = maxg(0.5f, blend), mk = __fsub_rn(1.0f, mb);
                    nm0 = __fadd_rn(__fmul_rn(m0, mk), __fmul_rn(luma, mb));
                    nm1 = __fadd_rn(__fmul_rn(m1, mk), __fmul_rn(__fmul_rn(luma, luma), mb));
                    nvar = maxg(0.0f, __fsub_rn(nm1, __fmul_rn(nm0, nm0)));
                    hl = min(hl + 1u, 64u);
                    ok = true;
                }
            }
        }
    }
    if (ok) {
        out[i] = pack(nr, ng, nb, nvar, cur.depth, cur.albedo);
        moments[i] = f2h(nm0) | (f2h(nm1) << 16);
        hist_out[i] = (uint8_t)hl;
    } else {  // main(), :104-108: irradiance stays what the blit wrote (w = 0)
        out[i] = q;
        moments[i] = 0u;
        hist_out[i] = 0;
    }
}

// ---- filter taps ----------------------------------------------------------------------------------------------------------
// The à-trous passes are INSTRUCTION-bound, not memory-bound (first B200 run: 0.8 ms per 4K pass for 32 B/px of HBM traffic,
// ~90 instructions per tap), so the tap is spelled for the SM's pipes while producing the oracle's bits:
//   * pow(clamp(dot(n, cn), .001, 1), 128): normals are integer vectors, so the dot product is an integer, the clamp gives 1
//     (dot >= 1) or 0.001, and seven squarings give exactly 1.0f or 0.0f — computed as an integer dot product and a select;
//   * exp(): rint() through the 1.5*2^23 magic add (FADD on the FMA pipe instead of FRND/F2I on the quarter-rate pipe), the
//     2^n scaling as an integer add on the exponent field (exact, like the multiplication it replaces);
//   * the two quotients are a * (1/b) by definition (oracle header); 1/lumaPhi is per pixel, 1/(length + 0.001) per tap and
//     computed on the host (IEEE sqrt and division), so no division is left in the loop;
//   * kernels are templated on the pass number: tap offsets and kernel weights are immediates, the 25 record addresses are
//     five row pointers plus constant offsets; CTAs whose taps all fall inside the image skip the bounds tests.
__device__ __forceinline__ float post_exp_neg(float x) {  // == post_exp(x), restated (x <= 0 or NaN in the filters)
    if (x != x) return x;
    if (!(x > -87.0f)) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    const float tm = __fadd_rn(__fmul_rn(x, 1.44269502f), 12582912.0f);  // low mantissa bits = rint(t), two's complement
    const float n = __fsub_rn(tm, 12582912.0f);
    float r = __fmaf_rn(n, -0.693145752f, x);
    r = __fmaf_rn(n, -1.42860677e-06f, r);
    float p = 1.38888892e-03f;
    p = __fmaf_rn(p, r, 8.33333377e-03f);
    p = __fmaf_rn(p, r, 4.16666679e-02f);
    p = __fmaf_rn(p, r, 1.66666672e-01f);
    p = __fmaf_rn(p, r, 0.5f);
    p = __fmaf_rn(p, r, 1.0f);
    p = __fmaf_rn(p, r, 1.0f);
    return __int_as_float(__float_as_int(p) + ((__float_as_int(tm) - 0x4B400000) << 23));  // p in [0.7, 1.42), n >= -126: stays normal
}
// integer components of unpackGNormal: each in {-1, 0, 1, 2}
struct INormal {
    int x, y, z;
};
__device__ __forceinline__ INormal inormal(uint32_t albedo_normal) {
    const uint32_t a = albedo_normal >> 24;
    return {(int)(a & 3u) - 1, (int)((a >> 2) & 3u) - 1, (int)((a >> 4) & 3u) - 1};
}
// exp(-(w_luma + w_depth)) * w_normal.  Almost every tap has the centre's normal code: then dot(n, cn) = |cn|^2 and w_normal is
// the per-pixel constant `self`; only other codes take the integer dot product.
__device__ __forceinline__ float tap_weight(float w_luma, float w_depth, uint32_t tap_albedo, uint32_t centre_albedo, const INormal cn, float self) {
    float w_normal = self;
    if ((tap_albedo ^ centre_albedo) >> 24) {
        const INormal n = inormal(tap_albedo);
        w_normal = (n.x * cn.x + n.y * cn.y + n.z * cn.z) >= 1 ? 1.0f : 0.0f;
    }
    return __fmul_rn(post_exp_neg(-__fadd_rn(w_luma, w_depth)), w_normal);
}
__device__ __forceinline__ float self_weight(const INormal cn) { return (cn.x * cn.x + cn.y * cn.y + cn.z * cn.z) >= 1 ? 1.0f : 0.0f; }

struct TapRcp {
    float v[49];  // 1 / (length(offset) + 0.001) per tap, row-major over the (2R+1)^2 window
};
// 1 / (length(offset) + 0.001) for the (2r+1)^2 taps at dilation d: IEEE sqrt, add, divide, as the oracle computes them per tap
inline TapRcp tap_rcp(int r, int d) {
    TapRcp t{};
    const int n = 2 * r + 1;
    for (int ky = -r; ky <= r; ky++)
        for (int kx = -r; kx <= r; kx++) {
            const float fx = (float)(kx * d), fy = (float)(ky * d);
            volatile float sq = fx * fx;  // volatile: no contraction of fx*fx + fy*fy by the host compiler
            volatile float sq2 = fy * fy;
            volatile float len = std::sqrt((float)(sq + sq2));
            volatile float den = len + 0.001f;
            t.v[(ky + r) * n + (kx + r)] = 1.0f / den;
        }
    return t;
}


// ---- Filter.comp pass -1: varianceEstim (:17-68).  in = IrradianceTex records, io_temp = TempIrradianceTex records -------
__global__ void __launch_bounds__(256) k_variance(const uint4* __restrict__ in, const uint8_t* __restrict__ hist, uint4* io_temp, int w, int h,
                                                  const __grid_constant__ TapRcp rcp) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t i = (size_t)y * w + x;
    const uint4 cq = in[i];
    const Rec c = unpack(cq);
    const uint32_t hl = hist[i];
    if (hl > 4u || c.depth < 0.0f) {
        io_temp[i] = cq;
        return;
    }
    const Rec stale = unpack(io_temp[i]);  // :26 reads the centre from the TEMP texture
    const float cl = luminance(stale.r, stale.g, stale.b);
    const INormal cn = inormal(c.albedo);
    const float self = self_weight(cn);
    float sr = 0.0f, sg = 0.0f, sb = 0.0f, s0 = 0.0f, s1 = 0.0f, wsum = 0.0f;
#pragma unroll
    for (int ky = -3; ky <= 3; ky++) {
        const int sy = y + ky;
        if ((uint32_t)sy >= (uint32_t)h) continue;
        const uint4* row = in + (size_t)sy * w + x;
#pragma unroll
        for (int kx = -3; kx <= 3; kx++) {
            if ((uint32_t)(x + kx) >= (uint32_t)w) continue;
            const Rec t = unpack(__ldg(row + kx));
            const float l = luminance(t.r, t.g, t.b);
            const float w_luma = __fmul_rn(fabsf(__fsub_rn(l, cl)), 0.1f);  // 1/lumaPhi, lumaPhi = 10
            const float w_depth = __fmul_rn(fabsf(__fsub_rn(c.depth, t.depth)), rcp.v[(ky + 3) * 7 + (kx + 3)]);
            const float wgt = tap_weight(w_luma, w_depth, t.albedo, c.albedo, cn, self);
            sr = __fmaf_rn(t.r, wgt, sr);  // sums are FMAs (oracle: same)
            sg = __fmaf_rn(t.g, wgt, sg);
            sb = __fmaf_rn(t.b, wgt, sb);
            s0 = __fmaf_rn(l, wgt, s0);
            s1 = __fmaf_rn(__fmul_rn(l, l), wgt, s1);
            wsum = __fadd_rn(wsum, wgt);
        }
    }
    wsum = maxg(wsum, 0.001f);
    s0 = __fdiv_rn(s0, wsum);
    s1 = __fdiv_rn(s1, wsum);
    float var = maxg(0.0f, __fsub_rn(s1, __fmul_rn(s0, s0)));
    var = __fmul_rn(var, __fmul_rn(__fsub_rn(4.0f, (float)hl), 3.0f));
    io_temp[i] = pack(__fdiv_rn(sr, wsum), __fdiv_rn(sg, wsum), __fdiv_rn(sb, wsum), var, c.depth, c.albedo);
}

// ---- Filter.comp pass >= 0: svgfAtrous + getFilteredVariance (:70-135) -------------------------------------------------
template <int PASS, bool INTERIOR>
__device__ __forceinline__ void atrous_pixel(const uint4* __restrict__ in, uint4* __restrict__ out, int x, int y, int w, int h, const TapRcp& rcp) {
    constexpr int D = 1 << PASS;
    const size_t i = (size_t)y * w + x;
    const uint4* centre = in + i;
    const uint4 cq = __ldg(centre);
    const Rec c = unpack(cq);
    if (c.depth < 0.0f) {
        out[i] = cq;
        return;
    }
    float cv = 0.0f;
#pragma unroll
    for (int ky = -1; ky <= 1; ky++)
#pragma unroll
        for (int kx = -1; kx <= 1; kx++) {
            float v = 0.0f;  // imageLoad outside the image returns 0
            if ((uint32_t)(x + kx) < (uint32_t)w && (uint32_t)(y + ky) < (uint32_t)h) v = h2f(__ldg(&(centre + (ptrdiff_t)ky * w + kx)->y) >> 16);
            const float k = (kx == 0 ? 0.25f : 0.125f) * (ky == 0 ? 1.0f : 0.5f);  // kernel[|kx|][|ky|] = {1/4,1/8;1/8,1/16}
            cv = __fmaf_rn(v, k, cv);
        }
    const INormal cn = inormal(c.albedo);
    const float self = self_weight(cn);
    const float cl = luminance(c.r, c.g, c.b);
    const float inv_phi = __fdiv_rn(1.0f, __fmul_rn(__fsqrt_rn(maxg(0.0001f, cv)), 4.0f));
    float sr = c.r, sg = c.g, sb = c.b, sv = c.var, wsum = 1.0f;
#pragma unroll
    for (int ky = -2; ky <= 2; ky++) {
        if (!INTERIOR && (uint32_t)(y + ky * D) >= (uint32_t)h) continue;
        const uint4* row = centre + (ptrdiff_t)(ky * D) * w;
#pragma unroll
        for (int kx = -2; kx <= 2; kx++) {
            if (kx == 0 && ky == 0) continue;
            if (!INTERIOR && (uint32_t)(x + kx * D) >= (uint32_t)w) continue;
            const Rec t = unpack(__ldg(row + kx * D));
            const float w_luma = __fmul_rn(fabsf(__fsub_rn(luminance(t.r, t.g, t.b), cl)), inv_phi);
            const float w_depth = __fmul_rn(fabsf(__fsub_rn(c.depth, t.depth)), rcp.v[(ky + 2) * 5 + (kx + 2)]);
            const float kxw = kx == 0 ? 0.375f : ((kx == 1 || kx == -1) ? 0.25f : 0.0625f);
            const float kyw = ky == 0 ? 0.375f : ((ky == 1 || ky == -1) ? 0.25f : 0.0625f);
            const float wgt = __fmul_rn(kxw * kyw, tap_weight(w_luma, w_depth, t.albedo, c.albedo, cn, self));
            sr = __fmaf_rn(t.r, wgt, sr);
            sg = __fmaf_rn(t.g, wgt, sg);
            sb = __fmaf_rn(t.b, wgt, sb);
            sv = __fmaf_rn(t.var, __fmul_rn(wgt, wgt), sv);
            wsum = __fadd_rn(wsum, wgt);
        }
    }
    if (wsum < 0.001f) wsum = 0.001f;
    out[i] = pack(__fdiv_rn(sr, wsum), __fdiv_rn(sg, wsum), __fdiv_rn(sb, wsum), __fdiv_rn(sv, __fmul_rn(wsum, wsum)), c.depth, c.albedo);
}

template <int PASS>
__global__ void __launch_bounds__(256) k_atrous(const uint4* __restrict__ in, uint4* __restrict__ out, int w, int h, const __grid_constant__ TapRcp rcp) {
    constexpr int R = 2 << PASS;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    const bool interior = x0 >= R && y0 >= R && x0 + 31 + R < w && y0 + 7 + R < h;  // uniform per CTA
    if (interior) {
        atrous_pixel<PASS, true>(in, out, x, y, w, h, rcp);
    } else {
        if (x >= w || y >= h) return;
        atrous_pixel<PASS, false>(in, out, x, y, w, h, rcp);
    }
}

// ---- GBufferBlit.frag:8-46 ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float aces(float v) {
    v = __fmul_rn(v, 0.6f);
    const float num = __fmul_rn(v, __fadd_rn(__fmul_rn(2.51f, v), 0.03f));
    const float den = __fadd_rn(__fmul_rn(v, __fadd_rn(__fmul_rn(2.43f, v), 0.59f)), 0.14f);
    const float r = __fdiv_rn(num, den);
    return r < 0.0f ? 0.0f : (r > 1.0f ? 1.0f : r);
}
__device__ __forceinline__ uint32_t unorm8(float c) {
    if (!(c > 0.0f)) return 0u;
    if (c > 1.0f) c = 1.0f;
    return (uint32_t)__fadd_rn(__fmul_rn(c, 255.0f), 0.5f);
}
// albedo and normal come from the tile framebuffer (always the CURRENT frame, like u_AlbedoNormalTex); the irradiance from
// whichever record buffer carries the name IrradianceTex after the pass rotation (GBuffer.h:100-128)
__global__ void __launch_bounds__(256) k_present(const uint32_t* __restrict__ tiles, const uint4* __restrict__ irr, uint32_t* __restrict__ rgba,
                                                 int w, int h, int channel) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t i = (size_t)y * w + x;
    const uint32_t a = load_tile_albedo(tiles, x, y, w);
    const Rec t = unpack(__ldg(irr + i));
    const float ar = __fdiv_rn((float)(a & 255u), 255.0f), ag = __fdiv_rn((float)((a >> 8) & 255u), 255.0f), ab = __fdiv_rn((float)((a >> 16) & 255u), 255.0f);
    float cr, cg, cb;
    if (channel == 1) {
        cr = ar, cg = ag, cb = ab;
    } else if (channel == 2) {
        cr = aces(t.r), cg = aces(t.g), cb = aces(t.b);
    } else if (channel == 3) {
        const float3 n = unpack_normal(a);
        cr = __fadd_rn(__fmul_rn(n.x, 0.5f), 0.5f), cg = __fadd_rn(__fmul_rn(n.y, 0.5f), 0.5f), cb = __fadd_rn(__fmul_rn(n.z, 0.5f), 0.5f);
    } else if (channel == 4) {
        const float it = t.var;
        if (it < 64.0f) {
            cr = cg = cb = __fdiv_rn(it, 64.0f);
        } else {
            const float tt = __fdiv_rn(__fsub_rn(it, 64.0f), 128.0f), k = __fsub_rn(1.0f, tt);
            cr = __fadd_rn(__fmul_rn(1.0f, k), __fmul_rn(1.0f, tt));
            cg = cb = __fadd_rn(__fmul_rn(1.0f, k), __fmul_rn(0.0f, tt));
        }
    } else if (channel == 5) {
        cr = cg = cb = __fmul_rn(__fsqrt_rn(t.var), 3.0f);
    } else {
        cr = pow045(aces(__fmul_rn(__fmul_rn(ar, t.r), 0.48f)));
        cg = pow045(aces(__fmul_rn(__fmul_rn(ag, t.g), 0.48f)));
        cb = pow045(aces(__fmul_rn(__fmul_rn(ab, t.b), 0.48f)));
    }
    rgba[i] = unorm8(cr) | (unorm8(cg) << 8) | (unorm8(cb) << 16) | 0xFF000000u;
}

// TraversalIters frames skip the denoiser (GBuffer.h:90) but the reference's blit still refreshes IrradianceTex, and SetCamera
// still makes this frame's depth / albedo the next frame's "previous" ones: write the blit into `irr`, and refresh the geometry
// half of the records that carry PrevIrradianceTex.
__global__ void __launch_bounds__(256) k_blit_only(const uint32_t* __restrict__ tiles, uint4* __restrict__ irr, uint4* __restrict__ prev, int w, int h) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t i = (size_t)y * w + x;
    const uint4 q = load_tile_pixel(tiles, x, y, w);
    irr[i] = q;
    prev[i].z = q.z;
    prev[i].w = q.w;
}

}  // namespace vrtpost

#ifndef VRT_HOST_EMULATION
using namespace vrtpost;

struct VrtGBuffer {
    int device = 0;
    uint32_t w = 0, h = 0;
    // names follow GBuffer.h:12-16; the three record buffers rotate exactly like Irradiance/PrevIrradiance/TempIrradiance
    uint4 *irr = nullptr, *prev_irr = nullptr, *temp_irr = nullptr;
    uint32_t *moments = nullptr, *prev_moments = nullptr;
    uint8_t *hist = nullptr, *hist_prev = nullptr;  // hist_prev = the values before this frame's Reproject dispatch
    Mat4 cur_proj{}, hist_proj{}, cur_inv{}, hist_inv{};
    double cur_pos[3] = {0, 0, 0}, hist_pos[3] = {0, 0, 0};
    bool have_camera = false, camera_set_for_frame = false;
    uint32_t reset_history = 0;
    uint32_t frame_no = 0, num_passes = 5, channel = 0;
    uint64_t last_launches = 0;
    // scratch of the host-pointer entry point
    uint32_t *d_tiles = nullptr, *d_rgba = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
};

namespace {
thread_local std::string g_gb_create_error;

int gfail(VrtGBuffer* g, int status, const std::string& msg) {
    if (g) g->err = msg;
    else g_gb_create_error = msg;
    return status;
}
#define GCU(call)                                                                                                    \
    do {                                                                                                             \
        cudaError_t e__ = (call);                                                                                    \
        if (e__ != cudaSuccess) return gfail(gb, VRT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

void free_planes(VrtGBuffer* g) {
    void* ps[] = {g->irr, g->prev_irr, g->temp_irr, g->moments, g->prev_moments, g->hist, g->hist_prev, g->d_tiles, g->d_rgba};
    for (void* p : ps)
        if (p) cudaFree(p);
    g->irr = g->prev_irr = g->temp_irr = nullptr;
    g->moments = g->prev_moments = nullptr;
    g->hist = g->hist_prev = nullptr;
    g->d_tiles = g->d_rgba = nullptr;
}

int alloc_planes(VrtGBuffer* gb, uint32_t w, uint32_t h) {
    free_planes(gb);
    const size_t n = (size_t)w * h;
    void** rec[] = {(void**)&gb->irr, (void**)&gb->prev_irr, (void**)&gb->temp_irr};
    for (void** p : rec) {
        GCU(cudaMalloc(p, n * 16));
        GCU(cudaMemsetAsync(*p, 0, n * 16, gb->stream));
    }
    void** mom[] = {(void**)&gb->moments, (void**)&gb->prev_moments};
    for (void** p : mom) {
        GCU(cudaMalloc(p, n * 4));
        GCU(cudaMemsetAsync(*p, 0, n * 4, gb->stream));
    }
    void** hs[] = {(void**)&gb->hist, (void**)&gb->hist_prev};
    for (void** p : hs) {
        GCU(cudaMalloc(p, n));
        GCU(cudaMemsetAsync(*p, 0, n, gb->stream));
    }
    GCU(cudaStreamSynchronize(gb->stream));
    gb->w = w;
    gb->h = h;
    return VRT_OK;
}
}  // namespace

extern "C" {

VRT_API int vrt_gbuffer_create(int32_t device, VrtGBuffer** out) {
    if (!out) return gfail(nullptr, VRT_ERR_INVALID, "vrt_gbuffer_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return gfail(nullptr, VRT_ERR_CUDA, std::string("vrt_gbuffer_create: no usable CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= count) return gfail(nullptr, VRT_ERR_INVALID, "vrt_gbuffer_create: device ordinal out of range");
    if (cudaSetDevice(device) != cudaSuccess) return gfail(nullptr, VRT_ERR_CUDA, "vrt_gbuffer_create: cudaSetDevice failed");
    VrtGBuffer* g = new VrtGBuffer();
    g->device = device;
    if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete g;
        return gfail(nullptr, VRT_ERR_CUDA, "vrt_gbuffer_create: cudaStreamCreate failed");
    }
    *out = g;
    return VRT_OK;
}

VRT_API void vrt_gbuffer_destroy(VrtGBuffer* gb) {
    if (!gb) return;
    cudaSetDevice(gb->device);
    cudaDeviceSynchronize();
    free_planes(gb);
    if (gb->stream) cudaStreamDestroy(gb->stream);
    delete gb;
}

VRT_API const char* vrt_gbuffer_last_error(const VrtGBuffer* gb) { return gb ? gb->err.c_str() : g_gb_create_error.c_str(); }

VRT_API int vrt_gbuffer_set_passes(VrtGBuffer* gb, uint32_t num_passes) {
    if (!gb) return VRT_ERR_INVALID;
    if (num_passes > 5) return gfail(gb, VRT_ERR_INVALID, "vrt_gbuffer_set_passes: 0..5 (GBuffer.h:23, CpuRenderer.cpp:480)");
    gb->num_passes = num_passes;
    return VRT_OK;
}

VRT_API int vrt_gbuffer_set_debug_channel(VrtGBuffer* gb, uint32_t channel) {
    if (!gb) return VRT_ERR_INVALID;
    if (channel > 5) return gfail(gb, VRT_ERR_INVALID, "vrt_gbuffer_set_debug_channel: 0..5 (GBuffer.h:8)");
    gb->channel = channel;
    return VRT_OK;
}

VRT_API int vrt_gbuffer_set_camera(VrtGBuffer* gb, const VrtGBufferCamera* cam) {
    if (!gb) return VRT_ERR_INVALID;
    if (!cam) return gfail(gb, VRT_ERR_INVALID, "vrt_gbuffer_set_camera: cam is NULL");
    if (cam->width == 0 || cam->height == 0 || (cam->width & 3u) || (cam->height & 3u) || cam->width > 32768u || cam->height > 32768u)
        return gfail(gb, VRT_ERR_INVALID, "vrt_gbuffer_set_camera: width/height must be positive multiples of 4 (CpuRenderer.cpp:419)");
    GCU(cudaSetDevice(gb->device));
    if (cam->width != gb->w || cam->height != gb->h) {  // GBuffer.h:32-47: new textures, history gone
        int st = alloc_planes(gb, cam->width, cam->height);
        if (st != VRT_OK) return st;
    }
    if (!gb->have_camera) {  // no history yet: start it at the current camera
        std::memcpy(gb->cur_proj.m, cam->proj, 64);
        std::memcpy(gb->cur_inv.m, cam->inv_proj, 64);
        std::memcpy(gb->cur_pos, cam->position, 24);
        gb->have_camera = true;
    }
    gb->hist_proj = gb->cur_proj;
    gb->hist_inv = gb->cur_inv;
    std::memcpy(gb->hist_pos, gb->cur_pos, 24);
    std::memcpy(gb->cur_proj.m, cam->proj, 64);
    std::memcpy(gb->cur_inv.m, cam->inv_proj, 64);
    std::memcpy(gb->cur_pos, cam->position, 24);
    std::swap(gb->moments, gb->prev_moments);  // GBuffer.h:55 (albedo / depth travel inside the records)
    gb->frame_no++;
    gb->reset_history = cam->reset_history;
    gb->camera_set_for_frame = true;
    return VRT_OK;
}

VRT_API int vrt_gbuffer_denoise_present_device(VrtGBuffer* gb, const void* d_tiles, uint32_t* d_out_rgba8, void* stream_) {
    if (!gb) return VRT_ERR_INVALID;
    if (!d_tiles || !d_out_rgba8) return gfail(gb, VRT_ERR_INVALID, "vrt_gbuffer_denoise_present: NULL buffer");
    if (!gb->camera_set_for_frame) return gfail(gb, VRT_ERR_STATE, "vrt_gbuffer_denoise_present: call vrt_gbuffer_set_camera first (CpuRenderer.cpp:423)");
    GCU(cudaSetDevice(gb->device));
    cudaStream_t s = (cudaStream_t)stream_;
    const uint32_t* tiles = (const uint32_t*)d_tiles;
    const int w = (int)gb->w, h = (int)gb->h;
    const dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8);
    uint64_t launches = 0;
    if (gb->channel != VRT_CHANNEL_TRAVERSAL_ITERS) {
        FrameParams P;
        P.cur_inv = gb->cur_inv;
        P.hist_proj = gb->hist_proj;
        P.hist_inv = gb->hist_inv;
        for (int k = 0; k < 3; k++) P.delta[k] = (float)(gb->cur_pos[k] - gb->hist_pos[k]);  // GBuffer.h:82
        P.w = w;
        P.h = h;
        P.reset = gb->reset_history ? 1 : 0;
        std::swap(gb->hist, gb->hist_prev);  // this dispatch reads the old lengths and writes every pixel of the new plane
        k_reproject<<<grid, block, 0, s>>>(tiles, gb->prev_irr, gb->prev_moments, gb->hist_prev, gb->irr, gb->moments, gb->hist, P);
        launches++;
        if (gb->num_passes > 0) {
            k_variance<<<grid, block, 0, s>>>(gb->irr, gb->hist, gb->temp_irr, w, h, tap_rcp(3, 1));
            launches++;
            for (uint32_t i = 0; i < gb->num_passes; i++) {  // GBuffer.h:103-121
                const uint4* in = i == 1 ? gb->prev_irr : (i % 2 == 0 ? gb->temp_irr : gb->irr);
                uint4* out = i % 2 == 0 ? gb->irr : gb->temp_irr;
                const TapRcp tr = tap_rcp(2, 1 << i);
                switch (i) {
                case 0: k_atrous<0><<<grid, block, 0, s>>>(in, out, w, h, tr); break;
                case 1: k_atrous<1><<<grid, block, 0, s>>>(in, out, w, h, tr); break;
                case 2: k_atrous<2><<<grid, block, 0, s>>>(in, out, w, h, tr); break;
                case 3: k_atrous<3><<<grid, block, 0, s>>>(in, out, w, h, tr); break;
                default: k_atrous<4><<<grid, block, 0, s>>>(in, out, w, h, tr); break;
                }
                launches++;
                if (i == 0) std::swap(gb->prev_irr, gb->irr);
            }
            if (gb->num_passes % 2 != 0) std::swap(gb->temp_irr, gb->irr);
        }
    } else {
        k_blit_only<<<grid, block, 0, s>>>(tiles, gb->irr, gb->prev_irr, w, h);
        launches++;
    }
    k_present<<<grid, block, 0, s>>>(tiles, gb->irr, d_out_rgba8, w, h, (int)gb->channel);
    launches++;
    if (gb->num_passes == 0) std::swap(gb->prev_irr, gb->irr);  // GBuffer.h:128-130
    GCU(cudaGetLastError());
    gb->last_launches = launches;
    gb->camera_set_for_frame = false;
    return VRT_OK;
}

VRT_API int vrt_gbuffer_denoise_present(VrtGBuffer* gb, const void* tiles, uint32_t* out_rgba8) {
    if (!gb) return VRT_ERR_INVALID;
    if (!tiles || !out_rgba8) return gfail(gb, VRT_ERR_INVALID, "vrt_gbuffer_denoise_present: NULL buffer");
    if (!gb->camera_set_for_frame) return gfail(gb, VRT_ERR_STATE, "vrt_gbuffer_denoise_present: call vrt_gbuffer_set_camera first (CpuRenderer.cpp:423)");
    GCU(cudaSetDevice(gb->device));
    const size_t n = (size_t)gb->w * gb->h;
    if (!gb->d_tiles) GCU(cudaMalloc((void**)&gb->d_tiles, n * 16));
    if (!gb->d_rgba) GCU(cudaMalloc((void**)&gb->d_rgba, n * 4));
    GCU(cudaMemcpyAsync(gb->d_tiles, tiles, n * 16, cudaMemcpyHostToDevice, gb->stream));
    int st = vrt_gbuffer_denoise_present_device(gb, gb->d_tiles, gb->d_rgba, gb->stream);
    if (st != VRT_OK) return st;
    GCU(cudaMemcpyAsync(out_rgba8, gb->d_rgba, n * 4, cudaMemcpyDeviceToHost, gb->stream));
    GCU(cudaStreamSynchronize(gb->stream));
    return VRT_OK;
}

VRT_API int vrt_gbuffer_render_present(VrtGBuffer* gb, VrtContext* ctx, const VrtFrame* frame, uint32_t* out_rgba8) {
    if (!gb) return VRT_ERR_INVALID;
    if (!ctx || !frame || !out_rgba8) return gfail(gb, VRT_ERR_INVALID, "vrt_gbuffer_render_present: NULL argument");
    if (!gb->camera_set_for_frame) return gfail(gb, VRT_ERR_STATE, "vrt_gbuffer_render_present: call vrt_gbuffer_set_camera first (CpuRenderer.cpp:423)");
    if (frame->width != gb->w || frame->height != gb->h) return gfail(gb, VRT_ERR_INVALID, "vrt_gbuffer_render_present: frame size differs from the camera's view size");
    if (frame->flags & VRT_FRAME_LINEAR_OUTPUT) return gfail(gb, VRT_ERR_INVALID, "vrt_gbuffer_render_present: the blit reads the tile layout (CopyTiledFramebuffer.comp)");
    GCU(cudaSetDevice(gb->device));
    const size_t n = (size_t)gb->w * gb->h;
    if (!gb->d_tiles) GCU(cudaMalloc((void**)&gb->d_tiles, n * 16));
    if (!gb->d_rgba) GCU(cudaMalloc((void**)&gb->d_rgba, n * 4));
    int st = vrt_render_device(ctx, frame, gb->d_tiles, nullptr, gb->stream);  // the traversal path, result stays in HBM
    if (st != VRT_OK) return gfail(gb, st, std::string("vrt_render_device: ") + vrt_last_error(ctx));
    st = vrt_gbuffer_denoise_present_device(gb, gb->d_tiles, gb->d_rgba, gb->stream);
    if (st != VRT_OK) return st;
    GCU(cudaMemcpyAsync(out_rgba8, gb->d_rgba, n * 4, cudaMemcpyDeviceToHost, gb->stream));  // 4 B/px leave the GPU, not 16
    GCU(cudaStreamSynchronize(gb->stream));
    return VRT_OK;
}

VRT_API int vrt_gbuffer_read(VrtGBuffer* gb, uint32_t plane, void* out) {
    if (!gb) return VRT_ERR_INVALID;
    if (!out || !gb->irr) return gfail(gb, VRT_ERR_STATE, "vrt_gbuffer_read: nothing allocated yet");
    GCU(cudaSetDevice(gb->device));
    const size_t n = (size_t)gb->w * gb->h;
    const void* src = nullptr;
    size_t bytes = 0;
    switch (plane) {
    case VRT_PLANE_IRRADIANCE: src = gb->irr, bytes = n * 16; break;
    case VRT_PLANE_PREV_IRRADIANCE: src = gb->prev_irr, bytes = n * 16; break;
    case VRT_PLANE_TEMP_IRRADIANCE: src = gb->temp_irr, bytes = n * 16; break;
    case VRT_PLANE_MOMENTS: src = gb->moments, bytes = n * 4; break;
    case VRT_PLANE_HISTORY_LEN: src = gb->hist, bytes = n; break;
    default: return gfail(gb, VRT_ERR_INVALID, "vrt_gbuffer_read: unknown plane");
    }
    GCU(cudaDeviceSynchronize());
    GCU(cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
    return VRT_OK;
}

VRT_API int vrt_gbuffer_last_launches(const VrtGBuffer* gb, uint64_t* out) {
    if (!gb || !out) return VRT_ERR_INVALID;
    *out = gb->last_launches;
    return VRT_OK;
}

}  // extern "C"
#endif  // !VRT_HOST_EMULATION
