// VRCP14PS / VRSQRT14PS, bit-exact.
//
// The reference's CPU renderer normalises every ray direction with simd::approx_rsqrt = _mm512_rsqrt14_ps, takes
// approx_sqrt = rsqrt14(x) * x in SampleDirection and approx_rcp = _mm512_rcp14_ps in the cube-map projection
// (src/LibGlimpsw/SwRast/SIMD_AVX512.h:136-138, used by SIMD.h:113, CpuRenderer.cpp:288, Texture.h:285).  Both instructions
// are architecturally defined (relative error < 2^-14, same bits on every AVX-512 implementation), so "the reference's
// frame" depends on their exact values.  Measured on the instruction itself (tools/x86_approx14/): the result depends on
// the top 16 (rcp14) / 15 (rsqrt14, plus the exponent's parity) mantissa bits only, and equals a 64-segment piecewise-linear
// function in integer arithmetic —  bits = ((c0[seg] - c1[seg] * t) >> 9) << 7,  t = the 10 mantissa bits below the segment
// index — with exact results for exact powers of two (four).  The coefficients below are the UNIQUE solution of that model
// (tools/x86_approx14/fit_tables.py); tools/x86_approx14/verify_exhaustive.c and tests/test_x86_approx14.py compare the
// functions with the instruction on all 2^32 inputs (zero mismatches, incl. zeros, denormals, infinities, NaNs, negative
// arguments, denormal results).
#pragma once
#include <stdint.h>
#ifndef VRT_X86_APPROX_QUAL
#if defined(__CUDACC__)
#define VRT_X86_APPROX_QUAL __host__ __device__ __forceinline__
#else
#define VRT_X86_APPROX_QUAL static inline
#endif
#endif

namespace vrt_x86 {

struct Coef {
    uint32_t c0, c1;
};
#if defined(__CUDA_ARCH__)
#define VRT_X86_TABLE __device__ const
#else
#define VRT_X86_TABLE static const
#endif
VRT_X86_TABLE Coef kRcp14[64] = {
    {0xFDFFF900u, 1009u}, {0xFDF03600u, 977u}, {0xFDE0F200u, 949u}, {0xFDD22000u, 921u},
    {0xFDC3BB00u, 893u}, {0xFDB5C700u, 869u}, {0xFDA83300u, 843u}, {0xFD9B0600u, 821u},
    {0xFD8E3200u, 797u}, {0xFD81BC00u, 777u}, {0xFD759800u, 755u}, {0xFD69CA00u, 735u},
    {0xFD5E4C00u, 717u}, {0xFD531B00u, 699u}, {0xFD483100u, 681u}, {0xFD3D8C00u, 663u},
    {0xFD332F00u, 647u}, {0xFD291100u, 631u}, {0xFD1F3600u, 617u}, {0xFD159300u, 601u},
    {0xFD0C2D00u, 587u}, {0xFD02FF00u, 573u}, {0xFCFA0A00u, 561u}, {0xFCF14500u, 547u},
    {0xFCE8B600u, 535u}, {0xFCE05800u, 523u}, {0xFCD82D00u, 513u}, {0xFCD02A00u, 501u},
    {0xFCC85700u, 491u}, {0xFCC0AD00u, 479u}, {0xFCB92E00u, 469u}, {0xFCB1D700u, 459u},
    {0xFCAAAA00u, 451u}, {0xFCA39F00u, 441u}, {0xFC9CBC00u, 433u}, {0xFC95F800u, 423u},
    {0xFC8F5A00u, 415u}, {0xFC88DD00u, 407u}, {0xFC828000u, 399u}, {0xFC7C4300u, 391u},
    {0xFC762800u, 385u}, {0xFC702500u, 377u}, {0xFC6A4100u, 369u}, {0xFC647B00u, 363u},
    {0xFC5ED100u, 357u}, {0xFC593D00u, 349u}, {0xFC53C600u, 343u}, {0xFC4E6800u, 337u},
    {0xFC492300u, 331u}, {0xFC43F500u, 325u}, {0xFC3EDE00u, 319u}, {0xFC39E200u, 315u},
    {0xFC34F600u, 309u}, {0xFC302100u, 303u}, {0xFC2B6400u, 299u}, {0xFC26B700u, 293u},
    {0xFC222200u, 289u}, {0xFC1D9F00u, 285u}, {0xFC192D00u, 279u}, {0xFC14D300u, 275u},
    {0xFC108900u, 271u}, {0xFC0C4F00u, 267u}, {0xFC082500u, 263u}, {0xFC040B00u, 259u},
};
// [0, 32): even exponent (argument mantissa in [1, 2)); [32, 64): odd exponent (argument taken as [2, 4))
VRT_X86_TABLE Coef kRsqrt14[64] = {
#include "x86_approx14_rsqrt.inc"
};

VRT_X86_APPROX_QUAL uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } v;
    v.f = f;
    return v.u;
#endif
}
VRT_X86_APPROX_QUAL float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } v;
    v.u = u;
    return v.f;
#endif
}
VRT_X86_APPROX_QUAL Coef coef(const Coef* table, uint32_t i) {
#if defined(__CUDA_ARCH__)
    const uint2 c = __ldg(reinterpret_cast<const uint2*>(table) + i);
    return Coef{c.x, c.y};
#else
    return table[i];
#endif
}
VRT_X86_APPROX_QUAL int clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return __builtin_clz(v);
#endif
}

// _mm512_rsqrt14_ps for a POSITIVE NORMAL argument whose result is normal too (any x in [2^-126, 2^127]: the result lies in
// [2^-64, 2^63]).  12 integer instructions and one 8-byte table load.
VRT_X86_APPROX_QUAL float rsqrt14_pos_normal(float x) {
    const uint32_t u = f2u(x);
    const uint32_t idx = ((u >> 18) & 63u) ^ 32u;  // (exponent - 127 odd) << 5 | top 5 mantissa bits
    const uint32_t t = (u >> 8) & 0x3FFu;
    const Coef c = coef(kRsqrt14, idx);
    uint32_t rb = ((c.c0 - c.c1 * t) >> 9) << 7;
    if ((u & 0x00FFFFFFu) == 0x00800000u) rb = 0x3F800000u;  // exact power of four
    // x = (1.m * 2^odd) * 4^k  ->  result = table value * 2^-k,  k = (e - 127 - odd) / 2
    const int k = ((int)(u >> 23) - 127 - (int)(idx >> 5)) >> 1;
    return u2f(rb - ((uint32_t)k << 23));
}

#if defined(__CUDACC__)
// The same rsqrt14 coefficients in CONSTANT memory, for callers whose lanes mostly share the segment (camera rays: the squared length
// of neighbouring pixels' directions differs in low mantissa bits): an indexed LDC instead of a global load in front of the traversal.
__constant__ Coef kRsqrt14Const[64] = {
#define VRT_X86_COEF_LIST_RSQRT
#include "x86_approx14_rsqrt.inc"
#undef VRT_X86_COEF_LIST_RSQRT
};
#endif
VRT_X86_APPROX_QUAL float rsqrt14_pos_normal_uniform(float x) {
    const uint32_t u = f2u(x);
    const uint32_t idx = ((u >> 18) & 63u) ^ 32u;
    const uint32_t t = (u >> 8) & 0x3FFu;
#if defined(__CUDA_ARCH__)
    const Coef c = kRsqrt14Const[idx];
#else
    const Coef c = kRsqrt14[idx];
#endif
    uint32_t rb = ((c.c0 - c.c1 * t) >> 9) << 7;
    if ((u & 0x00FFFFFFu) == 0x00800000u) rb = 0x3F800000u;
    const int k = ((int)(u >> 23) - 127 - (int)(idx >> 5)) >> 1;
    return u2f(rb - ((uint32_t)k << 23));
}

// _mm512_rsqrt14_ps, every argument.
VRT_X86_APPROX_QUAL float rsqrt14(float x) {
    const uint32_t u = f2u(x), sign = u & 0x80000000u, a = u & 0x7FFFFFFFu;
    if (a > 0x7F800000u) return u2f(u | 0x00400000u);  // NaN -> quiet NaN, payload kept
    if (a == 0u) return u2f(sign | 0x7F800000u);       // +-0 -> +-inf
    if (sign) return u2f(0xFFC00000u);                 // negative -> real indefinite
    if (a == 0x7F800000u) return 0.0f;                 // +inf -> +0
    int e = (int)(a >> 23);
    uint32_t m = a & 0x7FFFFFu;
    if (e == 0) {  // denormal argument (MXCSR.DAZ = 0): normalise
        const int sh = clz32(m) - 8;
        m = (m << sh) & 0x7FFFFFu;
        e = 1 - sh;
    }
    const int E = e - 127, odd = E & 1, k = (E - odd) / 2;
    uint32_t rb = 0x3F800000u;
    if (m != 0u || odd) {
        const Coef c = coef(kRsqrt14, ((uint32_t)odd << 5) | (m >> 18));
        rb = ((c.c0 - c.c1 * ((m >> 8) & 0x3FFu)) >> 9) << 7;
    }
    return u2f(rb - ((uint32_t)k << 23));
}

// _mm512_rcp14_ps, every argument (denormal arguments and denormal results included: MXCSR.DAZ = FTZ = 0).
VRT_X86_APPROX_QUAL float rcp14(float x) {
    const uint32_t u = f2u(x), sign = u & 0x80000000u, a = u & 0x7FFFFFFFu;
    if (a > 0x7F800000u) return u2f(u | 0x00400000u);
    if (a == 0x7F800000u) return u2f(sign);
    if (a == 0u) return u2f(sign | 0x7F800000u);
    int e = (int)(a >> 23);
    uint32_t m = a & 0x7FFFFFu;
    if (e == 0) {
        const int sh = clz32(m) - 8;
        m = (m << sh) & 0x7FFFFFu;
        e = 1 - sh;
    }
    uint32_t rb = 0x3F800000u;
    if (m != 0u) {
        const Coef c = coef(kRcp14, m >> 17);
        rb = ((c.c0 - c.c1 * ((m >> 7) & 0x3FFu)) >> 9) << 7;
    }
    const int re = (int)(rb >> 23) - (e - 127);
    const uint32_t rm = rb & 0x7FFFFFu;
    if (re >= 255) return u2f(sign | 0x7F800000u);
    if (re <= 0) {  // denormal result: truncated
        const int sh = 1 - re;
        return u2f(sh > 24 ? sign : (sign | ((rm | 0x800000u) >> sh)));
    }
    return u2f(sign | ((uint32_t)re << 23) | rm);
}

}  // namespace vrt_x86
