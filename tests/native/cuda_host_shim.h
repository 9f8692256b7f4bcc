// The few CUDA device-side names voxelrt_b200/csrc/vrt_post.cu and vrt_glsl.cuh use, defined for a HOST build (tests/native/emu_*.cpp):
// those kernels are plain per-thread functions (no shared memory, no barriers), so running every (block, thread) of
// the grid in a loop executes exactly the arithmetic the GPU executes.  Build with -ffp-contract=off: __fmul_rn(a,b) + c must not
// fuse, FMAs happen only where the source says __fmaf_rn.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <immintrin.h>

#include <algorithm>
#include <cfenv>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__
#define __align__(n) alignas(n)

struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct dim3 { unsigned x = 1, y = 1, z = 1; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
static inline float3 make_float3(float x, float y, float z) { return {x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }

static thread_local dim3 blockIdx, threadIdx, blockDim, gridDim;

template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
struct __half { unsigned short v; };
static inline __half __ushort_as_half(unsigned short u) { return {u}; }
static inline unsigned short __half_as_ushort(__half h) { return h.v; }
static inline float __half2float(__half h) { return _cvtsh_ss(h.v); }                                              // vcvtph2ps: exact
static inline __half __float2half_rn(float f) { return {(unsigned short)_cvtss_sh(f, _MM_FROUND_TO_NEAREST_INT)}; }  // vcvtps2ph, RNE
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __float2int_rd(float x) { return (int)std::floor(x); }
static inline int __float2int_rn(float x) { return (int)std::lrintf(x); }
static inline float __int2float_rn(int x) { return (float)x; }
// directed rounding (FADD.RM & co.): the x87/SSE rounding mode is switched around ONE volatile operation
static inline float __fadd_rd(float a, float b) { std::fesetround(FE_DOWNWARD); volatile float x = a, y = b; volatile float r = x + y; std::fesetround(FE_TONEAREST); return r; }
static inline float __fsub_rd(float a, float b) { std::fesetround(FE_DOWNWARD); volatile float x = a, y = b; volatile float r = x - y; std::fesetround(FE_TONEAREST); return r; }
static inline float __fmaf_rd(float a, float b, float c) { std::fesetround(FE_DOWNWARD); volatile float x = a, y = b, z = c; volatile float r = __builtin_fmaf(x, y, z); std::fesetround(FE_TONEAREST); return r; }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const unsigned long long v = (unsigned long long)x | ((unsigned long long)y << 32);
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xF;
        unsigned b = (unsigned)(v >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) b = (b & 0x80) ? 0xFF : 0x00;
        r |= b << (8 * i);
    }
    return r;
}
// one lane at a time: a warp of one (the traversal's results are defined per lane, whatever the warp it runs in)
static inline unsigned __activemask() { return 1u; }
static inline int __all_sync(unsigned, int p) { return p; }
static inline int __any_sync(unsigned, int p) { return p; }
template <class... A> static inline void __syncwarp(A...) {}
static inline void __syncthreads() {}
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
// Warp REPLAY for kernels whose lanes exchange data through a fixed sequence of xor-shuffles and have no divergent control flow around
// them (the warp-per-brick occupancy build): the emulator runs all 32 lanes once per shuffle call; call #k of round r returns what lane
// (l ^ mask) passed to call #k in the PREVIOUS round — correct for every k < r by induction, a placeholder otherwise — so after
// (number of calls + 1) rounds every lane has computed with the right values (stores of earlier rounds are simply overwritten).
struct ShflReplay {
    unsigned prev[64][32], cur[64][32];
    int call;   // index of the next shuffle call of the running lane
    int lane;
};
static thread_local ShflReplay* g_shfl_replay = nullptr;
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int mask) {
    ShflReplay* r = g_shfl_replay;
    if (!r) return v;  // one-lane warp
    static_assert(sizeof(T) == 4, "replay handles 32-bit values");
    unsigned bits;
    std::memcpy(&bits, &v, 4);
    const int k = r->call++;
    r->cur[k][r->lane] = bits;
    const unsigned other = r->prev[k][r->lane ^ mask];
    T out;
    std::memcpy(&out, &other, 4);
    return out;
}
static inline unsigned __ballot_sync(unsigned, int p) {
    ShflReplay* r = g_shfl_replay;
    if (!r) return p ? 1u : 0u;  // one-lane warp
    const int k = r->call++;
    r->cur[k][r->lane] = p ? 1u : 0u;
    unsigned word = 0;
    for (int l = 0; l < 32; l++) word |= (r->prev[k][l] & 1u) << l;
    return word;
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned) { return v; }
template <class T, class U> static inline T atomicAdd(T* p, U v) { T o = *p; *p = (T)(o + v); return o; }
using std::isinf;
using std::max;
using std::min;
