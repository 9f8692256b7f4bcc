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
#include <ucontext.h>

#include <cstdlib>
#include <vector>

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
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct dim3 { unsigned x = 1, y = 1, z = 1; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return {x, y}; }
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
// ---- warps ------------------------------------------------------------------------------------------------------------------------
// Three ways to run device code that talks to the other lanes of its warp:
//  (1) one lane at a time, a warp of one: votes see only the caller (per-lane functions: the traversal's results are defined per lane);
//  (2) LOCKSTEP (EmuWarp, run_warp_lockstep): the 32 lanes of a warp are 32 host threads and every collective is a rendezvous — what the
//      frame kernels need now that a half-warp is one SIMD packet of the reference (packet_votes), and what the persistent trace pass needs
//      (its loop control is a vote per trip).  A lane that returns leaves the rendezvous (it votes 0 from then on, like an exited thread);
//  (3) shuffle REPLAY (ShflReplay below) for the straight-line warp-per-brick kernels.
struct EmuWarp {
    ucontext_t sched;      // the scheduler (run_warp_lockstep's own context)
    ucontext_t ctx[32];    // one fiber per lane, all on the calling OS thread
    bool done[32];
    unsigned xchg[32];     // words deposited at the current rendezvous (0 for lanes that have returned)
    unsigned result[32];   // snapshot handed to the lanes when they resume
    unsigned block_idx, block_dim, tid0;
    void (*body)(void*, int);
    void* arg;
};
static thread_local EmuWarp* t_emu_warp = nullptr;
static thread_local int t_emu_lane = 0;
static inline void emu_enter_lane(EmuWarp* w, int l) {  // (all fibers share the OS thread's thread_locals: set them on every switch)
    blockIdx.x = w->block_idx, blockIdx.y = blockIdx.z = 0;
    blockDim.x = w->block_dim, blockDim.y = blockDim.z = 1;
    threadIdx.x = w->tid0 + (unsigned)l, threadIdx.y = threadIdx.z = 0;
    t_emu_lane = l;
}
static inline unsigned emu_exchange(unsigned mine, unsigned out[32]) {  // every live lane deposits a word, then all read all of them
    EmuWarp* w = t_emu_warp;
    const int l = t_emu_lane;
    w->xchg[l] = mine;
    swapcontext(&w->ctx[l], &w->sched);  // back to the scheduler until every live lane has arrived
    for (int k = 0; k < 32; k++) out[k] = w->result[k];
    return 0;
}
static void emu_lane_entry(unsigned lo, unsigned hi) {
    EmuWarp* w = reinterpret_cast<EmuWarp*>(((unsigned long long)hi << 32) | lo);
    const int l = t_emu_lane;
    w->body(w->arg, l);
    w->done[l] = true;
    w->xchg[l] = 0u;  // an exited lane votes 0 / false from now on
    swapcontext(&w->ctx[l], &w->sched);
}
// runs fn(lane) for the 32 lanes of one warp in LOCKSTEP: 32 fibers on the calling thread; a collective hands control back to the
// scheduler, which resumes the lanes in turn until all of them wait at the same rendezvous (or have returned), then publishes the words
template <class F>
static void run_warp_lockstep(unsigned block_idx, unsigned block_dim, unsigned tid0, F fn) {
    static thread_local char* stacks = nullptr;
    const size_t kStack = 256u << 10;
    if (!stacks) stacks = static_cast<char*>(std::malloc(32 * kStack));
    EmuWarp w;
    w.block_idx = block_idx, w.block_dim = block_dim, w.tid0 = tid0;
    w.body = [](void* a, int lane) { (*static_cast<F*>(a))(lane); };
    w.arg = &fn;
    const unsigned long long wp = reinterpret_cast<unsigned long long>(&w);
    for (int l = 0; l < 32; l++) {
        w.done[l] = false;
        w.xchg[l] = w.result[l] = 0u;
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = stacks + (size_t)l * kStack;
        w.ctx[l].uc_stack.ss_size = kStack;
        w.ctx[l].uc_link = nullptr;
        makecontext(&w.ctx[l], reinterpret_cast<void (*)()>(emu_lane_entry), 2, (unsigned)(wp & 0xFFFFFFFFu), (unsigned)(wp >> 32));
    }
    EmuWarp* outer = t_emu_warp;
    t_emu_warp = &w;
    for (;;) {
        bool any = false;
        for (int l = 0; l < 32; l++)
            if (!w.done[l]) {
                any = true;
                emu_enter_lane(&w, l);
                swapcontext(&w.sched, &w.ctx[l]);  // until the lane reaches a collective or returns
            }
        if (!any) break;
        for (int l = 0; l < 32; l++) w.result[l] = w.xchg[l];
    }
    t_emu_warp = outer;
}
static inline unsigned __activemask() { return t_emu_warp ? 0xFFFFFFFFu : 1u; }
template <class... A> static inline void __syncwarp(A...) {}
static inline void __syncthreads() {}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
    if (!t_emu_warp) return v;
    static_assert(sizeof(T) == 4, "lockstep shuffles handle 32-bit values");
    unsigned bits, all[32];
    std::memcpy(&bits, &v, 4);
    emu_exchange(bits, all);
    T out;
    std::memcpy(&out, &all[src & 31], 4);
    return out;
}
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
static inline unsigned __ballot_sync(unsigned mask, int p) {
    if (t_emu_warp) {
        unsigned all[32], word = 0;
        emu_exchange(p ? 1u : 0u, all);
        for (int l = 0; l < 32; l++) word |= (all[l] & 1u) << l;
        return word & mask;
    }
    ShflReplay* r = g_shfl_replay;
    if (!r) return p ? 1u : 0u;  // one-lane warp
    const int k = r->call++;
    r->cur[k][r->lane] = p ? 1u : 0u;
    unsigned word = 0;
    for (int l = 0; l < 32; l++) word |= (r->prev[k][l] & 1u) << l;
    return word;
}
static inline int __any_sync(unsigned mask, int p) { return t_emu_warp ? (__ballot_sync(mask, p) != 0u) : p; }
// (only ever used with the mask of the lanes that entered a branch together; the emulated cast_ray takes that decision per lane)
static inline int __all_sync(unsigned, int p) { return p; }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned) { return v; }
template <class T, class U> static inline T atomicAdd(T* p, U v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_RELAXED); }

using std::isinf;
using std::max;
using std::min;
