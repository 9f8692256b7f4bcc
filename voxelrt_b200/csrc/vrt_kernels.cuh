// Kernels of the traversal path (sm_100a).  See DESIGN.md §5 for the launch shapes and the
// roofline that bounds each one.
#pragma once
#include "vrt_shade.cuh"

namespace vrt {

// ---------------------------------------------------------------------------------------------
// K_trace: explicit rays, one thread per ray.  Replaces the file-static RayCast
// (CpuRenderer.cpp:172-224) for callers that bring their own rays.
// ---------------------------------------------------------------------------------------------
template <bool METRICS>
__global__ void __launch_bounds__(128, 8) k_trace(const __grid_constant__ DevScene S, const __grid_constant__ RayFrame W,
                                               const float* __restrict__ origin3, const float* __restrict__ dir3, uint32_t max_iters,
                                               uint64_t n, VrtHit* __restrict__ out, DevMetrics* metrics) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = i < n;
    HitLane H;
    CastResult R;
    R.iters = R.n_sector = R.n_cell = 0;
    R.capped = false;
    H.hit = false;
    if (valid) {
        float ox = __ldg(origin3 + 3 * i), oy = __ldg(origin3 + 3 * i + 1), oz = __ldg(origin3 + 3 * i + 2);
        float dx = __ldg(dir3 + 3 * i), dy = __ldg(dir3 + 3 * i + 1), dz = __ldg(dir3 + 3 * i + 2);
        cast_ray<METRICS>(S, W, ox, oy, oz, dx, dy, dz, max_iters, H, R);
        store_hit(out + i, H, R);
    }
    if (METRICS) {
        __syncwarp();
        metrics_add(metrics, R, valid, valid && H.hit);
    }
}

// ---------------------------------------------------------------------------------------------
// K_render (v1): one warp per 8x4-pixel tile = two of the reference's 4x4 SIMD tiles side by
// side, so lanes 0-15 / 16-31 fill one Framebuffer::Tile each (CpuRenderer.cpp:299-309) with
// 64-byte coalesced stores.  Warp tiles are numbered inside 32x32-pixel macro tiles (the unit
// of the multi-GPU screen split): macro tile t belongs to rank t % part_count.
// ---------------------------------------------------------------------------------------------
// (ROWS is a template parameter: a run-time test of F.flags here costs the default kernel 4 % — with the extra branch in the
// prologue ptxas no longer keeps the global-memory descriptor in a uniform register across the traversal loop)
template <bool ROWS>
__device__ __forceinline__ bool warp_tile_origin(const FrameParams& F, uint32_t work, uint32_t& x0, uint32_t& y0) {
    if (ROWS) {
        // band split: the unit is a VRT_BAND_ROWS (8) pixel high band; inside a band warp tiles are numbered in 32x8-pixel
        // groups (4 across, 2 down) so the 4 warps of a CTA still cover one 32x4 strip
        uint32_t group = work >> 3, sub = work & 7u;
        uint32_t lb = F.macros_x_magic ? __umulhi(group, F.macros_x_magic) : group, gx = group - lb * F.macros_x;
        uint32_t band = lb * F.part_count + F.part_index;
        x0 = (gx << 5) + ((sub & 3u) << 3);
        y0 = band * VRT_BAND_ROWS + ((sub >> 2) << 2);
        return x0 < F.width && y0 < F.height;
    }
    uint32_t macro_local = work >> 5, sub = work & 31u;
    uint32_t macro = macro_local * F.part_count + F.part_index;
    // macro / macros_x by multiplication with ceil(2^32 / macros_x): exact while macro * macros_x < 2^32
    uint32_t my = F.macros_x_magic ? __umulhi(macro, F.macros_x_magic) : macro, mx = macro - my * F.macros_x;
    x0 = (mx << 5) + ((sub & 3u) << 3);
    y0 = (my << 5) + ((sub >> 2) << 2);
    return x0 < F.width && y0 < F.height;
}

__device__ __forceinline__ void store_pixel(const FrameParams& F, uint32_t x, uint32_t y, const PixelOut& P) {
    if (F.flags & VRT_FRAME_COMPACT) {  // primary-only frames: albedo + depth, the constant irradiance is not stored (include/voxelrt_b200.h)
        if (F.flags & VRT_FRAME_LINEAR_OUTPUT) {
            uint32_t* o = reinterpret_cast<uint32_t*>(F.out);
            const size_t n = (size_t)F.width * F.height, p = (size_t)y * F.width + x;
            o[p] = P.albedo;
            o[n + p] = __float_as_uint(P.depth);
        } else {
            VrtTileAD* t = reinterpret_cast<VrtTileAD*>(F.out) + ((size_t)(y >> 2) * (F.width >> 2) + (x >> 2));
            const uint32_t lane = (x & 3u) | ((y & 3u) << 2);
            t->albedo[lane] = P.albedo;
            t->depth[lane] = P.depth;
        }
    } else if (F.flags & VRT_FRAME_LINEAR_OUTPUT) {
        uint32_t* o = reinterpret_cast<uint32_t*>(F.out);
        size_t n = (size_t)F.width * F.height, p = (size_t)y * F.width + x;
        o[p] = P.albedo;
        o[n + p] = __float_as_uint(P.depth);
        o[2 * n + p] = P.irr_rg;
        o[3 * n + p] = P.irr_bx;
    } else {
        VrtTile* t = reinterpret_cast<VrtTile*>(F.out) + ((size_t)(y >> 2) * (F.width >> 2) + (x >> 2));
        uint32_t lane = (x & 3u) | ((y & 3u) << 2);
        t->albedo[lane] = P.albedo;
        t->depth[lane] = P.depth;
        t->irr_rg[lane] = P.irr_rg;
        t->irr_bx[lane] = P.irr_bx;
    }
}

#ifndef VRT_RENDER_THREADS
#define VRT_RENDER_THREADS 128  // 4 warp tiles per CTA
#endif
// resident warps per SM: the primary-only kernel fits 56 registers (36 warps); the bounce kernel would take 71, runs best capped
// at 48 (40 warps, a few spills outside the loop: +3 % over 32 warps at 64 registers; 44/48 warps at 40 registers lose it again)
#ifndef VRT_RENDER_WARPS_PRIMARY
#define VRT_RENDER_WARPS_PRIMARY 36
#endif
#ifndef VRT_RENDER_WARPS_BOUNCE
#define VRT_RENDER_WARPS_BOUNCE 40
#endif
#define VRT_RENDER_CTAS(PRIMARY) (((PRIMARY) ? VRT_RENDER_WARPS_PRIMARY : VRT_RENDER_WARPS_BOUNCE) * 32 / VRT_RENDER_THREADS)
template <bool METRICS, bool PRIMARY, bool ROWS = false, bool OCC = false>
__device__ __forceinline__ void render_warp_tile(const DevScene& S, const FrameParams& F, uint32_t work) {
    uint32_t x0, y0;
    if (!warp_tile_origin<ROWS>(F, work, x0, y0)) return;  // warp-uniform
    uint32_t lane = threadIdx.x & 31u;
    uint32_t x = x0 + ((lane >> 4) << 2) + (lane & 3u);
    uint32_t y = y0 + ((lane >> 2) & 3u);
    bool valid = x < F.width && y < F.height;
    PixelOut P;
    if (PRIMARY) shade_pixel_primary<METRICS>(S, F, x, y, valid, P);
    else shade_pixel<METRICS, OCC>(S, F, x, y, valid, P);
    if (valid) store_pixel(F, x, y, P);
}

template <bool METRICS, bool PRIMARY, bool ROWS = false, bool OCC = false>
__global__ void __launch_bounds__(VRT_RENDER_THREADS, VRT_RENDER_CTAS(PRIMARY && !METRICS)) k_render(const __grid_constant__ DevScene S, const __grid_constant__ FrameParams F) {
    const uint32_t g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= F.n_work - F.work_offset) return;  // warp-uniform
    render_warp_tile<METRICS, PRIMARY, ROWS, OCC>(S, F, F.work_add + (uint32_t)(F.work_mul * (int32_t)g));
}

// ---------------------------------------------------------------------------------------------
// Wavefront form of a frame with bounces (vrt_shade.cuh): camera pass, then per bounce level a TRACE pass and a SHADE pass.
// ---------------------------------------------------------------------------------------------
template <bool ROWS>
__global__ void __launch_bounds__(VRT_RENDER_THREADS, VRT_RENDER_CTAS(false)) k_wave_primary(const __grid_constant__ DevScene S, const __grid_constant__ FrameParams F,
                                                                                             const __grid_constant__ WaveBuffers B) {
    const uint32_t g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= F.n_work - F.work_offset) return;  // warp-uniform
    const uint32_t work = F.work_add + (uint32_t)(F.work_mul * (int32_t)g);  // (grid order: see FrameParams::work_add)
    uint32_t x0 = 0, y0 = 0;
    const bool in_frame = warp_tile_origin<ROWS>(F, work, x0, y0);  // warp-uniform; the warp stays for the votes
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t x = x0 + ((lane >> 4) << 2) + (lane & 3u);
    const uint32_t y = y0 + ((lane >> 2) & 3u);
    wave_primary_pixel(S, F, B, (work - F.work_offset) * 32u + lane, x, y, in_frame && x < F.width && y < F.height);
}
template <bool ROWS>
__global__ void __launch_bounds__(VRT_RENDER_THREADS) k_wave_shade(const __grid_constant__ DevScene S, const __grid_constant__ FrameParams F,
                                                                   const __grid_constant__ WaveBuffers B, uint32_t level) {
    const uint32_t work = F.work_offset + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (work >= F.n_work) return;
    uint32_t x0 = 0, y0 = 0;
    const bool in_frame = warp_tile_origin<ROWS>(F, work, x0, y0);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t x = x0 + ((lane >> 4) << 2) + (lane & 3u);
    const uint32_t y = y0 + ((lane >> 2) & 3u);
    wave_shade_pixel(S, F, B, (work - F.work_offset) * 32u + lane, x, y, in_frame && x < F.width && y < F.height, level);
}

// The trace pass: persistent warps, one ray per lane, lanes refilled from the level's queue whenever fewer than `refill` of the warp's
// 32 rays are still in flight (and once more when none is).  The queue holds FAST rays and rays whose direction is NaN in all three
// components (quirk Q7: 0.8 % of all bounce rays; they take ONE lean trip — the reference's second trip lands at INT_MIN, outside the view,
// whatever the first one did); rays outside the fast loop's domain (an exactly zero component: 0.4 % of the blue-noise samples; far or
// outside origins) were queued apart by their producer and are traced by k_wave_trace_generic, 32 to a warp — one such ray traced inside
// this kernel would drag its warp through the generic loop alone (measured: 10 % of the kernel's instructions at 1 thread per instruction).
#ifndef VRT_TRACE_REFILL
#define VRT_TRACE_REFILL 24
#endif
#ifndef VRT_TRACE_CTAS
#define VRT_TRACE_CTAS 10  // measured (terrain / Sponza / 10 GB terrain): 10 CTAs = 40 warps per SM at 47 registers beats 8 (52 registers) by 2-5 %, 12 adds nothing
#endif
struct TraceArgs {
    const RayRec* rays;
    const uint32_t* n;  // rays queued for this level (from the front of `rays`)
    uint32_t* head;     // refill cursor (zeroed per frame)
    HitRec* hits;
    uint32_t max_iters;
    uint32_t refill;    // refill when fewer lanes than this are in flight
    const uint32_t* n_generic;  // k_wave_trace_generic: rays queued from the end of `rays`
    uint32_t capacity;
};
// CTAS = resident CTAs per SM the kernel is compiled for (8: <= 64 registers, 10: <= 48, 12: <= 40): the loop is latency-bound (two dependent
// loads per trip behind ~100 dependent ALU instructions), so warps in flight matter more than a few registers
template <int CTAS>
__global__ void __launch_bounds__(VRT_RENDER_THREADS, CTAS) k_wave_trace(const __grid_constant__ DevScene S, const __grid_constant__ RayFrame W,
                                                                         const __grid_constant__ TraceArgs A) {
    const uint32_t n = *A.n;
    const unsigned lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    // loop constants pinned in registers (ptxas otherwise re-loads kernel parameters from the constant bank every trip: see cast_loop_fast)
    const int opaque0 = (int)blockDim.z - 1;
    const float opaque0f = __int_as_float(opaque0);
    LeanFrame C;
    C.mgx = __fadd_rn(W.mgx, opaque0f), C.mgy = __fadd_rn(W.mgy, opaque0f), C.mgz = __fadd_rn(W.mgz, opaque0f), C.r32 = __fadd_rn(0.03125f, opaque0f);
    C.strz = (int)S.sxp + opaque0, C.stry = (int)S.sxzp + opaque0, C.hoff = W.hoff + opaque0;
    {
        unsigned long long hp = (unsigned long long)S.hdr, cp = (unsigned long long)S.cells;
#ifndef VRT_HOST_EMULATION
        asm volatile("add.u64 %0, %0, %1;" : "+l"(hp) : "l"((unsigned long long)(unsigned)opaque0));
        asm volatile("add.u64 %0, %0, %1;" : "+l"(cp) : "l"((unsigned long long)(unsigned)opaque0));
#endif
        C.hdrp = reinterpret_cast<const uint4*>(hp);
        C.cellp = reinterpret_cast<const char*>(cp);
    }
    VRT_PIN_F(C.mgx);
    VRT_PIN_F(C.mgy);
    VRT_PIN_F(C.mgz);
    VRT_PIN_F(C.r32);
    VRT_PIN_R(C.strz);
    VRT_PIN_R(C.stry);
    VRT_PIN_R(C.hoff);
    const uint32_t refill = A.refill + (uint32_t)opaque0;
    LeanRay r;
    r.ox = r.oy = r.oz = r.dx = r.dy = r.dz = r.ix = r.iy = r.iz = r.tx = r.ty = r.tz = r.cx = r.cy = r.cz = r.sdx = r.sdy = r.sdz = 0.0f;
    r.nmx = r.nmy = r.nmz = 0;
    uint32_t left = 0, budget = 0, slot = 0;
    unsigned long long hit_at = 0;
    // lane state: 0 idle, 1 a ray is in flight, 2..4 it has ended (2 solid voxel, 3 left the view, 4 out of trips) and its record is not written yet
    uint32_t state = 0;
    uint32_t more = 1;  // the queue may still hold rays (warp-uniform)
    unsigned act = 0u;  // lanes with a ray in flight
    for (;;) {
        // ---- epilogue of the rays that ended, then refill: entered with fewer than `refill` rays in flight (none, once the queue is empty)
        if (state >= 2u) {  // RayCast epilogue (CpuRenderer.cpp:204-223) of the ray that ended in this lane
            if (budget == 1u && state == 4u) state = 3u;  // the second trip of a NaN ray: currPos = NaN -> voxel INT_MIN -> outside
            uint32_t flags = HITREC_CAST | normal_code(r.sdx, r.sdy, r.sdz, r.dx, r.dy, r.dz);
            uint32_t material = 0u;  // (a miss's material is never used at bounce levels >= 1: RenderRow replaces it by the sky, :363-368)
            if (state == 2u) {
                material = ldg_u2(S.palette + __ldg(S.voxels + hit_at)).x;  // :120-132
                flags |= HITREC_HIT;
            } else if (state == 4u) flags |= HITREC_CAPPED;
            if (left == budget && state != 4u) flags |= HITREC_FIRST;  // no completed step
            store_hit_rec(A.hits + slot, r.cx, r.cy, r.cz, material, r.dx, r.dy, r.dz, flags);
            state = 0u;
        }
        if (more) {
            const unsigned want = ~act;
            const uint32_t cnt = (uint32_t)__popc(want);
            const int leader = __ffs((int)want) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(A.head, cnt);
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            more = base + cnt < n ? 1u : 0u;
            const uint32_t idx = base + (uint32_t)__popc(want & lt_mask);
            if (state == 0u && idx < n) {
                const float4* rp = reinterpret_cast<const float4*>(A.rays + idx);
                const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
                r.ox = r0.x, r.oy = r0.y, r.oz = r0.z, r.dx = r1.x, r.dy = r1.y, r.dz = r1.z;
                slot = __float_as_uint(r0.w);
                r.ix = rcp_rn_normal(r.dx), r.iy = rcp_rn_normal(r.dy), r.iz = rcp_rn_normal(r.dz);  // :173 (== IEEE 1/x for fast rays; NaN stays NaN)
                r.tx = __fmul_rn(__fsub_rn(r.dx < 0.0f ? 0.0f : 1.0f, r.ox), r.ix);                  // :175-179
                r.ty = __fmul_rn(__fsub_rn(r.dy < 0.0f ? 0.0f : 1.0f, r.oy), r.iy);
                r.tz = __fmul_rn(__fsub_rn(r.dz < 0.0f ? 0.0f : 1.0f, r.oz), r.iz);
                r.nmx = __float_as_int(r.dx) >> 31, r.nmy = __float_as_int(r.dy) >> 31, r.nmz = __float_as_int(r.dz) >> 31;
                r.cx = r.ox, r.cy = r.oy, r.cz = r.oz;  // :181
                r.sdx = r.sdy = r.sdz = 0.0f;           // :180
                budget = left = __float_as_uint(r1.w) == RAY_NAN ? 1u : A.max_iters;
                state = 1u;
            }
        }
        act = __ballot_sync(0xFFFFFFFFu, state == 1u);
        if (act == 0u) {
            if (!more) break;
            continue;
        }
        // ---- the trip loop: every lane with a ray in flight takes one trip per turn, until too few are left
        const uint32_t need = more ? refill : 1u;
        do {
            if (state == 1u) lean_trip(C, r, state, left, hit_at);
            act = __ballot_sync(0xFFFFFFFFu, state == 1u);
        } while ((uint32_t)__popc(act) >= need);
    }
}

// The rays outside the fast loop's domain (RAY_GENERIC, queued from the end of the ray buffer): cast_loop_generic spells out the x86
// semantics (infinities and NaNs of zero components, cvtps2dq overflow).  One thread per ray; a few thousand rays per level.
__global__ void __launch_bounds__(128) k_wave_trace_generic(const __grid_constant__ DevScene S, const __grid_constant__ RayFrame W,
                                                            const __grid_constant__ TraceArgs A) {
    const uint32_t n = *A.n_generic;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const float4* rp = reinterpret_cast<const float4*>(A.rays + (A.capacity - 1u - k));
        const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
        CastResult R;
        HitLane H;
        cast_loop_generic(S, r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, W.wx, W.wy, W.wz, A.max_iters, R);
        cast_finish<true>(S, R, r1.x, r1.y, r1.z, H);
        const uint32_t flags = HITREC_CAST | H.ncode | (H.hit ? HITREC_HIT : 0u) | (R.capped ? HITREC_CAPPED : 0u) | ((R.iters == 1u && !R.capped) ? HITREC_FIRST : 0u);
        store_hit_rec(A.hits + __float_as_uint(r0.w), H.px, H.py, H.pz, H.material, r1.x, r1.y, r1.z, flags);
    }
}

// K_render, persistent form: the grid is sized to the machine (SMs x resident CTAs) and every WARP pulls warp tiles
// from a ticket counter until the frame is done, so a warp slot never idles waiting for the slowest warp of its CTA
// (tile costs differ by an order of magnitude between sky and horizon tiles).  A warp's first tile is its global warp
// index; every further tile costs one atomicAdd by lane 0.  The counter is never reset: a launch consumes exactly
// (n_work - work_offset) tickets (each warp that ran ends on exactly one failing fetch), so the host advances
// `ticket_base` by that amount per launch; the unsigned difference survives the 32-bit wrap.
template <bool PRIMARY>
__global__ void __launch_bounds__(VRT_RENDER_THREADS, VRT_RENDER_CTAS(PRIMARY)) k_render_persist(const __grid_constant__ DevScene S, const __grid_constant__ FrameParams F,
                                                                                                uint32_t* __restrict__ ticket, uint32_t ticket_base) {
    const uint32_t n_tiles = F.n_work - F.work_offset;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
    uint32_t local = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    while (local < n_tiles) {
        // the ticket of the NEXT tile is requested before this tile is traced, so its latency is hidden
        uint32_t t = 0;
        if ((threadIdx.x & 31u) == 0u) t = atomicAdd(ticket, 1u);
        render_warp_tile<false, PRIMARY>(S, F, F.work_offset + local);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        local = n_warps + (t - ticket_base);
    }
}

// K_palette_albedo: per palette entry, the albedo bits the primary-only frame kernel would compute from the material.
__global__ void k_palette_albedo(const uint2* __restrict__ palette, uint32_t* __restrict__ albedo) {
    albedo[threadIdx.x] = albedo_rgb_bits(palette[threadIdx.x].x);
}

// ---------------------------------------------------------------------------------------------
// K_upload (K2): one warp per dirty brick.  Copies the 512 voxel bytes from the staging buffer
// into the brick's slot and rebuilds its eight 4x4x4 occupancy masks
// (FlatVoxelStorage::UpdateOccupancy, CpuRenderer.cpp:63-83; UpdateOccupancy.comp:8-34).
// Lane l holds voxels [16 l, 16 l + 16): y = l>>2, z in {2(l&3), 2(l&3)+1}, x = 0..7.
// Algorithmic bytes per brick: 512 read + 512 + 64 written (HBM-bound).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_upload_bricks(const uint4* __restrict__ staging, const uint32_t* __restrict__ slots, uint32_t n,
                                                       uint8_t* __restrict__ voxels, uint2* __restrict__ cells) {
    uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp >= n) return;
    uint32_t lane = threadIdx.x & 31u;
    uint32_t slot = __ldg(slots + warp);
    uint4 v = __ldg(staging + (size_t)warp * 32u + lane);
    reinterpret_cast<uint4*>(voxels + (size_t)slot * 512u)[lane] = v;

    // non-zero byte -> 1 bit, 16 voxels -> 16 bits: bits 0-7 row z0 (x=0..7), bits 8-15 row z0+1
    auto nz4 = [](uint32_t w) -> uint32_t {
        uint32_t t = (w | (w >> 4)) & 0x0F0F0F0Fu;  // fold nibbles
        t = (t | (t >> 2)) & 0x03030303u;
        t = (t | (t >> 1)) & 0x01010101u;            // one bit per byte
        return (t | (t >> 7) | (t >> 14) | (t >> 21)) & 0xFu;
    };
    uint32_t r0 = nz4(v.x) | (nz4(v.y) << 4);  // row z0: x 0..7
    uint32_t r1 = nz4(v.z) | (nz4(v.w) << 4);  // row z0+1
    // cell (cx, cz = (l>>1)&1, cy = l>>4); bit = vx + 4 vz + 16 vy, vz0 = 2(l&1), vy = (l>>2)&3
    uint32_t shift = 8u * (lane & 1u) + 16u * ((lane >> 2) & 1u);  // within a 32-bit half
    uint32_t c0 = ((r0 & 0xFu) | ((r1 & 0xFu) << 4)) << shift;     // cx = 0
    uint32_t c1 = ((r0 >> 4) | ((r1 >> 4) << 4)) << shift;         // cx = 1
    bool upper = (lane >> 3) & 1u;                                 // vy >= 2 -> high word
    uint32_t c0lo = upper ? 0u : c0, c0hi = upper ? c0 : 0u, c1lo = upper ? 0u : c1, c1hi = upper ? c1 : 0u;
    // OR over the 8 lanes that share (cz, cy): lane bits 0, 2, 3
#pragma unroll
    for (int o = 1; o <= 8; o <<= 1) {
        if (o == 2) continue;
        c0lo |= __shfl_xor_sync(0xFFFFFFFFu, c0lo, o);
        c0hi |= __shfl_xor_sync(0xFFFFFFFFu, c0hi, o);
        c1lo |= __shfl_xor_sync(0xFFFFFFFFu, c1lo, o);
        c1hi |= __shfl_xor_sync(0xFFFFFFFFu, c1hi, o);
    }
    if ((lane & 0xDu) == 0u) {  // lanes 0, 2, 16, 18: one per (cz, cy)
        uint32_t cz = (lane >> 1) & 1u, cy = lane >> 4;
        uint2* c = cells + (size_t)slot * 8u + (cz << 1) + (cy << 2);
        c[0] = make_uint2(c0lo, c0hi);
        c[1] = make_uint2(c1lo, c1hi);
    }
}

// K_move: device-side relocation of resident bricks when a sector's slot range changes (the
// reference re-uploads the whole sector instead, BrickSlotAllocator.cpp:19-23).  One warp per
// brick: 512 + 64 bytes read and written.
__global__ void __launch_bounds__(256) k_move_bricks(const uint2* __restrict__ pairs, uint32_t n, uint8_t* voxels, uint2* cells) {
    uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (warp >= n) return;
    uint32_t lane = threadIdx.x & 31u;
    uint2 p = __ldg(pairs + warp);  // {src slot, dst slot}
    uint4 v = reinterpret_cast<const uint4*>(voxels + (size_t)p.x * 512u)[lane];
    uint2 c = make_uint2(0, 0);
    if (lane < 8) c = cells[(size_t)p.x * 8u + lane];
    reinterpret_cast<uint4*>(voxels + (size_t)p.y * 512u)[lane] = v;
    if (lane < 8) cells[(size_t)p.y * 8u + lane] = c;
}

// K_headers: scatter of the per-sector {allocMask, baseSlot} records that changed.
struct HeaderUpdate {
    uint32_t index, mask_lo, mask_hi, base;  // index = hdr_index() in the bordered grid
};
__global__ void k_write_headers(const HeaderUpdate* __restrict__ upd, uint32_t n, uint4* hdr) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    HeaderUpdate u = upd[i];
    // {allocMask, base slot, base slot of the upper 32 bricks}; an empty mask leaves z/w free for the box builder
    bool any = (u.mask_lo | u.mask_hi) != 0u;
    hdr[u.index] = make_uint4(u.mask_lo, u.mask_hi, any ? u.base : 0u, any ? u.base + (uint32_t)__popc(u.mask_lo) : 0u);
}

// K_init_headers: zero the in-view entries, mark the one-sector border OUTSIDE.  `hdr` points at grid
// entry 0; `guard` further OUTSIDE entries precede and follow the grid (the traversal loop does not clamp
// its index: DESIGN.md §5).
__global__ void k_init_headers(uint4* hdr, uint32_t sxp, uint32_t syp, uint32_t guard) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = sxp * sxp * syp;
    if (j >= n + 2u * guard) return;
    bool border = true;
    if (j >= guard && j < guard + n) {
        uint32_t i = j - guard;
        uint32_t x = i % sxp, z = (i / sxp) % sxp, y = i / (sxp * sxp);
        border = x == 0 || z == 0 || y == 0 || x == sxp - 1 || z == sxp - 1 || y == syp - 1;
    }
    (hdr - guard)[j] = make_uint4(0u, 0u, 0u, border ? VRT_HDR_OUTSIDE : 0u);
}

// K_occ: the one-bit-per-entry view of the header table that the step-by-step loop consults first (DevScene::occ).  One warp
// per 32 entries; bit set = resident sector, border or guard entry.  `hdr_all` points at the first guard entry.
__global__ void __launch_bounds__(256) k_build_occ(const uint4* __restrict__ hdr_all, uint32_t n_all, uint32_t* __restrict__ occ) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool bit = false;
    if (i < n_all) {
        const uint4 h = hdr_all[i];
        bit = (h.x | h.y) != 0u || (int)h.w < 0;
    }
    const unsigned word = __ballot_sync(0xFFFFFFFFu, bit);
    if ((threadIdx.x & 31u) == 0u && i < n_all) occ[i >> 5] = word;
}

// ---------------------------------------------------------------------------------------------
// Empty-box builder.  For every empty in-view sector: a box of empty sectors that
// contains it, grown greedily one slab at a time; slab emptiness is a 3-D summed-volume-table query
// over the bordered grid (sat index of sector coordinate c is c + 1; the border counts as empty but
// boxes never leave [0, extent)).  Rebuilt only when some sector's emptiness changed.
// ---------------------------------------------------------------------------------------------
__global__ void k_box_occupancy(const uint4* __restrict__ hdr, uint32_t* __restrict__ sat, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 h = hdr[i];
    sat[i] = ((h.x | h.y) != 0u) ? 1u : 0u;
}
// inclusive prefix sums along one axis: `lines` independent lines of `len` elements, element j of
// line l at base(l) + j * stride, base(l) = (l / inner) * outer_stride + (l % inner) * inner_stride
__global__ void k_box_scan(uint32_t* sat, uint32_t lines, uint32_t len, uint32_t stride, uint32_t inner, uint32_t inner_stride,
                           uint32_t outer_stride) {
    uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= lines) return;
    uint32_t* p = sat + (size_t)(l / inner) * outer_stride + (size_t)(l % inner) * inner_stride;
    uint32_t acc = 0;
    for (uint32_t j = 0; j < len; j++) {
        acc += p[(size_t)j * stride];
        p[(size_t)j * stride] = acc;
    }
}
__device__ __forceinline__ uint32_t sat_count(const uint32_t* __restrict__ sat, uint32_t sxp, uint32_t sxzp, int x0, int x1, int y0, int y1,
                                              int z0, int z1) {
    // sum over [x0,x1] x [y0,y1] x [z0,z1] (sector coords): indices c + 1 (high) and c (low - 1 + 1)
    auto at = [&](int X, int Y, int Z) -> uint32_t { return __ldg(sat + (size_t)X + (size_t)Z * sxp + (size_t)Y * sxzp); };
    int X1 = x1 + 1, Y1 = y1 + 1, Z1 = z1 + 1;
    return at(X1, Y1, Z1) - at(x0, Y1, Z1) - at(X1, y0, Z1) - at(X1, Y1, z0) + at(x0, y0, Z1) + at(x0, Y1, z0) + at(X1, y0, z0) -
           at(x0, y0, z0);
}
__global__ void __launch_bounds__(128) k_box_grow(uint4* __restrict__ hdr, const uint32_t* __restrict__ sat, uint32_t sxp, uint32_t sxzp,
                                                  int ext_xz, int ext_y) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n_view = (uint32_t)ext_xz * (uint32_t)ext_xz * (uint32_t)ext_y;
    if (i >= n_view) return;
    int x = (int)(i % (uint32_t)ext_xz), z = (int)((i / (uint32_t)ext_xz) % (uint32_t)ext_xz), y = (int)(i / ((uint32_t)ext_xz * (uint32_t)ext_xz));
    uint32_t hi = hdr_index(sxp, sxzp, x, y, z);
    uint4 h = hdr[hi];
    if ((h.x | h.y) != 0u) return;  // resident sector: z/w hold its slot base and popcount
    int x0 = x, x1 = x, y0 = y, y1 = y, z0 = z, z1 = z;
    for (bool grew = true; grew;) {
        grew = false;
        if (x1 + 1 < ext_xz && sat_count(sat, sxp, sxzp, x1 + 1, x1 + 1, y0, y1, z0, z1) == 0u) x1++, grew = true;
        if (x0 > 0 && sat_count(sat, sxp, sxzp, x0 - 1, x0 - 1, y0, y1, z0, z1) == 0u) x0--, grew = true;
        if (z1 + 1 < ext_xz && sat_count(sat, sxp, sxzp, x0, x1, y0, y1, z1 + 1, z1 + 1) == 0u) z1++, grew = true;
        if (z0 > 0 && sat_count(sat, sxp, sxzp, x0, x1, y0, y1, z0 - 1, z0 - 1) == 0u) z0--, grew = true;
        if (y1 + 1 < ext_y && sat_count(sat, sxp, sxzp, x0, x1, y1 + 1, y1 + 1, z0, z1) == 0u) y1++, grew = true;
        if (y0 > 0 && sat_count(sat, sxp, sxzp, x0, x1, y0 - 1, y0 - 1, z0, z1) == 0u) y0--, grew = true;
    }
    bool big = (x1 - x0) >= 2 || (y1 - y0) >= 2 || (z1 - z0) >= 2;  // worth a macro step: >= 3 sectors along some axis
    big = big && ext_xz <= 256 && ext_y <= 256;  // corners are stored in 8 bits per axis
    h.z = ((uint32_t)x0 & 0xFFu) | (((uint32_t)y0 & 0xFFu) << 8) | (((uint32_t)z0 & 0xFFu) << 16);
    h.w = ((uint32_t)x1 & 0xFFu) | (((uint32_t)y1 & 0xFFu) << 8) | (((uint32_t)z1 & 0xFFu) << 16) | (big ? VRT_HDR_HASBOX : 0u);
    hdr[hi] = h;
}

// ---------------------------------------------------------------------------------------------
// K_hit_query: VoxelMap::RayCast + GetStepLevel (VoxelMap.cpp:125-170) in fp64, one thread per
// query.  Steps 32 (no sector) / 8 (no brick) / 1, bias 1e-4.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int step_level(const DevScene& S, int x, int y, int z) {
    int sx = x >> 5, sy = y >> 5, sz = z >> 5;
    if (((uint32_t)(sx | sz) >> S.sxz) != 0u || ((uint32_t)sy >> S.sy) != 0u) return 5;
    uint4 h = ldg_hdr(S.hdr + hdr_index(S.sxp, S.sxzp, sx, sy, sz));
    if ((h.x | h.y) == 0u) return 5;
    uint32_t bi = ((uint32_t)(x >> 3) & 3u) | (((uint32_t)(z >> 3) & 3u) << 2) | (((uint32_t)(y >> 3) & 3u) << 4);
    uint32_t half = (bi & 32u) ? h.y : h.x;
    if (!((half >> (bi & 31u)) & 1u)) return 3;
    uint32_t vi = ((uint32_t)x & 7u) | (((uint32_t)z & 7u) << 3) | (((uint32_t)y & 7u) << 6);
    return __ldg(S.voxels + (size_t)brick_slot(h, bi) * 512u + vi) == 0 ? 0 : -1;
}
__device__ __forceinline__ double glm_min(double x, double y) { return (y < x) ? y : x; }
__device__ __forceinline__ int floor2i_d(double v) {
    double f = floor(v);
    return (f >= -2147483648.0 && f < 2147483648.0) ? (int)f : (int)0x80000000;
}

__global__ void __launch_bounds__(128) k_hit_query(const __grid_constant__ DevScene S, const double* __restrict__ origin3, const double* __restrict__ dir3,
                                                   uint32_t max_iters, uint64_t n, VrtHitD* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double o[3] = {origin3[3 * i], origin3[3 * i + 1], origin3[3 * i + 2]};
    double d[3] = {dir3[3 * i], dir3[3 * i + 1], dir3[3 * i + 2]};
    double inv[3], ts[3];
    int pos[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        inv[a] = __ddiv_rn(1.0, d[a]);
        ts[a] = __dmul_rn(__dsub_rn(d[a] < 0.0 ? 0.0 : 1.0, o[a]), inv[a]);
        pos[a] = floor2i_d(o[a]);
    }
    VrtHitD r;
    r.dist = -1.0;
    r.nx = r.ny = r.nz = r.u = r.v = 0.0f;
    r.vx = r.vy = r.vz = 0;
    r.iters = max_iters;
    r._pad = 0;
    for (uint32_t it = 0; it < max_iters; it++) {
        double sd[3], hp[3];
#pragma unroll
        for (int a = 0; a < 3; a++) sd[a] = __fma_rn((double)pos[a], inv[a], ts[a]);
        double tmin = __dadd_rn(glm_min(glm_min(sd[0], sd[1]), sd[2]), 0.0001);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            hp[a] = __fma_rn(tmin, d[a], o[a]);
            pos[a] = floor2i_d(hp[a]);
        }
        int k = step_level(S, pos[0], pos[1], pos[2]);
        if (k < 0) {
            bool smx = tmin >= sd[0], smy = tmin >= sd[1], smz = tmin >= sd[2];
            r.dist = tmin;
            r.nx = smx ? (float)-((d[0] > 0.0) - (d[0] < 0.0)) : 0.0f;
            r.ny = smy ? (float)-((d[1] > 0.0) - (d[1] < 0.0)) : 0.0f;
            r.nz = smz ? (float)-((d[2] > 0.0) - (d[2] < 0.0)) : 0.0f;
            float fu = smx ? (float)hp[1] : (float)hp[0];
            float fv = smz ? (float)hp[1] : (float)hp[2];
            r.u = __fsub_rn(fu, floorf(fu));
            r.v = __fsub_rn(fv, floorf(fv));
            r.vx = pos[0];
            r.vy = pos[1];
            r.vz = pos[2];
            r.iters = it + 1;
            break;
        }
        int mk = (1 << k) - 1;
#pragma unroll
        for (int a = 0; a < 3; a++) pos[a] = d[a] < 0.0 ? (pos[a] & ~mk) : (pos[a] | mk);
    }
    out[i] = r;
}

}  // namespace vrt
