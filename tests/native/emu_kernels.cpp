// libemu_kernels.so — per-thread kernels of voxelrt_b200/csrc/vrt_kernels.cuh compiled for the host: K_hit_query (the fp64 picking ray of
// VoxelMap::RayCast) and the header / empty-box builders (k_init_headers, k_write_headers, k_box_occupancy, k_box_scan, k_box_grow), run
// thread by thread; with the boxes in place the traversal can be emulated WITH macro steps.  TEST INFRASTRUCTURE ONLY.
#define VRT_HOST_EMULATION 1
#include "cuda_host_shim.h"
#include "../../voxelrt_b200/csrc/vrt_kernels.cuh"
#include "../../voxelrt_b200/csrc/vrt_glsl_frame.cuh"

using namespace vrt;

template <class F>
static void run_1d(uint64_t n, unsigned block, F kernel) {
    const int64_t blocks = (int64_t)((n + block - 1) / block);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t b = 0; b < blocks; b++) {
        blockDim.x = block, blockDim.y = blockDim.z = 1;
        blockIdx.x = (unsigned)b;
        for (unsigned t = 0; t < block; t++) {
            threadIdx.x = t;
            kernel();
        }
    }
}

extern "C" {
#define EMU_API __attribute__((visibility("default")))

struct EmuScene {
    uint4* hdr;  // entry 0 of the bordered grid; 2 * sxp^2 guard entries precede and follow it
    const uint2* cells;
    const uint8_t* voxels;
    const uint2* palette;
    uint32_t sxz, sy;
};
static DevScene scene_of(const EmuScene* e) {
    DevScene S{};
    S.hdr = e->hdr, S.cells = e->cells, S.voxels = e->voxels, S.palette = e->palette;
    S.sxz = e->sxz, S.sy = e->sy;
    S.lim_xz = 32u << e->sxz, S.lim_y = 32u << e->sy;
    S.sxp = (1u << e->sxz) + 2, S.sxzp = S.sxp * S.sxp;
    S.n_hdr = S.sxzp * ((1u << e->sy) + 2);
    return S;
}

EMU_API void emu_hit_query(const EmuScene* e, const double* o3, const double* d3, uint32_t max_iters, uint64_t n, VrtHitD* out) {
    const DevScene S = scene_of(e);
    run_1d(n, 128, [&] { k_hit_query(S, o3, d3, max_iters, n, out); });
}

// vrt_create + vrt_sync's header path: k_init_headers, then one HeaderUpdate per resident sector through k_write_headers
EMU_API void emu_build_headers(const EmuScene* e, const uint32_t* index, const uint32_t* mask_lo, const uint32_t* mask_hi, const uint32_t* base, uint32_t n) {
    const DevScene S = scene_of(e);
    const uint32_t syp = (1u << e->sy) + 2, guard = 2u * S.sxzp;
    run_1d((uint64_t)S.n_hdr + 2u * guard, 256, [&] { k_init_headers(e->hdr, S.sxp, syp, guard); });
    HeaderUpdate* upd = new HeaderUpdate[n ? n : 1];
    for (uint32_t i = 0; i < n; i++) upd[i] = HeaderUpdate{index[i], mask_lo[i], mask_hi[i], base[i]};
    run_1d(n, 128, [&] { k_write_headers(upd, n, e->hdr); });
    delete[] upd;
}

// K_upload (K2, one WARP per brick: voxel copy + the eight 4x4x4 occupancy masks through xor-shuffle OR-reductions), by warp replay
EMU_API void emu_upload_bricks(const uint4* staging, const uint32_t* slots, uint32_t n, uint8_t* voxels, uint2* cells) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t w = 0; w < (int64_t)n; w++) {
        ShflReplay R;
        std::memset(&R, 0, sizeof(R));
        g_shfl_replay = &R;
        blockDim.x = 32, blockDim.y = blockDim.z = 1;
        blockIdx.x = (unsigned)w;
        int calls = 0;
        for (int round = 0; round <= 64; round++) {
            for (unsigned lane = 0; lane < 32; lane++) {
                threadIdx.x = lane;
                R.lane = (int)lane, R.call = 0;
                k_upload_bricks(staging, slots, n, voxels, cells);
                calls = R.call;
            }
            std::memcpy(R.prev, R.cur, sizeof(R.prev));
            if (round >= calls) break;  // every call has seen correct inputs
        }
        g_shfl_replay = nullptr;
    }
}

// K_move: device-side relocation of resident bricks ({src slot, dst slot} pairs, one warp per brick, no lane exchange).  All reads of a launch
// precede its writes on the GPU only per warp, so (as in vrt_sync) sources and destinations of one launch must not overlap.
EMU_API void emu_move_bricks(const uint2* pairs, uint32_t n, uint8_t* voxels, uint2* cells) {
    run_1d((uint64_t)n * 32u, 256, [&] { k_move_bricks(pairs, n, voxels, cells); });
}

// k_build_occ (one ballot per warp of 32 header entries) by warp replay: the one-bit-per-entry table of the OCC traversal loop
EMU_API void emu_build_occ(const uint4* hdr_all, uint32_t n_all, uint32_t* occ) {
    const int64_t warps = ((int64_t)n_all + 31) / 32;
#pragma omp parallel for schedule(static)
    for (int64_t w = 0; w < warps; w++) {
        ShflReplay R;
        std::memset(&R, 0, sizeof(R));
        g_shfl_replay = &R;
        blockDim.x = 32, blockDim.y = blockDim.z = 1;
        blockIdx.x = (unsigned)w;
        for (int round = 0; round < 2; round++) {
            for (unsigned lane = 0; lane < 32; lane++) {
                threadIdx.x = lane;
                R.lane = (int)lane, R.call = 0;
                k_build_occ(hdr_all, n_all, occ);
            }
            std::memcpy(R.prev, R.cur, sizeof(R.prev));
        }
        g_shfl_replay = nullptr;
    }
}

// k_render itself (one-thread-per-pixel frame kernel: warp tiles of 8x4 pixels numbered inside macro tiles or bands, the multi-GPU screen
// split, store_pixel), launched like launch_render does: fill_frame_params + fill_frame_partition, blocks of VRT_RENDER_THREADS threads.
EMU_API int emu_render_kernel(const EmuScene* e, const VrtFrame* f, const uint8_t* bn, const uint32_t* sky, const VrtSkyDesc* sky_desc, void* out,
                              uint32_t row0, uint32_t row1) {
    DevScene S = scene_of(e);
    uint32_t albedo[256];
    for (int i = 0; i < 256; i++) albedo[i] = albedo_rgb_bits(e->palette[i].x);
    S.albedo = albedo;
    FrameParams F;
    fill_frame_params(F, f, S.sxp, 0, bn, sky, sky_desc);
    F.out = out;
    if (!fill_frame_partition(F, f, row0, row1)) return -1;
    if (F.n_work <= F.work_offset) return 0;
    static int order = 0;  // alternate the two grid orders of launch_render ("tile_order"): results must not depend on it
    order ^= 1;
    F.work_add = order ? F.n_work - 1u : F.work_offset;
    F.work_mul = order ? -1 : 1;
    const unsigned wpb = VRT_RENDER_THREADS / 32;
    const int64_t blocks = (F.n_work - F.work_offset + wpb - 1) / wpb;
    const bool rows = (F.flags & VRT_FRAME_PART_ROWS) != 0u, primary = F.bounces == 0;
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t w = 0; w < blocks * (int64_t)wpb; w++) {  // every warp of the grid, its 32 lanes in lockstep (packet votes)
        run_warp_lockstep((unsigned)(w / wpb), VRT_RENDER_THREADS, (unsigned)(w % wpb) * 32u, [&](int) {
            if (primary) {
                if (rows) k_render<false, true, true>(S, F);
                else k_render<false, true, false>(S, F);
            } else {
                if (rows) k_render<false, false, true>(S, F);
                else k_render<false, false, false>(S, F);
            }
        });
    }
    return (int)blocks;
}

// k_render_glsl (VRT_FRAME_GLSL frames: the GPU renderer's frame shader per pixel), launched like launch_render does.  groups / lut: the
// 128^3 level and the interaction LUT as vrt_api.cu's glsl_scene() prepares them (k_build_groups is run here).
EMU_API int emu_render_glsl(const EmuScene* e, const VrtFrame* f, const uint8_t* bn, const uint32_t* sky, const VrtSkyDesc* sky_desc, const uint2* lut, void* out) {
    DevScene S = scene_of(e);
    if (e->sxz < 2 || e->sy < 2) return -2;
    const uint32_t gxz = e->sxz - 2, n_groups = 1u << (2 * gxz + e->sy - 2);
    std::vector<uint2> groups(n_groups);
    gridDim.x = (n_groups + 127) / 128, blockDim.x = 128;
    for (unsigned b = 0; b < gridDim.x; b++)
        for (unsigned t = 0; t < 128; t++) {
            blockIdx.x = b, threadIdx.x = t;
            k_build_groups(S, groups.data(), gxz, n_groups);
        }
    const GlslScene G{groups.data(), lut, gxz};
    FrameParams F;
    fill_frame_params(F, f, S.sxp, 0, bn, sky, sky_desc);
    F.out = out;
    if (!fill_frame_partition(F, f, 0, 0)) return -1;
    if (F.n_work <= F.work_offset) return 0;
    const unsigned wpb = VRT_RENDER_THREADS / 32;
    const int64_t blocks = (F.n_work - F.work_offset + wpb - 1) / wpb;
    const bool rows = (F.flags & VRT_FRAME_PART_ROWS) != 0u;
    const uint32_t cast_flags = (F.flags & VRT_FRAME_GLSL_ANISOTROPIC) ? VRT_GLSL_ANISOTROPIC : 0u;
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t w = 0; w < blocks * (int64_t)wpb; w++)
        run_warp_lockstep((unsigned)(w / wpb), VRT_RENDER_THREADS, (unsigned)(w % wpb) * 32u, [&](int) {
            if (rows) k_render_glsl<true>(S, G, F, cast_flags);
            else k_render_glsl<false>(S, G, F, cast_flags);
        });
    return (int)blocks;
}

// The WAVEFRONT form of a frame with bounces, launched like launch_render does: k_wave_primary, then per bounce level k_wave_trace (the
// persistent, lane-refilling trace pass with the branch-lean trip) and k_wave_shade.  Warps run in lockstep; the trace pass runs `trace_warps`
// persistent warps one after the other (the first ones drain most of the queue — results do not depend on who traces a ray).
EMU_API int emu_wave_frame(const EmuScene* e, const VrtFrame* f, const uint8_t* bn, const uint32_t* sky, const VrtSkyDesc* sky_desc, void* out, VrtHit* aux,
                           uint32_t trace_warps) {
    DevScene S = scene_of(e);
    uint32_t albedo[256];
    for (int i = 0; i < 256; i++) albedo[i] = albedo_rgb_bits(e->palette[i].x);
    S.albedo = albedo;
    FrameParams F;
    fill_frame_params(F, f, S.sxp, 0, bn, sky, sky_desc);
    F.out = out;
    F.aux = aux;
    if (!fill_frame_partition(F, f, 0, 0)) return -1;
    if (F.n_work <= F.work_offset || F.bounces == 0) return 0;
    static int wave_order = 0;  // both grid orders of the camera pass ("tile_order")
    wave_order ^= 1;
    F.work_add = wave_order ? F.n_work - 1u : F.work_offset;
    F.work_mul = wave_order ? -1 : 1;
    const unsigned wpb = VRT_RENDER_THREADS / 32;
    const int64_t warps = (int64_t)((F.n_work - F.work_offset + wpb - 1) / wpb) * wpb;
    const bool rows = (F.flags & VRT_FRAME_PART_ROWS) != 0u;
    const size_t cap = (size_t)(F.n_work - F.work_offset) * 32u;
    std::vector<RayRec> rays(cap);
    std::vector<HitRec> hits(cap);
    std::vector<float4> path_a(cap);
    std::vector<float2> path_b(cap);
    std::vector<uint16_t> pk(cap / 16 + 1);
    uint32_t counters[30] = {};
    WaveBuffers B{rays.data(), (uint32_t)cap, counters, counters + 20, counters + 10, hits.data(), path_a.data(), path_b.data(), pk.data()};
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t w = 0; w < warps; w++)
        run_warp_lockstep((unsigned)(w / wpb), VRT_RENDER_THREADS, (unsigned)(w % wpb) * 32u, [&](int) {
            if (rows) k_wave_primary<true>(S, F, B);
            else k_wave_primary<false>(S, F, B);
        });
    for (uint32_t level = 1; level <= F.bounces; level++) {
        TraceArgs A{B.rays, B.n_rays + level, B.head + level, B.hits, F.max_iters, 24u, B.n_generic + level, B.capacity};
#pragma omp parallel for schedule(dynamic, 1)
        for (int64_t w = 0; w < (int64_t)trace_warps; w++)
            run_warp_lockstep((unsigned)(w / wpb), VRT_RENDER_THREADS, (unsigned)(w % wpb) * 32u, [&](int) { k_wave_trace<8>(S, F.W, A); });
        {
            gridDim.x = 1, blockDim.x = 128, blockIdx.x = 0;
            for (unsigned t = 0; t < 128; t++) {
                threadIdx.x = t;
                k_wave_trace_generic(S, F.W, A);
            }
        }
#pragma omp parallel for schedule(dynamic, 8)
        for (int64_t w = 0; w < warps; w++)
            run_warp_lockstep((unsigned)(w / wpb), VRT_RENDER_THREADS, (unsigned)(w % wpb) * 32u, [&](int) {
                if (rows) k_wave_shade<true>(S, F, B, level);
                else k_wave_shade<false>(S, F, B, level);
            });
    }
    return (int)(counters[1] + counters[21]);
}

// rebuild_boxes of vrt_api.cu: the same five launches, in order
EMU_API void emu_build_boxes(const EmuScene* e, uint32_t* sat) {
    const DevScene S = scene_of(e);
    const uint32_t sxp = S.sxp, syp = (1u << e->sy) + 2, sxzp = S.sxzp, n = S.n_hdr;
    run_1d(n, 256, [&] { k_box_occupancy(e->hdr, sat, n); });
    run_1d(sxp * syp, 128, [&] { k_box_scan(sat, sxp * syp, sxp, 1u, sxp * syp, sxp, 0u); });
    run_1d(sxp * syp, 128, [&] { k_box_scan(sat, sxp * syp, sxp, sxp, sxp, 1u, sxzp); });
    run_1d(sxzp, 128, [&] { k_box_scan(sat, sxzp, syp, sxzp, sxzp, 1u, 0u); });
    const uint32_t n_view = 1u << (2 * e->sxz + e->sy);
    run_1d(n_view, 128, [&] { k_box_grow(e->hdr, sat, sxp, sxzp, 1 << e->sxz, 1 << e->sy); });
}
}
