// libemu_trace.so — the traversal itself (voxelrt_b200/csrc/vrt_device.cuh: cast_ray = the magic-number fast loop or the generic loop,
// cast_finish, store_hit — what k_trace runs per thread) compiled for the host and run one lane at a time over a device-layout
// brickmap built by the test with numpy.  The PTX helpers have plain-C twins under VRT_HOST_EMULATION (lop3, the pinning no-ops,
// 1/x for the rcp + Newton sequence); the directed-rounding adds go through fesetround.  Macro steps are off (they need the box
// builder's data); everything else is the loop the 20 Grays/s kernel runs.  TEST INFRASTRUCTURE ONLY.
#define VRT_HOST_EMULATION 1
#include "cuda_host_shim.h"
#include "../../voxelrt_b200/csrc/vrt_device.cuh"

using namespace vrt;

extern "C" {
#define EMU_API __attribute__((visibility("default")))

struct EmuScene {
    const uint4* hdr;  // entry 0 of the bordered grid; 2 * sxp^2 OUTSIDE guard entries precede and follow it
    const uint2* cells;
    const uint8_t* voxels;
    const uint2* palette;
    uint32_t sxz, sy;
};

// mode 0: cast_ray as k_trace calls it, macro steps off; 1: force the generic loop (what special rays take);
// 2: macro steps ON (the headers must carry the boxes of the box builder, tests/native/emu_kernels.cpp);
// 3: the OCC form of the step-by-step loop (big views: a one-bit sector table is consulted before the header);
// 4: the METRICS instantiation (what bench.py's roofline counters come from); n_fast then receives 6 words: fast rays, iterations, sector-mask
//    fetches, cell-mask fetches, hits, capped rays (DevMetrics, summed here instead of through the warp-aggregated atomics)
static const uint32_t* g_occ = nullptr;  // mode 3: the one-bit table of k_build_occ (bit index = header index + guard)
EMU_API void emu_set_occ(const uint32_t* occ) { g_occ = occ; }

EMU_API void emu_trace(const EmuScene* e, const int32_t wo[3], const float* o3, const float* d3, uint32_t max_iters, uint64_t n, VrtHit* out,
                       int mode, uint64_t* n_fast) {
    DevScene S{};
    S.hdr = e->hdr, S.cells = e->cells, S.voxels = e->voxels, S.palette = e->palette;
    S.sxz = e->sxz, S.sy = e->sy;
    S.lim_xz = 32u << e->sxz, S.lim_y = 32u << e->sy;
    S.sxp = (1u << e->sxz) + 2, S.sxzp = S.sxp * S.sxp;
    S.n_hdr = S.sxzp * ((1u << e->sy) + 2);
    S.occ = g_occ, S.occ_bias = 2u * S.sxzp;
    RayFrame W = make_ray_frame(S.sxp, mode == 2 ? 1 : 0, wo);
    if (mode == 1) W.fast_ok = 0;
    if (max_iters == 0) max_iters = VRT_MAX_ITERS_DEFAULT;
    uint64_t fast = 0, m_it = 0, m_ns = 0, m_nc = 0, m_hit = 0, m_cap = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : fast, m_it, m_ns, m_nc, m_hit, m_cap)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        blockDim.x = 128, blockDim.y = blockDim.z = 1;
        threadIdx.x = (unsigned)(i & 127);
        HitLane H;
        CastResult R;
        R.iters = R.n_sector = R.n_cell = 0;
        R.capped = false;
        H.hit = false;
        const float ox = o3[3 * i], oy = o3[3 * i + 1], oz = o3[3 * i + 2], dx = d3[3 * i], dy = d3[3 * i + 1], dz = d3[3 * i + 2];
        fast += (W.fast_ok && ray_is_fast(ox, oy, oz, dx, dy, dz)) ? 1 : 0;
        if (mode == 4) {
            cast_ray<true>(S, W, ox, oy, oz, dx, dy, dz, max_iters, H, R);
            m_it += R.iters, m_ns += R.n_sector, m_nc += R.n_cell, m_hit += H.hit ? 1 : 0, m_cap += R.capped ? 1 : 0;
        } else if (mode == 3) cast_ray<false, true, true>(S, W, ox, oy, oz, dx, dy, dz, max_iters, H, R);
        else cast_ray<false>(S, W, ox, oy, oz, dx, dy, dz, max_iters, H, R);
        store_hit(out + i, H, R);
    }
    if (n_fast) {
        n_fast[0] = fast;
        if (mode == 4) n_fast[1] = m_it, n_fast[2] = m_ns, n_fast[3] = m_nc, n_fast[4] = m_hit, n_fast[5] = m_cap;
    }
}
}
