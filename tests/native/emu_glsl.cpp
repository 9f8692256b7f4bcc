// libemu_glsl.so — voxelrt_b200/csrc/vrt_glsl.cuh (k_build_groups, k_trace_glsl: the GLSL renderer's ray casts) compiled for the
// host and run thread by thread over device-layout arrays that tests/test_glsl_kernel_on_cpu.py builds with numpy, so that the CPU
// test tier compares the kernel source itself with the oracle.  TEST INFRASTRUCTURE ONLY.
#define VRT_HOST_EMULATION 1
#include "cuda_host_shim.h"
#include "../../voxelrt_b200/csrc/vrt_glsl.cuh"

using namespace vrt;

template <class F>
static void run_1d(uint64_t n, unsigned block, F kernel) {
    const int64_t blocks = (int64_t)((n + block - 1) / block);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t b = 0; b < blocks; b++) {
        blockDim.x = block;
        blockIdx.x = (unsigned)b;
        for (unsigned t = 0; t < block; t++) {
            threadIdx.x = t;
            kernel();
        }
    }
}

extern "C" {
#define EMU_API __attribute__((visibility("default")))

struct EmuScene {  // what vrt_api.cu's dev_scene() hands the kernels
    const uint4* hdr;
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
EMU_API void emu_build_groups(const EmuScene* e, uint2* groups) {
    const DevScene S = scene_of(e);
    const uint32_t gxz = e->sxz - 2, n_groups = 1u << (2 * gxz + e->sy - 2);
    run_1d(n_groups, 128, [&] { k_build_groups(S, groups, gxz, n_groups); });
}
EMU_API void emu_trace_glsl(const EmuScene* e, const uint2* groups, const uint2* lut, const int32_t wo[3], const float* o3, const float* d3,
                            uint32_t flags, uint64_t n, VrtHit* out) {
    const DevScene S = scene_of(e);
    const GlslScene G{groups, lut, e->sxz - 2};
    run_1d(n, 128, [&] { k_trace_glsl(S, G, wo[0], wo[1], wo[2], o3, d3, flags, n, out); });
}
}
