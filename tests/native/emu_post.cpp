// libemu_post.so — the DEVICE half of voxelrt_b200/csrc/vrt_post.cu (the GBuffer kernels) compiled for the host and run over the
// whole grid in a loop, so that tests/test_post_kernels_on_cpu.py can compare the very kernel source with the image-space oracle in
// the CPU test tier.  TEST INFRASTRUCTURE ONLY (nothing in the product links this).
#define VRT_HOST_EMULATION 1
#include "cuda_host_shim.h"
#include "../../voxelrt_b200/csrc/vrt_post.cu"

using namespace vrtpost;

template <class F>
static void run_grid(int w, int h, F kernel) {
    const int gx = (w + 31) / 32, gy = (h + 7) / 8;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < gx * gy; b++) {
        blockDim.x = 32, blockDim.y = 8;
        blockIdx.x = (unsigned)(b % gx), blockIdx.y = (unsigned)(b / gx);
        for (unsigned ty = 0; ty < 8; ty++)
            for (unsigned tx = 0; tx < 32; tx++) {
                threadIdx.x = tx, threadIdx.y = ty;
                kernel();
            }
    }
}

extern "C" {
#define EMU_API __attribute__((visibility("default")))

EMU_API void emu_reproject(const uint32_t* tiles, const uint4* prev, const uint32_t* prev_moments, const uint8_t* hist_in, uint4* out,
                           uint32_t* moments, uint8_t* hist_out, const float* cur_inv, const float* hist_proj, const float* hist_inv,
                           const float* delta, int w, int h, int reset) {
    FrameParams P;
    std::memcpy(P.cur_inv.m, cur_inv, 64);
    std::memcpy(P.hist_proj.m, hist_proj, 64);
    std::memcpy(P.hist_inv.m, hist_inv, 64);
    std::memcpy(P.delta, delta, 12);
    P.w = w, P.h = h, P.reset = reset;
    run_grid(w, h, [&] { k_reproject(tiles, prev, prev_moments, hist_in, out, moments, hist_out, P); });
}
EMU_API void emu_variance(const uint4* in, const uint8_t* hist, uint4* io_temp, int w, int h) {
    const TapRcp tr = tap_rcp(3, 1);
    run_grid(w, h, [&] { k_variance(in, hist, io_temp, w, h, tr); });
}
EMU_API void emu_atrous(const uint4* in, uint4* out, int w, int h, int pass) {
    const TapRcp tr = tap_rcp(2, 1 << pass);
    switch (pass) {
    case 0: run_grid(w, h, [&] { k_atrous<0>(in, out, w, h, tr); }); break;
    case 1: run_grid(w, h, [&] { k_atrous<1>(in, out, w, h, tr); }); break;
    case 2: run_grid(w, h, [&] { k_atrous<2>(in, out, w, h, tr); }); break;
    case 3: run_grid(w, h, [&] { k_atrous<3>(in, out, w, h, tr); }); break;
    default: run_grid(w, h, [&] { k_atrous<4>(in, out, w, h, tr); }); break;
    }
}
EMU_API void emu_present(const uint32_t* tiles, const uint4* irr, uint32_t* rgba, int w, int h, int channel) {
    run_grid(w, h, [&] { k_present(tiles, irr, rgba, w, h, channel); });
}
EMU_API void emu_blit_only(const uint32_t* tiles, uint4* irr, uint4* prev, int w, int h) {
    run_grid(w, h, [&] { k_blit_only(tiles, irr, prev, w, h); });
}
}
