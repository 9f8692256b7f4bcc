// libemu_render.so — the per-pixel body of the FRAME kernels (voxelrt_b200/csrc/vrt_shade.cuh: primary_ray, cast_ray, shade_pixel_primary
// for bounces = 0 and shade_pixel for frames with blue-noise bounces and the sky cube) compiled for the host and run warp tile by warp tile
// (8x4 pixels = 32 lanes in LOCKSTEP: a half-warp is one 4x4 packet of the reference, and the packet coupling is carried by warp votes); the
// frame constants come from the product's own fill_frame_params.  Output: the 16 B/px tile framebuffer (VrtTile) like store_pixel writes
// it.  Macro steps off (they need the box builder).  TEST INFRASTRUCTURE ONLY.
#define VRT_HOST_EMULATION 1
#include "cuda_host_shim.h"
#include "../../voxelrt_b200/csrc/vrt_shade.cuh"

using namespace vrt;

extern "C" {
#define EMU_API __attribute__((visibility("default")))

struct EmuScene {
    const uint4* hdr;
    const uint2* cells;
    const uint8_t* voxels;
    const uint2* palette;
    uint32_t sxz, sy;
};

static const uint32_t* g_occ = nullptr;  // non-null: frames with bounces run the OCC instantiation (big views), bit index = header index + guard
EMU_API void emu_render_set_occ(const uint32_t* occ) { g_occ = occ; }

EMU_API void emu_render(const EmuScene* e, const VrtFrame* f, const uint8_t* blue_noise, const uint32_t* sky, const VrtSkyDesc* sky_desc, VrtTile* out,
                        VrtHit* aux) {
    DevScene S{};
    S.hdr = e->hdr, S.cells = e->cells, S.voxels = e->voxels, S.palette = e->palette;
    S.sxz = e->sxz, S.sy = e->sy;
    S.lim_xz = 32u << e->sxz, S.lim_y = 32u << e->sy;
    S.sxp = (1u << e->sxz) + 2, S.sxzp = S.sxp * S.sxp;
    S.n_hdr = S.sxzp * ((1u << e->sy) + 2);
    uint32_t albedo[256];  // k_palette_albedo: albedo_rgb_bits of every palette entry
    for (int i = 0; i < 256; i++) albedo[i] = albedo_rgb_bits(e->palette[i].x);
    S.albedo = albedo;
    S.occ = g_occ, S.occ_bias = 2u * S.sxzp;
    FrameParams F;
    fill_frame_params(F, f, S.sxp, 0, blue_noise, sky, sky_desc);
    F.aux = aux;
    const int w = (int)f->width, h = (int)f->height;
    const int tiles_x = (w + 7) / 8, tiles_y = (h + 3) / 4;
    const bool occ = g_occ != nullptr;
#pragma omp parallel for schedule(dynamic, 4)
    for (int t = 0; t < tiles_x * tiles_y; t++) {
        const int x0 = (t % tiles_x) * 8, y0 = (t / tiles_x) * 4;
        run_warp_lockstep((unsigned)t, 32, 0, [&](int lane) {
            const int x = x0 + ((lane >> 4) << 2) + (lane & 3), y = y0 + ((lane >> 2) & 3);  // render_warp_tile's lane -> pixel map
            const bool valid = x < w && y < h;
            PixelOut P;
            if (F.bounces == 0) shade_pixel_primary<false>(S, F, (uint32_t)x, (uint32_t)y, valid, P);
            else if (occ) shade_pixel<false, true>(S, F, (uint32_t)x, (uint32_t)y, valid, P);
            else shade_pixel<false, false>(S, F, (uint32_t)x, (uint32_t)y, valid, P);
            if (!valid) return;
            VrtTile* tile = out + ((size_t)(y >> 2) * (size_t)(w >> 2) + (size_t)(x >> 2));
            const int l = (x & 3) | ((y & 3) << 2);
            tile->albedo[l] = P.albedo;
            tile->depth[l] = P.depth;
            tile->irr_rg[l] = P.irr_rg;
            tile->irr_bx[l] = P.irr_bx;
        });
    }
}
}
