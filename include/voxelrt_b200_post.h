/*
 * voxelrt_b200_post.h — C ABI of the image-space step that FOLLOWS the traversal path
 * (SURVEY.md §8f, row N4): the reference's `GBuffer` (src/VoxelRT/GBuffer.h) — blit of the tiled
 * framebuffer, temporal reprojection, SVGF variance estimate + à-trous passes, tone-mapped present —
 * as CUDA kernels for sm_100a in the same shared library (libvoxelrt_b200.so).
 *
 * The object mirrors the reference's `GBuffer` member for member: it owns the history between frames;
 * the caller hands it, per frame, the camera (GBuffer::SetCamera) and the 16 B/px tile framebuffer the
 * traversal path produced (vrt_render_device's default output, left in device memory).  All functions
 * return VRT_OK (0) or a negative VrtStatus (voxelrt_b200.h); there is no CPU fallback.
 */
#ifndef VOXELRT_B200_POST_H
#define VOXELRT_B200_POST_H

#include "voxelrt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct VrtGBuffer VrtGBuffer;

/* GBuffer::DebugChannel (GBuffer.h:8) */
typedef enum VrtDebugChannel {
    VRT_CHANNEL_NONE = 0,
    VRT_CHANNEL_ALBEDO = 1,
    VRT_CHANNEL_IRRADIANCE = 2,
    VRT_CHANNEL_NORMALS = 3,
    VRT_CHANNEL_TRAVERSAL_ITERS = 4, /* skips the denoiser (GBuffer.h:90) */
    VRT_CHANNEL_VARIANCE = 5
} VrtDebugChannel;

/* Arguments of GBuffer::SetCamera (GBuffer.h:31-60) as the shaders see them (SetUniforms, :61-84). */
typedef struct VrtGBufferCamera {
    uint32_t width, height;   /* multiples of 4 (CpuRenderer.cpp:419); a change of size drops the history (:32-47) */
    float proj[16];           /* CurrentProj = P * V(rotation only), column-major (GBuffer.h:51)                    */
    float inv_proj[16];       /* GetInverseProjScreenMat(CurrentProj, viewSize) (GBuffer.h:133-139)                 */
    double position[3];       /* cam.ViewPosition (GBuffer.h:50); u_OriginDelta = vec3(current - history) (:82)     */
    uint32_t reset_history;   /* u_ForceResetHistory = "the world changed this frame" (CpuRenderer.cpp:421-423)     */
    uint32_t _pad;
} VrtGBufferCamera;

/* Planes vrt_gbuffer_read can copy out (tests / inspection).  Irradiance planes are w*h records of 16 bytes:
 * {f16 r, f16 g, f16 b, f16 variance, f32 depth, u32 albedo|normal<<24} — the reference's IrradianceTex texel
 * with the DepthTex / AlbedoTex texels of the same frame stored next to it (one 16-byte load per filter tap). */
typedef enum VrtGBufferPlane {
    VRT_PLANE_IRRADIANCE = 0,      /* IrradianceTex     */
    VRT_PLANE_PREV_IRRADIANCE = 1, /* PrevIrradianceTex */
    VRT_PLANE_TEMP_IRRADIANCE = 2, /* TempIrradianceTex */
    VRT_PLANE_MOMENTS = 3,         /* MomentsTex, w*h x 2 f16 */
    VRT_PLANE_HISTORY_LEN = 4      /* HistoryLenTex, w*h u8   */
} VrtGBufferPlane;

/* GBuffer::GBuffer (GBuffer.h:25-29).  device = CUDA ordinal, -1 = current. */
VRT_API int vrt_gbuffer_create(int32_t device, VrtGBuffer** out);
VRT_API void vrt_gbuffer_destroy(VrtGBuffer* gb);
VRT_API const char* vrt_gbuffer_last_error(const VrtGBuffer* gb);

/* GBuffer::NumDenoiserPasses (0..5, default 5; GBuffer.h:23, slider CpuRenderer.cpp:480) and
 * GBuffer::DebugChannelView (GBuffer.h:22). */
VRT_API int vrt_gbuffer_set_passes(VrtGBuffer* gb, uint32_t num_passes);
VRT_API int vrt_gbuffer_set_debug_channel(VrtGBuffer* gb, uint32_t channel);

/* GBuffer::SetCamera (GBuffer.h:31-60): rotates the history (matrices, position, albedo/depth/moments planes),
 * advances FrameNo, (re)allocates on a size change. */
VRT_API int vrt_gbuffer_set_camera(VrtGBuffer* gb, const VrtGBufferCamera* cam);

/* The tail of RenderFrame (CpuRenderer.cpp:466-473): CopyTiledFramebuffer.comp + GBuffer::DenoiseAndPresent
 * (GBuffer.h:86-130).  `tiles` = width*height*16 bytes in Framebuffer::Tile layout (VrtTile, 4x4 tiles, row-major
 * tile order); `out_rgba8` = width*height u32, row-major, R in the low byte, A = 255 — what GBufferBlit.frag
 * writes to the window.  The _device form takes device pointers and runs on `stream` (cudaStream_t, may be NULL)
 * without synchronising; the host form copies in and out and returns when the image is in `out_rgba8`. */
VRT_API int vrt_gbuffer_denoise_present(VrtGBuffer* gb, const void* tiles, uint32_t* out_rgba8);
VRT_API int vrt_gbuffer_denoise_present_device(VrtGBuffer* gb, const void* d_tiles, uint32_t* d_out_rgba8, void* stream);

/* Renderer::RenderFrame from the trace to the window in one call (CpuRenderer.cpp:440-473): vrt_render_device of `frame`
 * on `ctx` into a tile framebuffer the GBuffer owns, then the blit + denoise + present above, all on one stream; only the
 * presented RGBA8 image (4 B/px instead of the 16 B/px tile framebuffer) is copied to host memory `out_rgba8`.  `ctx` must
 * live on the same device; vrt_gbuffer_set_camera must have been called with the frame's size. */
VRT_API int vrt_gbuffer_render_present(VrtGBuffer* gb, VrtContext* ctx, const VrtFrame* frame, uint32_t* out_rgba8);

/* Inspection hook: copy one plane (VrtGBufferPlane) to host memory; synchronises. */
VRT_API int vrt_gbuffer_read(VrtGBuffer* gb, uint32_t plane, void* out);
/* Kernels launched by the last denoise_present call. */
VRT_API int vrt_gbuffer_last_launches(const VrtGBuffer* gb, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* VOXELRT_B200_POST_H */
