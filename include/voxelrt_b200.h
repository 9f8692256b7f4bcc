/*
 * voxelrt_b200.h — C ABI of the B200-native brickmap traversal path.
 *
 * This is the drop-in boundary for the ONE hot path of dubiousconst282/VoxelRT that this
 * repository accelerates: brickmap residency -> per-ray hierarchical-DDA traversal ->
 * primary / blue-noise secondary shading -> screen-tile multi-GPU.  Plain pointers and
 * sizes only; no C++/torch types.  Every entry point names the reference interface it
 * replaces (paths relative to the reference tree, `src/VoxelRT/...`).
 *
 * Threading: a context is externally single-threaded (the reference calls everything from
 * its one UI/GL thread, Main.cpp:83-127).  All functions return VRT_OK (0) or a negative
 * VrtStatus; the message is available from vrt_last_error().  There is NO CPU fallback:
 * if no CUDA device is usable, vrt_create fails with VRT_ERR_CUDA.
 */
#ifndef VOXELRT_B200_H
#define VOXELRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VRT_API
#else
#define VRT_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------------------------------
 * Fixed geometry of the reference brickmap (VoxelMap.h:100-102, CpuRenderer.cpp:14-18)
 * ------------------------------------------------------------------------------------------ */
#define VRT_BRICK_DIM 8u          /* BrickIndexer<3,3>: 8x8x8 voxels, 1 byte per voxel            */
#define VRT_BRICK_BYTES 512u      /* sizeof(Brick)                                                */
#define VRT_SECTOR_BRICKS 64u     /* MaskIndexer<2,2>: 4x4x4 bricks per sector                    */
#define VRT_SECTOR_DIM 32u        /* voxels per sector edge                                       */
#define VRT_CELLS_PER_BRICK 8u    /* 2x2x2 cells of 4x4x4 voxels, one u64 occupancy mask each     */
#define VRT_PALETTE_SIZE 256u
#define VRT_MAX_ITERS_DEFAULT 128u /* CpuRenderer.cpp:185                                         */
#define VRT_BLUE_NOISE_BYTES (128u * 128u * 64u * 2u) /* STBN vec2 128x128x64, (R,G) u8 pairs     */

typedef enum VrtStatus {
    VRT_OK = 0,
    VRT_ERR_INVALID = -1,   /* bad argument                                                       */
    VRT_ERR_CUDA = -2,      /* CUDA runtime/driver error (message holds cudaGetErrorString)       */
    VRT_ERR_OOM = -3,       /* "Could not allocate brick slots" (BrickSlotAllocator.cpp:15)       */
    VRT_ERR_STATE = -4,     /* call order violated (e.g. render before any sync)                  */
    VRT_ERR_UNSUPPORTED = -5
} VrtStatus;

typedef struct VrtContext VrtContext;

/* View = the window of the world the renderer keeps resident.  The reference hard-codes it:
 * CPU renderer 64x16x64 sectors (CpuRenderer.cpp:14-16), GPU renderer 128x64x128
 * (GpuRenderer.cpp:6-8).  Here it is a runtime power-of-two extent, origin at sector (0,0,0). */
typedef struct VrtConfig {
    uint32_t struct_size;            /* = sizeof(VrtConfig)                                       */
    int32_t device;                  /* CUDA ordinal, -1 = current device                         */
    uint32_t sectors_xz_log2;        /* view extent in sectors along X and Z (6 = reference CPU)  */
    uint32_t sectors_y_log2;         /* view extent in sectors along Y       (4 = reference CPU)  */
    uint32_t initial_brick_capacity; /* brick slots to reserve up front; grows on demand (0=auto) */
    uint32_t flags;                  /* reserved, 0                                               */
} VrtConfig;

/* One entry of VoxelMap::DirtyLocs (VoxelMap.h:185) as the renderer sees it inside
 * SyncBuffers (CpuRenderer.cpp:38-59, GpuRenderer.cpp:52-78): the sector position, the
 * sector's current allocation mask (Sector::GetAllocationMask, VoxelMap.cpp:53-66), which of
 * its bricks changed, and the voxel bytes of the changed bricks.
 * `bricks` holds popcount(dirty_mask & alloc_mask) bricks of 512 bytes, ascending bit order,
 * voxel index = x | z<<3 | y<<6 (BrickIndexer).  A record with VRT_SECTOR_REMOVED means the
 * sector key is gone from VoxelMap::Sectors (CpuRenderer.cpp:43-46). */
#define VRT_SECTOR_REMOVED 1u
typedef struct VrtDirtySector {
    int32_t sx, sy, sz;  /* sector position (WorldSectorIndexer::GetPos), view-relative       */
    uint32_t flags;
    uint64_t alloc_mask; /* bit i = brick i allocated, i = x | z<<2 | y<<4 (MaskIndexer)       */
    uint64_t dirty_mask;
    const uint8_t* bricks;
} VrtDirtySector;

/* Result of one ray — the lane-wise content of VHitResult (CpuRenderer.cpp:88-96,209-223)
 * plus the hit voxel and bookkeeping the parity tests need.  48 bytes. */
typedef struct VrtHit {
    int32_t vx, vy, vz; /* voxel the loop stopped in, world coords (CpuRenderer.cpp:186)      */
    uint32_t material;  /* low 32 bits of the palette entry: RGB565 | f16 emission << 16      */
    float dist;         /* min3(sideDist), no bias (CpuRenderer.cpp:204)                      */
    float px, py, pz;   /* currPos, relative to worldOrigin (CpuRenderer.cpp:212)             */
    float u, v;         /* face UV (CpuRenderer.cpp:218-221)                                  */
    uint32_t flags;     /* VRT_HIT_* below                                                    */
    uint32_t _pad;
} VrtHit;
/* flags layout */
#define VRT_HIT_NORMAL_MASK 0x3Fu   /* (nx+1) | (ny+1)<<2 | (nz+1)<<4  (CpuRenderer.cpp:372-374) */
#define VRT_HIT_HIT 0x100u          /* VHitResult::Mask: stopped on a solid voxel inside the view */
#define VRT_HIT_INBOUND 0x200u      /* last voxel was inside the view                             */
#define VRT_HIT_CAPPED 0x400u       /* ran out of iterations (a miss, CpuRenderer.cpp:222)        */
#define VRT_HIT_ITERS_SHIFT 16      /* loop iterations executed, bits 16..31                      */

/* Result of the fp64 picking query, VoxelMap::RayCast (VoxelMap.cpp:140-170) / HitResult
 * (VoxelMap.h:171-178). */
typedef struct VrtHitD {
    double dist;        /* tmin at the hit, -1 on miss                                            */
    float nx, ny, nz;   /* mix(0, -sign(dir), tmin >= sideDist)                                   */
    float u, v;
    int32_t vx, vy, vz;
    uint32_t iters;
    uint32_t _pad;
} VrtHitD;

/* Per-frame constants — FrameConstants of the CPU renderer (CpuRenderer.cpp:314-323,444-453).
 * Matrices are column-major float[16] exactly like glm::mat4 (m[col][row] = a[col*4+row]). */
#define VRT_FRAME_LINEAR_OUTPUT 1u /* out = 4 planes (albedo, depth, irrRG, irrBX) of w*h u32  */
#define VRT_FRAME_AUX_HITS 2u      /* also fill the VrtHit of every primary ray (aux_hits)     */
#define VRT_FRAME_COMPACT 4u       /* bounces == 0 only: 8 B/px — per 4x4 tile {albedo[16], depth[16]} (VrtTileAD; with
                                    * VRT_FRAME_LINEAR_OUTPUT: the two planes).  A primary-only frame's irradiance is the
                                    * constant 1.0 in every pixel (CpuRenderer.cpp:379-381: IrradianceRG = IrradianceBX =
                                    * 0x3C003C00), so half of the reference's 16 B/px tile carries no information; leaving it
                                    * out halves the bytes over PCIe (vrt_render) and NVLink (vrt_render_gather), losslessly */
#define VRT_FRAME_PART_ROWS 8u     /* multi-GPU split by BAND: rank r renders the VRT_BAND_ROWS-pixel-high
                                    * bands b with b % part_count == r (each band is one contiguous range
                                    * of the tile-layout framebuffer, so a rank's result moves with one
                                    * strided copy); default split is by macro tile, t % part_count       */
#define VRT_FRAME_GLSL 16u         /* render the frame like the reference's GPU renderer: main() of Shaders/VoxelRender.comp:29-93
                                    * per pixel (getPrimaryRay :20-25, rayCast, the sun shadow ray, `bounces` = u_MaxBounces
                                    * diffuse bounces through rayCastCoarse with a sun ray after the first two, sky on a miss),
                                    * i.e. GpuRenderer::RenderFrame's compute dispatch (GpuRenderer.cpp:257-268) instead of
                                    * CpuRenderer::RenderFrame's RenderRow.  Same frame constants, same 16 B/px tile output —
                                    * albedo | packGNormal << 24, depth, irradiance as 3 halfs, and the 4th half = hit.iters
                                    * (what the shader stores in IrradianceTex.w) — so every way of delivering a frame (host /
                                    * device buffers, band split, gather, the GBuffer step) applies.  Needs a view of at least
                                    * 4 sectors per axis (the 128^3 level), vrt_set_blue_noise for bounces > 0.  PARITY
                                    * UNPINNED (no GL device): bit-equal to its own oracle under the canonical arithmetic and
                                    * the two stand-ins (sky sampler, unassigned `out` fields) stated in vrt_glsl_frame.cuh  */
#define VRT_FRAME_GLSL_ANISOTROPIC 32u /* with VRT_FRAME_GLSL: u_UseAnisotropicLods (GpuRenderer.cpp:262)              */
#define VRT_BAND_ROWS 8u           /* two tile rows: 2.7 % load imbalance at 8 GPUs on the 4K terrain frame
                                    * (32-pixel bands: 14 %)                                              */
typedef struct VrtFrame {
    uint32_t width, height;  /* rounded down to multiples of 4 by the caller (CpuRenderer.cpp:419) */
    float inv_proj[16];      /* GBuffer::GetInverseProjScreenMat (GBuffer.h:133-139)               */
    float proj[16];          /* CurrentProj = P * V(rotation only) (GBuffer.h:51)                  */
    int32_t world_origin[3]; /* floor(camera position)                                             */
    float origin_frac[3];    /* fract(camera position)                                             */
    uint32_t frame_no;
    uint32_t bounces;        /* NumLightBounces (Renderer.h:65)                                    */
    uint32_t max_iters;      /* 0 = 128                                                            */
    uint32_t flags;
    /* screen-tile partition for multi-GPU: this context renders tiles t with
     * t % part_count == part_index (tile = 32x32 pixels, row-major tile index).  1-GPU: 0/1. */
    uint32_t part_index, part_count;
} VrtFrame;

/* Framebuffer::Tile of the reference (CpuRenderer.cpp:299-309) for 16-lane packets: a 4x4
 * pixel tile stored as four arrays of 16 u32, lane = (x&3) | (y&3)<<2.  Default output of
 * vrt_render is tiles in row-major tile order, TileStride = width/4. */
typedef struct VrtTile {
    uint32_t albedo[16];  /* RGBA8, A = packed normal code << 0 (bits 24..29)                  */
    float depth[16];      /* proj z/w of pos/16, -1 on miss                                    */
    uint32_t irr_rg[16];  /* 2 x f16                                                           */
    uint32_t irr_bx[16];  /* f16 | 0                                                           */
} VrtTile;

/* VRT_FRAME_COMPACT: the information-carrying half of a VrtTile of a primary-only frame. */
typedef struct VrtTileAD {
    uint32_t albedo[16];
    float depth[16];
} VrtTileAD;

typedef struct VrtStats {
    uint64_t resident_bricks;   /* FreeList::NumAllocated                                      */
    uint64_t brick_capacity;    /* slots in the device arena                                   */
    uint64_t free_ranges;       /* FreeList::FreeRanges.size() (GpuRenderer.cpp:301)           */
    uint64_t resident_sectors;
    uint64_t bytes_uploaded;    /* H2D bytes moved by the last vrt_sync                        */
    uint64_t bricks_uploaded;   /* bricks in the last vrt_sync                                 */
    uint64_t bricks_relocated;  /* bricks moved device-side by the last vrt_sync               */
    uint64_t device_bytes;      /* total device memory owned by the context                    */
    uint64_t last_launches;     /* kernels launched by the last ABI call                       */
    uint64_t bounce_form;       /* how the last frame with bounces was traced: 0 none yet, 1 one thread per pixel (all levels in one
                                   kernel), 2 wavefront passes; bit 8 set while the per-context choice between the two is still being
                                   timed (the first four such frames of a (size, bounces, scene) combination)                     */
} VrtStats;

/* ---- lifetime ------------------------------------------------------------------------------- */
/* Replaces CpuRenderer::CpuRenderer / GpuRenderer::GpuRenderer storage setup
 * (CpuRenderer.cpp:26-31,404-411; GpuRenderer.cpp:41-43,212-236). */
VRT_API int vrt_create(const VrtConfig* cfg, VrtContext** out);
VRT_API void vrt_destroy(VrtContext* ctx);
/* Message of the last failing call on ctx (ctx may be NULL for a failed vrt_create). */
VRT_API const char* vrt_last_error(const VrtContext* ctx);
VRT_API int vrt_get_stats(const VrtContext* ctx, VrtStats* out);

/* ---- residency ------------------------------------------------------------------------------ */
/* Palette[i] = Material::GetEncoded() (VoxelMap.h:27-41); the reference re-encodes it every
 * frame (CpuRenderer.cpp:34-36, GpuRenderer.cpp:100-102). */
VRT_API int vrt_set_palette(VrtContext* ctx, const uint64_t palette[256]);
/* FlatVoxelStorage::SyncBuffers (CpuRenderer.cpp:33-61) / GpuVoxelStorage::SyncBuffers
 * (GpuRenderer.cpp:45-167): allocate/free brick slots, upload ONLY the dirty bricks, rebuild
 * their 8 cell masks on the device (UpdateOccupancy, CpuRenderer.cpp:63-83 /
 * UpdateOccupancy.comp:8-34).  Host data is consumed before return.
 * - Each sector may appear in at most ONE record per call (DirtyLocs is a map keyed by sector, VoxelMap.h:185);
 *   a duplicate returns VRT_ERR_INVALID.
 * - The call is transactional: a record that fails validation (dirty bricks without payload, duplicate sector)
 *   or a failed allocation (VRT_ERR_OOM) returns before anything — host mirror, slot arena, device buffers —
 *   has changed; vrt_read_sector and rendering see the state of the previous successful call.
 * - A brick that enters alloc_mask without being in dirty_mask (VoxelMap::GetBrick creates bricks on lookup,
 *   VoxelMap.cpp:122) is resident as an EMPTY brick (its recycled slot is zero-filled). */
VRT_API int vrt_sync(VrtContext* ctx, uint32_t n, const VrtDirtySector* sectors);
/* Test/inspection hook: copy the resident state of one sector back to the host.
 * out_alloc_mask: allocation mask; out_bricks (64*512 B) / out_cells (64*8 u64) are filled for
 * allocated bricks in brick-index order (dense, unallocated = 0). Any pointer may be NULL. */
VRT_API int vrt_read_sector(VrtContext* ctx, int32_t sx, int32_t sy, int32_t sz, uint64_t* out_alloc_mask,
                            uint32_t* out_base_slot, uint8_t* out_bricks, uint64_t* out_cells);

/* ---- ray cast (explicit rays) ----------------------------------------------------------------- */
/* RayCast(map, origin, dir, mask, worldOrigin) -> VHitResult (CpuRenderer.cpp:172-224), one
 * ray per element.  origin3/dir3 are n x 3 floats (origin relative to world_origin).
 * Host-pointer version copies in and out; the _device version takes device pointers and
 * runs on `stream` (a cudaStream_t, may be NULL). */
VRT_API int vrt_trace(VrtContext* ctx, uint64_t n, const float* origin3, const float* dir3, const int32_t world_origin[3],
                      uint32_t max_iters, VrtHit* out);
VRT_API int vrt_trace_device(VrtContext* ctx, uint64_t n, const float* d_origin3, const float* d_dir3,
                             const int32_t world_origin[3], uint32_t max_iters, VrtHit* d_out, void* stream);

/* ---- ray cast with the GPU renderer's semantics ------------------------------- */
/* rayCast / rayCastCoarse of the reference's GLSL renderer (Shaders/VoxelTraversal.glsl:162-243) with its
 * getStepPos (:92-131): the extra 128^3 level (SectorMasks, one bit per sector of a 4x4x4 group), origins
 * outside the grid clipped to its box (:133-145), +5-ulp bias and 256 iterations — or, with VRT_GLSL_COARSE,
 * rayCastCoarse: +0.001 bias, 96 iterations, after 30 of them any occupied 4^3 cell counts as a hit (what the
 * GPU renderer uses for bounce and sun-shadow rays, VoxelRender.comp:58-83).  VRT_GLSL_ANISOTROPIC masks every
 * occupancy word with the ray/cell interaction LUT first (u_UseAnisotropicLods, GpuRenderer.cpp:193-210).
 * These rays visit other cells than the CPU renderer's, so normals / iteration counts differ from vrt_trace;
 * VrtHit fields: dist = biased tmin, flags carry -sign(dir) normals (code 21 = none on a miss), u/v as the shader. */
#define VRT_GLSL_COARSE 1u
#define VRT_GLSL_ANISOTROPIC 2u
VRT_API int vrt_trace_glsl(VrtContext* ctx, uint64_t n, const float* origin3, const float* dir3, const int32_t world_origin[3],
                           uint32_t flags, VrtHit* out);

/* ---- hit query (picking) ---------------------------------------------------------------------- */
/* HitResult VoxelMap::RayCast(dvec3 origin, dvec3 dir, maxIters=1024) (VoxelMap.cpp:140-170),
 * batched.  The reference walks the unbounded sector hash; here sectors outside the resident
 * view read as "no sector" (step 32). */
VRT_API int vrt_hit_query(VrtContext* ctx, uint64_t n, const double* origin3, const double* dir3, uint32_t max_iters,
                          VrtHitD* out);

/* ---- shading inputs --------------------------------------------------------------------------- */
/* VBlueNoise source image (CpuRenderer.cpp:235-252): 128 x (128*64) texels of (R,G) bytes,
 * row-major, slice s occupies rows [128 s, 128 s + 128). */
VRT_API int vrt_set_blue_noise(VrtContext* ctx, const uint8_t* rg, size_t bytes);
/* Sky cube as swr::HdrTexture2D holds it (Texture.h:401-446): 6 layers of R11G11B10f texels
 * with a mip chain; texel(layer, mip, x, y) = data[(layer << layer_shift) + mip_offset[mip]
 * + x + (y << (row_shift - mip))]. */
typedef struct VrtSkyDesc {
    uint32_t face_size;   /* power of two                                                      */
    uint32_t mip_levels;  /* <= 16                                                             */
    uint32_t layer_shift;
    uint32_t mip_offset[16];
    uint64_t texel_count; /* u32 elements in `texels`                                          */
} VrtSkyDesc;
VRT_API int vrt_set_sky(VrtContext* ctx, const VrtSkyDesc* desc, const uint32_t* texels);

/* ---- frame ------------------------------------------------------------------------------------ */
/* Renderer::RenderFrame minus present (Renderer.h:18; CpuRenderer.cpp:415-464 — everything
 * between SyncBuffers and the blit).  vrt_render writes w*h*16 bytes (w*h*8 with VRT_FRAME_COMPACT) to HOST
 * memory `out` (tiles, or planes with VRT_FRAME_LINEAR_OUTPUT).  With a screen split (part_count > 1 and
 * VRT_FRAME_PART_ROWS) only this rank's 8-pixel bands are traced and written, at their offsets of the whole frame:
 * N processes can fill ONE host frame (e.g. a shared, page-locked mapping) in parallel, each over its own PCIe link.
 * aux_hits (host, w*h VrtHit, row-major
 * pixel order) is filled when VRT_FRAME_AUX_HITS is set.  vrt_render_device leaves the result
 * in device memory on `stream` — this is what a CUDA/GL-interop presenter would consume. */
VRT_API int vrt_render(VrtContext* ctx, const VrtFrame* frame, void* out, VrtHit* aux_hits);
VRT_API int vrt_render_device(VrtContext* ctx, const VrtFrame* frame, void* d_out, VrtHit* d_aux_hits, void* stream);

/* ---- multi-GPU band gather over NVLink (no reference counterpart) ------------------ */
/* One process per GPU, the brickmap replicated, the screen split into 8-pixel bands (VRT_FRAME_PART_ROWS).  The presenting rank exports
 * its framebuffer as a CUDA IPC handle (vrt_fb_export) and every other rank opens it (vrt_fb_import).  The exchange is done by the COPY
 * ENGINES, not by stores of the frame kernel: vrt_render_gather renders this rank's bands into its OWN device buffer d_local_fb on
 * `stream`, then copies exactly those bands to the same offsets of d_owner_fb (the presenting rank's framebuffer, or a local pointer) on a
 * copy stream of the context — one strided device-to-device copy over NVLink — while the SMs trace the next frame.  (Why not peer
 * stores from the kernel's epilogue, as SURVEY 8e sketched: with all ranks in step, every rank would write into ONE GPU during a frame —
 * 7/8 of the frame through one 770 GB/s ingress, longer than the frame takes to trace — and the stores' back-pressure stalls the tracing
 * warps; the copy engines decouple the two)  The call returns at once; a d_local_fb may be reused every
 * VRT_GATHER_DEPTH-th call (the call waits for the copy issued that many calls ago), the owner may change from frame to frame, and when
 * d_local_fb == d_owner_fb (the presenting rank renders its own bands in place) no copy is issued.  With VRT_FRAME_COMPACT the bands are
 * 8 B/px.  vrt_gather_wait makes `stream` wait for every gather issued so far.  64-byte opaque handle. */
VRT_API int vrt_fb_export(VrtContext* ctx, uint64_t bytes, uint8_t handle_out[64], void** d_ptr_out);
VRT_API int vrt_fb_import(VrtContext* ctx, const uint8_t handle[64], void** d_ptr_out);
VRT_API int vrt_fb_release(VrtContext* ctx, void* d_ptr);
#define VRT_GATHER_DEPTH 8u
VRT_API int vrt_render_gather(VrtContext* ctx, const VrtFrame* frame, void* d_local_fb, void* d_owner_fb, void* stream);
VRT_API int vrt_gather_wait(VrtContext* ctx, void* stream);

/* Device-side traversal counters of the last trace/render (TRAVERSAL_METRICS,
 * VoxelTraversal.glsl:147-151): total loop iterations, sector-mask fetches, cell-mask fetches,
 * hits.  Enabled with vrt_set_option(ctx, "metrics", 1); costs a few atomics per warp. */
typedef struct VrtTraversalMetrics {
    uint64_t rays, iters, sector_fetches, cell_fetches, hits, capped;
} VrtTraversalMetrics;
VRT_API int vrt_get_metrics(VrtContext* ctx, VrtTraversalMetrics* out);
/* Settings (the role of Renderer::DrawSettings, Renderer.h:19).  Every combination produces the same bytes; they change speed only.
 *   "metrics" 0/1          traversal counters (above)
 *   "macro_steps" 0/1      exact empty-box space skipping for camera rays, default 1 (2 = diagnostic counters)
 *   "wavefront" 0/1/2      frames with bounces: one thread per pixel through all levels / wavefront passes (camera pass, then per level a
 *                          persistent trace pass with lane refill and a shade pass) / default 2: the first four such frames after a change
 *                          of scene emptiness, size, bounce count or split alternate between the forms, timed, and the faster is kept
 *                          (VrtStats.bounce_form reports it)
 *   "trace_refill" 1..32   trace pass: lanes in flight below which a warp refills from the queue, default 24
 *   "trace_ctas" 8/10/12   trace pass: resident CTAs per SM it is compiled for, default 10
 *   "tile_order" 0/1       primary frame kernel: warp tiles top-to-bottom / bottom-to-top (default: the cheap sky rows run last)
 *   "persistent" 0/n       resident-grid form of the frame kernel with n x (SMs x CTAs/SM) CTAs, default 0 (measured slower)
 *   "gather_threads" 0..64 host threads of vrt_sync's staging gather for batches of >= 8192 bricks: 0 = min(8, cores), 1 = caller only
 *   "compact_bounces"      removed (round 1's CTA-level re-dealing of bounce rays): any non-zero value returns VRT_ERR_UNSUPPORTED
 *   "reserve_slots" n      test hook: take the first n brick slots out of an EMPTY arena (exercises 64-bit brick addressing)
 * Unknown names return VRT_ERR_INVALID. */
VRT_API int vrt_set_option(VrtContext* ctx, const char* name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* VOXELRT_B200_H */
