/*
 * vrt_oracle.h — CPU ORACLE for the brickmap traversal path.  TEST INFRASTRUCTURE ONLY.
 *
 * A scalar, lane-wise restatement in plain C of the reference's CPU renderer
 * (src/VoxelRT/CpuRenderer.cpp) and picking ray cast (src/VoxelRT/VoxelMap.cpp).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product library (voxelrt_b200/lib/libvoxelrt_b200.so) never does.
 *
 * Parity status: PINNED for the traversal (a7-a10 of SURVEY §8a) against the reference's own
 * CpuRenderer.cpp compiled from /root/reference by oracle/Makefile (oracle/_ref/, see
 * oracle/ref_harness.cpp and tests/golden/); the reference ships no tests or golden vectors
 * of its own.  Ray generation and shading are PINNED too: whole frames of orc_render equal the reference's RenderRow output byte
 * for byte (tests/test_ref_pin.py) — rsqrt14 / rcp14 are reproduced bit-exactly and the 16-lane packet coupling of RayCast / RenderRow
 *  is part of orc_render's semantics.
 *
 * The structs are the public ABI's (include/voxelrt_b200.h) so tests feed identical bytes
 * to both sides.
 */
#ifndef VRT_ORACLE_H
#define VRT_ORACLE_H

#include "../include/voxelrt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcMap OrcMap;

/* per-call traversal statistics; these are the I_s / I_c / H terms of the algorithmic-bytes
 * formula: B_ray = 8 I_s + 8 I_c + 9 H + 16 P */
typedef struct OrcStats {
    uint64_t rays;
    uint64_t iters;          /* loop iterations started                                    */
    uint64_t sector_fetches; /* I_s: iterations that passed the bounds test                */
    uint64_t cell_fetches;   /* I_c: iterations whose brick bit was set                    */
    uint64_t hits;           /* H                                                          */
    uint64_t capped;
    uint64_t iter_hist[8];   /* 0-3,4-7,8-15,16-31,32-63,64-127,128-255,256+               */
    uint64_t lod_hist[6];    /* empty-cell steps by cell edge 1,2,4,8,16,32 voxels         */
} OrcStats;

OrcMap* orc_map_create(uint32_t sectors_xz_log2, uint32_t sectors_y_log2);
void orc_map_destroy(OrcMap* m);
void orc_map_set_palette(OrcMap* m, const uint64_t palette[256]);
/* FlatVoxelStorage::SyncBuffers, CpuRenderer.cpp:33-61 */
int orc_map_sync(OrcMap* m, uint32_t n, const VrtDirtySector* sectors);
/* FlatVoxelStorage::UpdateOccupancy, CpuRenderer.cpp:63-83 */
void orc_build_occupancy(const uint8_t brick[512], uint64_t cells[8]);
int orc_map_read_sector(const OrcMap* m, int32_t sx, int32_t sy, int32_t sz, uint64_t* alloc_mask, uint8_t* bricks,
                        uint64_t* cells);

void orc_set_blue_noise(OrcMap* m, const uint8_t* rg, size_t bytes);
void orc_set_sky(OrcMap* m, const VrtSkyDesc* desc, const uint32_t* texels);

/* RayCast, CpuRenderer.cpp:172-224 (lane-wise).  threads <= 0: all cores. */
void orc_trace(const OrcMap* m, uint64_t n, const float* origin3, const float* dir3, const int32_t world_origin[3],
               uint32_t max_iters, VrtHit* out, OrcStats* stats, int threads);
/* VoxelMap::RayCast, VoxelMap.cpp:125-170 */
void orc_hit_query(const OrcMap* m, uint64_t n, const double* origin3, const double* dir3, uint32_t max_iters,
                   VrtHitD* out, int threads);
/* RenderRow + RenderFrame row loop, CpuRenderer.cpp:326-402,455-462.
 * rows [row0,row1) in pixels (multiples of 4); pass 0,height for a full frame. */
void orc_render(const OrcMap* m, const VrtFrame* frame, void* out, VrtHit* aux_hits, OrcStats* stats, int threads,
                uint32_t row0, uint32_t row1);

/* primary ray of pixel (x,y): GetPrimaryRay + OriginFrac, CpuRenderer.cpp:226-233,327-334 */
void orc_primary_ray(const VrtFrame* frame, uint32_t x, uint32_t y, float origin[3], float dir[3]);
/* SampleDirection, CpuRenderer.cpp:273-291 */
void orc_sample_direction(float sx, float sy, float out[3]);
/* VBlueNoise::Sample for the pixel (x,y), CpuRenderer.cpp:254-270 (4x4 tiles) */
void orc_blue_noise_sample(const OrcMap* m, uint32_t x, uint32_t y, uint32_t frame_no, uint32_t sample_idx, float out[2]);
/* sky SampleCube<Nearest>(dir, mip) * 3, CpuRenderer.cpp:350-356 */
void orc_sky_sample(const OrcMap* m, const float dir[3], uint32_t mip, float out[3]);
/* R11G11B10f::Pack, Texture.h:158-177 */
uint32_t orc_pack_r11g11b10f(float r, float g, float b);
/* Material::GetEncoded, VoxelMap.h:27-41 */
uint64_t orc_encode_material(uint8_t r, uint8_t g, uint8_t b, uint8_t fuzz, float emission);

/* debugging aid: every ray pixel (x,y) casts (6 floats each) with its hit record; returns the count */
uint32_t orc_debug_pixel(const OrcMap* m, const VrtFrame* frame, uint32_t x, uint32_t y, float* rays6, VrtHit* hits, uint32_t out4[4]);

/* RGBA8u::Pack / RG16f::Pack (Texture.h:41-62,101-116), simd::sincos_2pi (SIMD.h:175-190) */
uint32_t orc_pack_unorm8x4(float r, float g, float b, float a);
uint32_t orc_pack_half2(float x, float y);
void orc_sincos_2pi(float x, float* s, float* c);

/* rayCast / rayCastCoarse + getStepPos of the GLSL renderer (Shaders/VoxelTraversal.glsl:92-243); flags = VRT_GLSL_*.
 * Parity unpinned (GLSL needs a GL device); canonical arithmetic in vrt_oracle.c. */
void orc_trace_glsl(const OrcMap* m, uint64_t n, const float* origin3, const float* dir3, const int32_t world_origin[3], uint32_t flags,
                    VrtHit* out, OrcStats* stats, int threads);
/* The GPU renderer's frame shader (Shaders/VoxelRender.comp:29-93) per pixel: VRT_FRAME_GLSL frames (include/voxelrt_b200.h).
 * Parity unpinned as above; canonical arithmetic and stand-ins in vrt_oracle.c. */
void orc_render_glsl(const OrcMap* m, const VrtFrame* frame, void* out, OrcStats* stats, int threads);
/* GenerateRayCellInteractionMaskLUT, GpuRenderer.cpp:193-210 */
void orc_interaction_lut(uint64_t table[512]);

/* _mm512_rsqrt14_ps / _mm512_rcp14_ps (SIMD_AVX512.h:136-138), bit-exact for every argument */
float orc_x86_rsqrt14(float x);
float orc_x86_rcp14(float x);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
