// oracle/_ref/libref_cpu.so — the REFERENCE's own CPU renderer, compiled from where it lies.
//
// TEST INFRASTRUCTURE ONLY (never linked into or loaded by the product).  This translation unit
// #includes the reference's src/VoxelRT/CpuRenderer.cpp and VoxelMap.cpp verbatim from
// /root/reference (nothing is copied into this repository) so that its file-static functions —
// FlatVoxelStorage::SyncBuffers/UpdateOccupancy, GetStepPos, RayCast, GetPrimaryRay, VBlueNoise,
// SampleDirection, RenderRow — and VoxelMap::RayCast can be called directly.  The libraries the
// reference takes from vcpkg (glm, imgui, glad, stb; absent from this image) are replaced by the
// small stand-ins under oracle/shim/, which are our own code; none of them is on the traversal's
// arithmetic path (the traversal uses only the reference's SIMD.h / SIMD_AVX512.h intrinsics).
//
// Built by oracle/Makefile with  g++ -O2 -std=c++20 -march=native -fno-fast-math  — i.e. the
// reference's flags minus -ffast-math (SURVEY §8a "canonical arithmetic": value-changing fast-math
// rewrites such as 1/x -> rcp14+Newton are a build artefact we deliberately do not pin).
// Used for: (1) pinning the oracle restatement (tests/test_ref_pin.py, tests/golden/), (2) the
// "reference" kind of bench.py's cpu_baseline / --impl reference arm.
#include <omp.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "../include/voxelrt_b200.h"

// ---- the reference, verbatim ---------------------------------------------------------------------
#include "CpuRenderer.cpp"
#include "VoxelMap.cpp"
// scene ingest: the reference's surface voxeliser and palette quantiser, VoxelMap::VoxelizeModel
// (Voxelize.cpp:77-147 + Common/PaletteBuilder.h).  Its model loader (Common/Scene.cpp) sits on assimp + stb_image, which are absent: the
// glim::Model arrives through ref_voxelize below instead, already decoded.
#include <cfloat>
#include <functional>
#include "Common/Scene.h"
#include "Voxelize.cpp"
glim::Model::Model(std::string_view) {}
// the editing tool that produces BASELINE configs[4]'s dirty bricks in the app: BrushSession::Dispatch (Brush.cpp:10-37) over
// VoxelMap::RegionDispatchSIMD (VoxelMap.h:223-263: brick creation, DirtyLocs, garbage collection of emptied bricks / sectors)
#include "Brush.cpp"

// ---- externals the two TUs expect from files we do not compile ------------------------------------
namespace swr {
// LibGlimpsw/SwRast/ImageHelpers.cpp:13-29 decodes a PNG with stb; here the static VBlueNoise ctor
// (CpuRenderer.cpp:238-252) gets a zero image and the real table arrives through ref_set_blue_noise.
StbImage StbImage::Load(std::string_view, PixelType type) {
    const uint32_t w = 128, h = 128 * 64;
    return {.Width = w, .Height = h, .Type = type, .Data = {(uint8_t*)std::calloc((size_t)w * h, 4), &std::free}};
}
namespace texutil {
// ImageHelpers.cpp:87-122 builds the cube from an .hdr panorama (stb); the static _skyBox
// (CpuRenderer.cpp:311) starts as a tiny black cube and ref_set_sky installs the real texels.
HdrTexture2D LoadCubemapFromPanoramaHDR(std::string_view, uint32_t mipLevels) { return HdrTexture2D(8, 8, mipLevels, 6); }
}  // namespace texutil
}  // namespace swr
// The "cvox 0004" (de)serialiser (VoxelMap.cpp:205-274) sits on the reference's Common/BinaryIO.cpp, which oracle/Makefile
// compiles as a second translation unit straight from /root/reference (its header has no include guard): <zstd.h> resolves to
// oracle/shim/zstd.h (declarations only) and the system libzstd.so.1 is linked.

// ---- C API -----------------------------------------------------------------------------------------
#define REF_API extern "C" __attribute__((visibility("default")))

struct RefCtx {
    VoxelMap map;
    std::unique_ptr<FlatVoxelStorage> storage;
    uint64_t palette[256] = {};
};

REF_API RefCtx* ref_create() {
    auto* c = new RefCtx();
    c->storage = std::make_unique<FlatVoxelStorage>();  // the reference's dense 2 GiB + 256 MiB view
    return c;
}
REF_API void ref_destroy(RefCtx* c) { delete c; }
REF_API int ref_vector_width() { return (int)simd::VectorWidth; }
REF_API int ref_num_threads() { return omp_get_max_threads(); }

REF_API void ref_set_palette(RefCtx* c, const uint64_t* enc) {
    std::memcpy(c->palette, enc, sizeof(c->palette));
    std::memcpy(c->storage->Palette, enc, sizeof(c->palette));
}
// Material::GetEncoded (VoxelMap.h:27-41)
REF_API uint64_t ref_encode_material(uint8_t r, uint8_t g, uint8_t b, uint8_t fuzz, float emission) {
    Material m;
    m.Color[0] = r, m.Color[1] = g, m.Color[2] = b;
    m.MetalFuzziness = fuzz;
    m.Emission = emission;
    return m.GetEncoded();
}

// Applies dirty-sector records to the reference VoxelMap (Sectors + DirtyLocs), then runs the
// reference's FlatVoxelStorage::SyncBuffers (CpuRenderer.cpp:33-61).
REF_API int ref_sync(RefCtx* c, uint32_t n, const VrtDirtySector* recs) {
    for (uint32_t r = 0; r < n; r++) {
        const VrtDirtySector& d = recs[r];
        glm::ivec3 pos(d.sx, d.sy, d.sz);
        if (!WorldSectorIndexer::CheckInBounds(pos)) continue;
        uint32_t idx = WorldSectorIndexer::GetIndex(pos);
        if (d.flags & VRT_SECTOR_REMOVED) {
            c->map.Sectors.erase(idx);
            c->map.DirtyLocs[idx] = ~0ull;
            continue;
        }
        Sector& s = c->map.Sectors[idx];
        s.DeleteBricks(s.GetAllocationMask() & ~d.alloc_mask);
        const uint8_t* src = d.bricks;
        for (uint32_t b = 0; b < 64; b++) {
            if (!((d.alloc_mask >> b) & 1)) continue;
            Brick* brick = s.GetBrick(b, true);
            if ((d.dirty_mask >> b) & 1) {
                std::memcpy(brick->Data, src, 512);
                src += 512;
            }
        }
        c->map.DirtyLocs[idx] |= d.dirty_mask;
    }
    c->storage->SyncBuffers(c->map);
    std::memcpy(c->storage->Palette, c->palette, sizeof(c->palette));  // SyncBuffers re-encodes map.Palette (unused here)
    return 0;
}

// One brush stroke on the reference's VoxelMap: capsule from a to b (voxel coordinates), action 0 = Fill / 1 = Replace, material 0 erases.
REF_API void ref_brush_dispatch(RefCtx* c, const int32_t a[3], const int32_t b[3], float radius, int action, uint8_t material) {
    BrushSession s;
    s.Pars.Action = action ? BrushAction::Replace : BrushAction::Fill;
    s.Pars.Radius = radius;
    s.Pars.PointA = glm::ivec3(a[0], a[1], a[2]);
    s.Pars.PointB = glm::ivec3(b[0], b[1], b[2]);
    s.Pars.Material = Voxel{material};
    s.Dispatch(c->map);
}
// VoxelMap::DirtyLocs as (sector position, dirty-brick mask) pairs; clears it (what SyncBuffers does after consuming it)
REF_API uint32_t ref_map_take_dirty(RefCtx* c, int32_t* xyz, uint64_t* masks, uint32_t cap) {
    uint32_t n = 0;
    for (auto& [idx, mask] : c->map.DirtyLocs) {
        if (n >= cap) break;
        glm::ivec3 p = WorldSectorIndexer::GetPos(idx);
        xyz[3 * n] = p.x, xyz[3 * n + 1] = p.y, xyz[3 * n + 2] = p.z;
        masks[n++] = mask;
    }
    c->map.DirtyLocs.clear();
    return n;
}

// ---- cvox files: the reference's own VoxelMap::Serialize / Deserialize (VoxelMap.cpp:205-274) ----
REF_API int ref_serialize(RefCtx* c, const char* path) {
    try {
        c->map.Serialize(path);
        return 0;
    } catch (const std::exception&) { return -1; }
}
REF_API int ref_deserialize(RefCtx* c, const char* path) {
    try {
        c->map.Sectors.clear();
        c->map.Deserialize(path);
        return 0;
    } catch (const std::exception&) { return -1; }
}
REF_API void ref_set_material(RefCtx* c, int i, uint8_t r, uint8_t g, uint8_t b, uint8_t fuzz, float emission) {
    Material& m = c->map.Palette[i & 255];
    m.Color[0] = r, m.Color[1] = g, m.Color[2] = b;
    m.MetalFuzziness = fuzz;
    m.Emission = emission;
}
REF_API void ref_get_material(RefCtx* c, int i, uint8_t* rgbf4, float* emission) {
    const Material& m = c->map.Palette[i & 255];
    rgbf4[0] = m.Color[0], rgbf4[1] = m.Color[1], rgbf4[2] = m.Color[2], rgbf4[3] = m.MetalFuzziness;
    *emission = m.Emission;
}
REF_API uint32_t ref_map_sector_count(RefCtx* c) { return (uint32_t)c->map.Sectors.size(); }
// sectors of the VoxelMap itself (not the renderer's view), in hash-map order: world sector coordinates + allocation mask
REF_API uint32_t ref_map_list_sectors(RefCtx* c, int32_t* xyz, uint64_t* masks, uint32_t cap) {
    uint32_t n = 0;
    for (auto& [idx, sector] : c->map.Sectors) {
        if (n >= cap) break;
        glm::ivec3 p = WorldSectorIndexer::GetPos(idx);
        xyz[3 * n] = p.x, xyz[3 * n + 1] = p.y, xyz[3 * n + 2] = p.z;
        masks[n++] = sector.GetAllocationMask();
    }
    return n;
}
REF_API int ref_map_read_sector(RefCtx* c, int sx, int sy, int sz, uint64_t* mask, uint8_t* bricks /* popcount(mask) x 512, ascending */) {
    auto it = c->map.Sectors.find(WorldSectorIndexer::GetIndex(glm::ivec3(sx, sy, sz)));
    if (it == c->map.Sectors.end()) return -1;
    uint64_t m = it->second.GetAllocationMask();
    *mask = m;
    for (uint32_t b = 0; b < 64; b++)
        if ((m >> b) & 1) {
            std::memcpy(bricks, it->second.GetBrick(b)->Data, 512);
            bricks += 512;
        }
    return 0;
}

// glim::PaletteBuilder on its own: AddColor for n packed RGBA8 words, Build(max_colors); -> NumColors, colours in rgb[3 * i ..], and (when
// `index_of` is given) FindIndex of every input colour.
REF_API uint32_t ref_palette_build(const uint32_t* colors, uint64_t n, uint32_t max_colors, uint8_t* rgb, uint8_t* index_of) {
    glim::PaletteBuilder pb;
    for (uint64_t i = 0; i < n; i++) pb.AddColor(colors[i]);
    pb.Build(max_colors);
    for (uint32_t i = 0; i < pb.NumColors; i++) rgb[3 * i] = pb.ColorR[i], rgb[3 * i + 1] = pb.ColorG[i], rgb[3 * i + 2] = pb.ColorB[i];
    if (index_of)
        for (uint64_t i = 0; i < n; i++) index_of[i] = (uint8_t)pb.FindIndex(colors[i]);
    return pb.NumColors;
}

// VoxelMap::VoxelizeModel on a model handed over as plain arrays: triangles (3 x xyz, model space, node transforms already applied),
// their texture coordinates, a texture id per triangle, and the decoded RGBA8 base-colour images (power-of-two sizes, as swr::Texture2D
// requires).  Builds the glim::Model the way Common/Scene.cpp does (one texture with 8 mips per image, meshes of <= 65535 16-bit-indexed
// vertices, bounds per mesh / node) with a single identity node, then runs the reference's own code.  The voxels land in c->map
// (ref_map_list_sectors / ref_map_read_sector), the palette in c->map.Palette (ref_get_material).
REF_API int ref_voxelize(RefCtx* c, uint32_t n_tris, const float* pos9, const float* uv6, const int32_t* tri_tex, uint32_t n_tex, const uint8_t* const* tex_rgba,
                         const uint32_t* tex_w, const uint32_t* tex_h, uint32_t size) {
    glim::Model model("");
    model.Materials.reserve(n_tex);
    for (uint32_t t = 0; t < n_tex; t++) {
        char name[32];
        std::snprintf(name, sizeof(name), "tex%04u", t);
        if (!std::has_single_bit(tex_w[t]) || !std::has_single_bit(tex_h[t])) return -1;
        swr::RgbaTexture2D tex(tex_w[t], tex_h[t], 8, 1);  // Scene.cpp:105-116
        tex.SetPixels(tex_rgba[t], tex_w[t], 0);
        tex.GenerateMips();
        auto slot = model.Textures.insert({name, std::move(tex)});
        model.Materials.push_back(glim::Material{.Texture = &slot.first->second});
    }
    model.VertexBuffer = std::make_unique<glim::Vertex[]>((size_t)n_tris * 3);
    model.IndexBuffer = std::make_unique<glim::VertexIndex[]>((size_t)n_tris * 3);
    glim::ModelNode& root = model.RootNode;
    root.Transform = glm::mat4(1.0f);
    root.Bounds[0] = glm::vec3(FLT_MIN), root.Bounds[1] = glm::vec3(-FLT_MAX);  // Scene.cpp:127 (sic: FLT_MIN)
    const uint32_t kChunk = 65535 / 3;
    for (uint32_t t0 = 0; t0 < n_tris;) {
        uint32_t t1 = t0;
        while (t1 < n_tris && t1 - t0 < kChunk && tri_tex[t1] == tri_tex[t0]) t1++;
        if (tri_tex[t0] < 0 || (uint32_t)tri_tex[t0] >= n_tex) return -2;
        glim::Mesh mesh{.VertexOffset = t0 * 3, .IndexOffset = t0 * 3, .IndexCount = (t1 - t0) * 3, .Material = &model.Materials[(size_t)tri_tex[t0]],
                        .Bounds = {glm::vec3(FLT_MIN), glm::vec3(-FLT_MAX)}};  // Scene.cpp:199
        for (uint32_t t = t0; t < t1; t++)
            for (uint32_t j = 0; j < 3; j++) {
                glim::Vertex& v = model.VertexBuffer[(size_t)t * 3 + j];
                v = {};
                v.x = pos9[(size_t)t * 9 + j * 3], v.y = pos9[(size_t)t * 9 + j * 3 + 1], v.z = pos9[(size_t)t * 9 + j * 3 + 2];
                v.u = uv6[(size_t)t * 6 + j * 2], v.v = uv6[(size_t)t * 6 + j * 2 + 1];
                model.IndexBuffer[(size_t)t * 3 + j] = (glim::VertexIndex)((t - t0) * 3 + j);
                glm::vec3 p(v.x, v.y, v.z);
                mesh.Bounds[0] = glm::min(mesh.Bounds[0], p), mesh.Bounds[1] = glm::max(mesh.Bounds[1], p);  // :216-217
            }
        root.Bounds[0] = glm::min(root.Bounds[0], mesh.Bounds[0]), root.Bounds[1] = glm::max(root.Bounds[1], mesh.Bounds[1]);  // :134-135
        root.Meshes.push_back((uint32_t)model.Meshes.size());
        model.Meshes.push_back(mesh);
        t0 = t1;
    }
    c->map.Sectors.clear();
    c->map.VoxelizeModel(model, glm::uvec3(0), glm::uvec3(size));
    return 0;
}

// Dense view of one sector as the reference's FlatVoxelStorage holds it (CpuRenderer.cpp:20-31).
REF_API int ref_read_sector(RefCtx* c, int sx, int sy, int sz, uint64_t* mask, uint8_t* bricks, uint64_t* cells) {
    glm::ivec3 pos(sx, sy, sz);
    if (!ViewSectorIndexer::CheckInBounds(pos)) return -1;
    uint32_t v = ViewSectorIndexer::GetIndex(pos);
    if (mask) *mask = c->storage->SectorMasks[v];
    if (bricks) std::memcpy(bricks, &c->storage->StorageBuffer[(size_t)v * 64 * 512], 64 * 512);
    if (cells) std::memcpy(cells, &c->storage->OccupancyStorage[(size_t)v * 64 * 8], 64 * 8 * 8);
    return 0;
}

REF_API void ref_set_blue_noise(const uint8_t* rg) {  // same re-tiling as the VBlueNoise ctor, CpuRenderer.cpp:240-251
    uint16_t* ptr = _blueNoise.Data;
    for (uint32_t y = 0; y < 128 * 64; y += simd::TileHeight)
        for (uint32_t x = 0; x < 128; x += simd::TileWidth)
            for (uint32_t i = 0; i < simd::VectorWidth; i++) {
                uint32_t sx = x + (i % simd::TileWidth), sy = y + (i / simd::TileWidth);
                const uint8_t* s = &rg[(sx + sy * 128) * 2];
                *ptr++ = (uint16_t)(s[0] | s[1] << 8);
            }
}
REF_API int ref_set_sky(const VrtSkyDesc* d, const uint32_t* texels) {
    swr::HdrTexture2D t(d->face_size, d->face_size, d->mip_levels, 6);
    if (t.MipLevels != d->mip_levels || t.LayerShift != d->layer_shift) return -1;  // layout disagreement
    std::memcpy(t.Data.get(), texels, (size_t)d->texel_count * 4);
    _skyBox = std::move(t);
    return 0;
}

// RayCast (CpuRenderer.cpp:172-224).  lanes_per_packet = 16: rays are packed 16 to a SIMD packet as
// the renderer does (results then carry the packet-coupled quirks Q2/Q10 of DESIGN.md §3);
// lanes_per_packet = 1: one active lane per packet = the lane-wise semantics the oracle restates.
// out[i].iters is not available from the reference.
REF_API void ref_trace(RefCtx* c, uint64_t n, const float* o3, const float* d3, const int32_t* wo, VrtHit* out, int lanes_per_packet) {
    const uint32_t W = lanes_per_packet == 1 ? 1u : simd::VectorWidth;
    const glm::ivec3 worldOrigin(wo[0], wo[1], wo[2]);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t base = 0; base < (int64_t)n; base += W) {
        VFloat3 o, d;
        VMask mask = 0;
        for (uint32_t l = 0; l < simd::VectorWidth; l++) {
            uint64_t i = (l < W && (uint64_t)base + l < n) ? (uint64_t)base + l : (uint64_t)base;
            o.x[l] = o3[3 * i], o.y[l] = o3[3 * i + 1], o.z[l] = o3[3 * i + 2];
            d.x[l] = d3[3 * i], d.y[l] = d3[3 * i + 1], d.z[l] = d3[3 * i + 2];
            if (l < W && (uint64_t)base + l < n) mask |= (VMask)(1u << l);
        }
        VHitResult h = RayCast(*c->storage, o, d, mask, worldOrigin);
        for (uint32_t l = 0; l < W && (uint64_t)base + l < n; l++) {
            VrtHit& r = out[base + l];
            std::memset(&r, 0, sizeof(r));
            r.material = (uint32_t)h.MaterialData[l];
            r.dist = h.Distance[l];
            r.px = h.Pos.x[l], r.py = h.Pos.y[l], r.pz = h.Pos.z[l];
            r.u = h.UV.x[l], r.v = h.UV.y[l];
            r.vx = wo[0] + (int32_t)std::floor(r.px), r.vy = wo[1] + (int32_t)std::floor(r.py), r.vz = wo[2] + (int32_t)std::floor(r.pz);
            int nx = (int)h.Normal.x[l], ny = (int)h.Normal.y[l], nz = (int)h.Normal.z[l];
            r.flags = (uint32_t)((nx + 1) | ((ny + 1) << 2) | ((nz + 1) << 4)) | (((h.Mask >> l) & 1) ? VRT_HIT_HIT : 0u);
        }
    }
}

// VoxelMap::RayCast (VoxelMap.cpp:140-170)
REF_API void ref_hit_query(RefCtx* c, uint64_t n, const double* o3, const double* d3, uint32_t maxIters, VrtHitD* out) {
    for (uint64_t i = 0; i < n; i++) {  // the hash map is not thread-safe for GetBrick(create=true)
        HitResult h = c->map.RayCast(glm::dvec3(o3[3 * i], o3[3 * i + 1], o3[3 * i + 2]), glm::dvec3(d3[3 * i], d3[3 * i + 1], d3[3 * i + 2]), maxIters);
        VrtHitD& r = out[i];
        std::memset(&r, 0, sizeof(r));
        r.dist = h.Distance;
        if (h.Distance >= 0.0) {
            r.nx = h.Normal.x, r.ny = h.Normal.y, r.nz = h.Normal.z;
            r.u = h.UV.x, r.v = h.UV.y;
            r.vx = h.VoxelPos.x, r.vy = h.VoxelPos.y, r.vz = h.VoxelPos.z;
        }
    }
}

// GetPrimaryRay + OriginFrac for every pixel of a frame (CpuRenderer.cpp:226-233,327-334), row-major.
REF_API void ref_primary_rays(const VrtFrame* f, float* o3, float* d3) {
    glm::mat4 inv;
    std::memcpy(&inv, f->inv_proj, 64);
    for (uint32_t y = 0; y < f->height; y += simd::TileHeight) {
        VFloat v = simd::conv2f((int32_t)y + simd::TileOffsetsY) + 0.5f;
        for (uint32_t x = 0; x < f->width; x += simd::TileWidth) {
            VFloat u = simd::conv2f((int32_t)x + simd::TileOffsetsX) + 0.5f;
            VFloat3 origin, dir;
            GetPrimaryRay({u, v}, inv, origin, dir);
            origin += VFloat3(glm::vec3(f->origin_frac[0], f->origin_frac[1], f->origin_frac[2]));
            for (uint32_t l = 0; l < simd::VectorWidth; l++) {
                size_t p = (size_t)(y + l / 4) * f->width + x + (l & 3);
                o3[3 * p] = origin.x[l], o3[3 * p + 1] = origin.y[l], o3[3 * p + 2] = origin.z[l];
                d3[3 * p] = dir.x[l], d3[3 * p + 1] = dir.y[l], d3[3 * p + 2] = dir.z[l];
            }
        }
    }
}

// The frame loop of CpuRenderer::RenderFrame (CpuRenderer.cpp:444-462) minus GL: FrameConstants +
// RenderRow over tile rows, OpenMP instead of std::execution::par_unseq (no TBB here => serial).
// Returns the seconds spent in the row loop (what _frameTime brackets, :442,464).
REF_API double ref_render(RefCtx* c, const VrtFrame* f, void* out_tiles, int threads) {
    glm::mat4 proj, inv;
    std::memcpy(&proj, f->proj, 64);
    std::memcpy(&inv, f->inv_proj, 64);
    FrameConstants fc = {
        .Storage = *c->storage,
        .Size = glm::uvec2(f->width, f->height),
        .WorldOrigin = glm::ivec3(f->world_origin[0], f->world_origin[1], f->world_origin[2]),
        .OriginFrac = glm::vec3(f->origin_frac[0], f->origin_frac[1], f->origin_frac[2]),
        .FrameNo = f->frame_no,
        .NumLightBounces = f->bounces,
        .CurrentProj = proj,
        .InvProj = inv,
    };
    auto* tiles = (Framebuffer::Tile*)out_tiles;
    const uint32_t stride = f->width / simd::TileWidth;
    const int rows = (int)(f->height / simd::TileHeight);
    if (threads <= 0) threads = omp_get_max_threads();
    double t0 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int row = 0; row < rows; row++) RenderRow(fc, &tiles[(size_t)row * stride], (uint32_t)row * simd::TileHeight);
    return omp_get_wtime() - t0;
}

// GBuffer::GetInverseProjScreenMat (GBuffer.h:133-139) through the stand-in glm
REF_API void ref_inverse_proj_screen(const float* m16, int w, int h, float* out16) {
    glm::mat4 m;
    std::memcpy(&m, m16, 64);
    glm::mat4 r = GBuffer::GetInverseProjScreenMat(m, glm::ivec2(w, h));
    std::memcpy(out16, &r, 64);
}

// ---- small probes of the shading helpers (lane 0 of a packet) ---------------------------------------
// simd::sincos_2pi (SIMD.h:175-190)
REF_API void ref_sincos_2pi(float x, float* s, float* c) {
    VFloat vs, vc;
    simd::sincos_2pi(VFloat(x), vs, vc);
    *s = vs[0], *c = vc[0];
}
// SampleDirection (CpuRenderer.cpp:273-291) — uses rsqrt14, so only approximately the oracle's value
REF_API void ref_sample_direction(float sx, float sy, float* out3) {
    VFloat3 d = SampleDirection({VFloat(sx), VFloat(sy)});
    out3[0] = d.x[0], out3[1] = d.y[0], out3[2] = d.z[0];
}
// VBlueNoise::Sample (CpuRenderer.cpp:254-270) for the 4x4 tile at (x & ~3, y & ~3): 16 (R,G)/255 pairs
REF_API void ref_blue_noise_tile(uint32_t x, uint32_t y, uint32_t frame_no, uint32_t sample_idx, float* out32) {
    VFloat2 v = _blueNoise.Sample(glm::uvec2(x & ~3u, y & ~3u), frame_no, sample_idx);
    for (int l = 0; l < 16; l++) out32[2 * l] = v.x[l], out32[2 * l + 1] = v.y[l];
}
// sky: _skyBox.SampleCube<Nearest, mips>(dir, mip) * 3 (CpuRenderer.cpp:350-356) — ProjectCubemap uses rcp14
REF_API void ref_sky_sample(const float* dir3, int mip, float* out3) {
    constexpr swr::SamplerDesc SD = {.MagFilter = swr::FilterMode::Nearest, .MinFilter = swr::FilterMode::Nearest, .EnableMips = true};
    VFloat3 c = _skyBox.SampleCube<SD, false>(VFloat3(VFloat(dir3[0]), VFloat(dir3[1]), VFloat(dir3[2])), (float)mip);
    out3[0] = c.x[0] * 3.0f, out3[1] = c.y[0] * 3.0f, out3[2] = c.z[0] * 3.0f;
}
// swr::pixfmt::R11G11B10f::Pack / RGBA8u::Pack / RG16f::Pack (Texture.h:41-62,101-116,158-177), lane 0
REF_API uint32_t ref_pack_r11g11b10f(float r, float g, float b) { return (uint32_t)swr::pixfmt::R11G11B10f::Pack({VFloat(r), VFloat(g), VFloat(b)})[0]; }
REF_API uint32_t ref_pack_rgba8(float r, float g, float b, float a) { return (uint32_t)swr::pixfmt::RGBA8u::Pack({VFloat(r), VFloat(g), VFloat(b), VFloat(a)})[0]; }
REF_API uint32_t ref_pack_rg16f(float x, float y) { return (uint32_t)swr::pixfmt::RG16f::Pack({VFloat(x), VFloat(y)})[0]; }
