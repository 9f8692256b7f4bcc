#include "b200_renderer.h"

#include <chrono>
#include <cstring>

namespace vrt_host {

// ---- matrices (float, column-major; formulas of the GLM functions the reference calls) ------------
mat4 operator*(const mat4& a, const mat4& b) {
    mat4 r;
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++) {
            float acc = 0.0f;
            for (int k = 0; k < 4; k++) acc += a.at(k, row) * b.at(c, k);
            r.at(c, row) = acc;
        }
    return r;
}
// glm::inverse(mat4) — what GBuffer::GetInverseProjScreenMat calls (GBuffer.h:134): GLM's float cofactor expansion
// (18 2x2 sub-determinants, one reciprocal of the determinant).  One rounding per operation (this file is built
// -ffp-contract=off); tests/test_native_host.py pins the result bit for bit against the reference's own
// GetInverseProjScreenMat compiled into oracle/_ref.
mat4 Inverse(const mat4& a) {
    auto m = [&](int c, int r) { return a.at(c, r); };
    const float c00 = m(2, 2) * m(3, 3) - m(3, 2) * m(2, 3), c02 = m(1, 2) * m(3, 3) - m(3, 2) * m(1, 3), c03 = m(1, 2) * m(2, 3) - m(2, 2) * m(1, 3);
    const float c04 = m(2, 1) * m(3, 3) - m(3, 1) * m(2, 3), c06 = m(1, 1) * m(3, 3) - m(3, 1) * m(1, 3), c07 = m(1, 1) * m(2, 3) - m(2, 1) * m(1, 3);
    const float c08 = m(2, 1) * m(3, 2) - m(3, 1) * m(2, 2), c10 = m(1, 1) * m(3, 2) - m(3, 1) * m(1, 2), c11 = m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2);
    const float c12 = m(2, 0) * m(3, 3) - m(3, 0) * m(2, 3), c14 = m(1, 0) * m(3, 3) - m(3, 0) * m(1, 3), c15 = m(1, 0) * m(2, 3) - m(2, 0) * m(1, 3);
    const float c16 = m(2, 0) * m(3, 2) - m(3, 0) * m(2, 2), c18 = m(1, 0) * m(3, 2) - m(3, 0) * m(1, 2), c19 = m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2);
    const float c20 = m(2, 0) * m(3, 1) - m(3, 0) * m(2, 1), c22 = m(1, 0) * m(3, 1) - m(3, 0) * m(1, 1), c23 = m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1);
    const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    const float v0[4] = {m(1, 0), m(0, 0), m(0, 0), m(0, 0)}, v1[4] = {m(1, 1), m(0, 1), m(0, 1), m(0, 1)};
    const float v2[4] = {m(1, 2), m(0, 2), m(0, 2), m(0, 2)}, v3[4] = {m(1, 3), m(0, 3), m(0, 3), m(0, 3)};
    mat4 inv;
    for (int i = 0; i < 4; i++) {
        const float sa = (i & 1) ? -1.0f : 1.0f, sb = -sa;
        inv.at(0, i) = ((v1[i] * f0[i] - v2[i] * f1[i]) + v3[i] * f2[i]) * sa;
        inv.at(1, i) = ((v0[i] * f0[i] - v2[i] * f3[i]) + v3[i] * f4[i]) * sb;
        inv.at(2, i) = ((v0[i] * f1[i] - v1[i] * f3[i]) + v3[i] * f5[i]) * sa;
        inv.at(3, i) = ((v0[i] * f2[i] - v1[i] * f4[i]) + v2[i] * f5[i]) * sb;
    }
    const float d0 = m(0, 0) * inv.at(0, 0), d1 = m(0, 1) * inv.at(1, 0), d2 = m(0, 2) * inv.at(2, 0), d3 = m(0, 3) * inv.at(3, 0);
    const float one_over_det = 1.0f / ((d0 + d1) + (d2 + d3));
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) inv.at(c, r) = inv.at(c, r) * one_over_det;
    return inv;
}
mat4 Translate(const mat4& a, float x, float y, float z) {
    mat4 r = a;
    for (int row = 0; row < 4; row++) r.at(3, row) = ((a.at(0, row) * x + a.at(1, row) * y) + a.at(2, row) * z) + a.at(3, row);
    return r;
}
mat4 Scale(const mat4& a, float x, float y, float z) {
    mat4 r = a;
    for (int row = 0; row < 4; row++) {
        r.at(0, row) = a.at(0, row) * x;
        r.at(1, row) = a.at(1, row) * y;
        r.at(2, row) = a.at(2, row) * z;
    }
    return r;
}
mat4 Perspective(float fovy, float aspect, float zNear, float zFar) {
    float t = std::tan(fovy / 2.0f);
    mat4 r;
    std::memset(r.m, 0, sizeof(r.m));
    r.at(0, 0) = 1.0f / (aspect * t);
    r.at(1, 1) = 1.0f / t;
    r.at(2, 2) = -(zFar + zNear) / (zFar - zNear);
    r.at(2, 3) = -1.0f;
    r.at(3, 2) = -(2.0f * zFar * zNear) / (zFar - zNear);
    return r;
}
mat4 EulerAngleXY(float ax, float ay) {
    float cx = std::cos(ax), sx = std::sin(ax), cy = std::cos(ay), sy = std::sin(ay);
    mat4 r;
    const float v[16] = {cy, -sx * -sy, cx * -sy, 0, 0, cx, sx, 0, sy, -sx * cy, cx * cy, 0, 0, 0, 0, 1};
    std::memcpy(r.m, v, sizeof(v));
    return r;
}
mat4 GetInverseProjScreenMat(const mat4& projView, uint32_t w, uint32_t h) {
    mat4 inv = Inverse(projView);
    inv = Translate(inv, -1.0f, -1.0f, 0.0f);
    inv = Scale(inv, 2.0f / (float)w, 2.0f / (float)h, 1.0f);
    return Translate(inv, 0.5f, 0.5f, 0.0f);  // offset to pixel centre
}
mat4 Camera::GetViewMatrix(bool translateToView) const {
    mat4 m = EulerAngleXY(-Euler[1], Euler[0]);  // Camera.h:42 destRotation = eulerAngleXY(-Euler.y, Euler.x)
    if (translateToView) m = Translate(m, (float)-ViewPosition.x, (float)-ViewPosition.y, (float)-ViewPosition.z);
    return m;
}
mat4 Camera::GetProjMatrix() const { return Perspective(FieldOfView * 0.01745329251994329577f, AspectRatio, NearZ, FarZ); }

// ---- renderer -----------------------------------------------------------------------------------------
void B200Renderer::Check(int status, const char* what) {
    if (status != VRT_OK) throw std::runtime_error(std::string(what) + ": " + vrt_last_error(_ctx));
}

B200Renderer::B200Renderer(std::shared_ptr<VoxelMap> map, int device, uint32_t xz, uint32_t y) : _map(std::move(map)) {
    VrtConfig cfg{};
    cfg.struct_size = sizeof(cfg);
    cfg.device = device;
    cfg.sectors_xz_log2 = xz;
    cfg.sectors_y_log2 = y;
    _device = device;
    int st = vrt_create(&cfg, &_ctx);
    if (st != VRT_OK) throw std::runtime_error(std::string("vrt_create: ") + vrt_last_error(nullptr));
    _map->MarkAllDirty();  // CpuRenderer.cpp:410
}
B200Renderer::~B200Renderer() {
    vrt_gbuffer_destroy(_gbuffer);
    vrt_destroy(_ctx);
}

void B200Renderer::SyncBuffers(VoxelMap& map) {
    // palette: the reference re-encodes it every frame (CpuRenderer.cpp:34-36); upload only on change
    uint64_t enc[256];
    for (uint32_t i = 0; i < 256; i++) enc[i] = map.Palette[i].GetEncoded();
    if (!_paletteValid || std::memcmp(enc, _paletteEncoded, sizeof(enc)) != 0) {
        Check(vrt_set_palette(_ctx, enc), "vrt_set_palette");
        std::memcpy(_paletteEncoded, enc, sizeof(enc));
        _paletteValid = true;
    }
    if (map.DirtyLocs.empty()) return;

    size_t bricks = 0;
    for (auto& [idx, dirty] : map.DirtyLocs) {
        auto it = map.Sectors.find(idx);
        if (it != map.Sectors.end()) bricks += (size_t)__builtin_popcountll(dirty & it->second.GetAllocationMask());
    }
    _payload.resize(bricks * sizeof(Brick));
    std::vector<VrtDirtySector> recs;
    recs.reserve(map.DirtyLocs.size());
    size_t off = 0;
    for (auto& [idx, dirty] : map.DirtyLocs) {
        ivec3 pos = WorldSectorIndexer::GetPos(idx);
        VrtDirtySector r{};
        r.sx = pos.x, r.sy = pos.y, r.sz = pos.z;
        auto it = map.Sectors.find(idx);
        if (it == map.Sectors.end()) {  // sector deleted (CpuRenderer.cpp:43-46)
            r.flags = VRT_SECTOR_REMOVED;
            r.dirty_mask = ~0ull;
        } else {
            Sector& s = it->second;
            r.alloc_mask = s.GetAllocationMask();
            r.dirty_mask = dirty;
            r.bricks = _payload.data() + off;
            for (uint64_t m = dirty & r.alloc_mask; m; m &= m - 1) {
                std::memcpy(_payload.data() + off, s.GetBrick((uint32_t)__builtin_ctzll(m)), sizeof(Brick));
                off += sizeof(Brick);
            }
        }
        recs.push_back(r);
    }
    Check(vrt_sync(_ctx, (uint32_t)recs.size(), recs.data()), "vrt_sync");
    map.DirtyLocs.clear();  // CpuRenderer.cpp:60
}

void B200Renderer::RenderFrame(Camera& cam, uvec2 viewSize) {
    viewSize.x &= ~3u;  // round down to 4x4 steps (CpuRenderer.cpp:419)
    viewSize.y &= ~3u;
    const bool worldChanged = !_map->DirtyLocs.empty();  // CpuRenderer.cpp:421
    SyncBuffers(*_map);
    FrameNo++;  // GBuffer::SetCamera (GBuffer.h:57)

    mat4 proj = cam.GetProjMatrix() * cam.GetViewMatrix(false);  // GBuffer.h:51
    mat4 inv = GetInverseProjScreenMat(proj, viewSize.x, viewSize.y);
    VrtFrame f{};
    f.width = viewSize.x, f.height = viewSize.y;
    std::memcpy(f.inv_proj, inv.m, sizeof(inv.m));
    std::memcpy(f.proj, proj.m, sizeof(proj.m));
    const double p[3] = {cam.ViewPosition.x, cam.ViewPosition.y, cam.ViewPosition.z};
    for (int a = 0; a < 3; a++) {
        double fl = std::floor(p[a]);
        f.world_origin[a] = (int32_t)fl;       // CpuRenderer.cpp:447
        f.origin_frac[a] = (float)(p[a] - fl);  // :448
    }
    f.frame_no = FrameNo;
    f.bounces = NumLightBounces;
    f.part_count = 1;
    _tiles.resize((size_t)viewSize.x * viewSize.y / 16);
    _size = viewSize;
    auto t0 = std::chrono::steady_clock::now();
    if (DenoiseAndPresent) {
        if (!_gbuffer && vrt_gbuffer_create(_device, &_gbuffer) != VRT_OK)
            throw std::runtime_error(std::string("vrt_gbuffer_create: ") + vrt_gbuffer_last_error(nullptr));
        auto gcheck = [&](int st, const char* what) {
            if (st != VRT_OK) throw std::runtime_error(std::string(what) + ": " + vrt_gbuffer_last_error(_gbuffer));
        };
        VrtGBufferCamera gc{};
        gc.width = viewSize.x, gc.height = viewSize.y;
        std::memcpy(gc.proj, proj.m, sizeof(proj.m));
        std::memcpy(gc.inv_proj, inv.m, sizeof(inv.m));
        for (int a = 0; a < 3; a++) gc.position[a] = p[a];
        gc.reset_history = worldChanged ? 1u : 0u;  // GBuffer::SetCamera(cam, viewSize, worldChanged), CpuRenderer.cpp:423
        gcheck(vrt_gbuffer_set_passes(_gbuffer, NumDenoiserPasses), "vrt_gbuffer_set_passes");
        gcheck(vrt_gbuffer_set_debug_channel(_gbuffer, DebugChannelView), "vrt_gbuffer_set_debug_channel");
        gcheck(vrt_gbuffer_set_camera(_gbuffer, &gc), "vrt_gbuffer_set_camera");
        _rgba.resize((size_t)viewSize.x * viewSize.y);
        gcheck(vrt_gbuffer_render_present(_gbuffer, _ctx, &f, _rgba.data()), "vrt_gbuffer_render_present");
    } else {
        Check(vrt_render(_ctx, &f, _tiles.data(), nullptr), "vrt_render");
    }
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    _raysPerSec = (double)viewSize.x * viewSize.y * (NumLightBounces + 1) * (1000.0 / ms);  // CpuRenderer.cpp:490-492
}

std::vector<HitResult> B200Renderer::RayCast(const std::vector<dvec3>& origins, const std::vector<dvec3>& dirs, uint32_t maxIters) {
    if (origins.size() != dirs.size()) throw std::runtime_error("RayCast: origins/dirs size mismatch");
    SyncBuffers(*_map);
    std::vector<VrtHitD> raw(origins.size());
    static_assert(sizeof(dvec3) == 24, "dvec3 layout");
    Check(vrt_hit_query(_ctx, origins.size(), &origins[0].x, &dirs[0].x, maxIters, raw.data()), "vrt_hit_query");
    std::vector<HitResult> out(raw.size());
    for (size_t i = 0; i < raw.size(); i++) {
        out[i].Distance = raw[i].dist;
        out[i].Normal[0] = raw[i].nx, out[i].Normal[1] = raw[i].ny, out[i].Normal[2] = raw[i].nz;
        out[i].UV[0] = raw[i].u, out[i].UV[1] = raw[i].v;
        out[i].VoxelPos = {raw[i].vx, raw[i].vy, raw[i].vz};
    }
    return out;
}

void B200Renderer::SetBlueNoise(const uint8_t* rg) { Check(vrt_set_blue_noise(_ctx, rg, VRT_BLUE_NOISE_BYTES), "vrt_set_blue_noise"); }
void B200Renderer::SetSky(const VrtSkyDesc& desc, const uint32_t* texels) { Check(vrt_set_sky(_ctx, &desc, texels), "vrt_set_sky"); }

}  // namespace vrt_host

// Test hook (tests/test_native_host.py): GetInverseProjScreenMat for a column-major float[16], so that the adapter's matrix can be
// compared bit for bit with the reference's own GBuffer::GetInverseProjScreenMat (oracle/_ref) from Python.
extern "C" __attribute__((visibility("default"))) void vrt_host_inverse_proj_screen(const float* proj_view16, uint32_t w, uint32_t h, float* out16) {
    vrt_host::mat4 m;
    std::memcpy(m.m, proj_view16, sizeof(m.m));
    vrt_host::mat4 r = vrt_host::GetInverseProjScreenMat(m, w, h);
    std::memcpy(out16, r.m, sizeof(r.m));
}

