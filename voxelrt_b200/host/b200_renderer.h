// B200Renderer — the drop-in for the reference's `Renderer` implementations (src/VoxelRT/Renderer.h:15-66):
// same virtuals (`RenderFrame(Camera&, viewSize)`), same pull contract with the voxel map (drains and
// clears `VoxelMap::DirtyLocs` inside SyncBuffers like CpuRenderer.cpp:33-61 / GpuRenderer.cpp:45-79),
// same frame constants (GBuffer::SetCamera, GBuffer.h:31-60; CpuRenderer.cpp:444-453).  All device work
// goes through the C ABI of libvoxelrt_b200.so (include/voxelrt_b200.h); there is no CPU fallback —
// the constructor throws std::runtime_error when the CUDA path is unavailable.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/voxelrt_b200_post.h"
#include "voxel_map.h"

namespace vrt_host {

struct uvec2 {
    uint32_t x = 0, y = 0;
};

// Column-major 4x4 float matrix, m[col*4+row] (what glm::mat4 is in memory).
struct mat4 {
    float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    float& at(int col, int row) { return m[col * 4 + row]; }
    float at(int col, int row) const { return m[col * 4 + row]; }
};
mat4 operator*(const mat4& a, const mat4& b);
mat4 Inverse(const mat4& a);
mat4 Translate(const mat4& a, float x, float y, float z);
mat4 Scale(const mat4& a, float x, float y, float z);
mat4 Perspective(float fovyRad, float aspect, float zNear, float zFar);
mat4 EulerAngleXY(float angleX, float angleY);
// GBuffer::GetInverseProjScreenMat (GBuffer.h:133-139)
mat4 GetInverseProjScreenMat(const mat4& projView, uint32_t width, uint32_t height);

// glim::Camera as the renderers see it (LibGlimpsw/Common/Camera.h:8-34)
struct Camera {
    dvec3 ViewPosition;
    float Euler[2] = {0, 0};  // yaw, pitch
    float FieldOfView = 90.0f, AspectRatio = 1.0f, NearZ = 0.01f, FarZ = 1000.0f;
    mat4 GetViewMatrix(bool translateToView = true) const;
    mat4 GetProjMatrix() const;
};

struct Renderer {
    virtual ~Renderer() {}
    virtual void RenderFrame(Camera& cam, uvec2 viewSize) = 0;
};

class B200Renderer : public Renderer {
public:
    // viewSectorsXZLog2 / YLog2: extent of the resident view (6/4 = the reference CPU renderer's
    // 2048x512x2048 voxels, CpuRenderer.cpp:14-16; 7/6 = the GPU renderer's, GpuRenderer.cpp:6-8)
    B200Renderer(std::shared_ptr<VoxelMap> map, int device = -1, uint32_t viewSectorsXZLog2 = 6, uint32_t viewSectorsYLog2 = 4);
    ~B200Renderer() override;

    // CpuRenderer.cpp:415-474.  With DenoiseAndPresent == false (default) it stops before the blit and leaves the 16 B/px
    // tiles in Tiles(); with true the frame stays on the device through GBuffer::DenoiseAndPresent (GBuffer.h:86-130) and
    // only the presented RGBA8 image comes back (Presented()).
    void RenderFrame(Camera& cam, uvec2 viewSize) override;
    void SyncBuffers(VoxelMap& map);                         // CpuRenderer.cpp:33-61
    // VoxelMap::RayCast (VoxelMap.cpp:140-170) on the device, batched
    std::vector<HitResult> RayCast(const std::vector<dvec3>& origins, const std::vector<dvec3>& dirs, uint32_t maxIters = 1024);

    void SetBlueNoise(const uint8_t* rg128x8192);                  // VBlueNoise source texels (R,G)
    void SetSky(const VrtSkyDesc& desc, const uint32_t* texels);   // swr::HdrTexture2D cube

    uint32_t NumLightBounces = 1;  // Renderer.h:65
    bool DenoiseAndPresent = false;
    uint32_t NumDenoiserPasses = 5;  // GBuffer.h:23
    uint32_t DebugChannelView = 0;   // GBuffer::DebugChannel (GBuffer.h:8,22)
    const std::vector<uint32_t>& Presented() const { return _rgba; }  // w*h RGBA8, what GBufferBlit.frag draws
    uint32_t FrameNo = 0;          // GBuffer::FrameNo
    // The last frame, 16 B/px in Framebuffer::Tile order (CpuRenderer.cpp:299-309); host copy for a presenter.
    const std::vector<VrtTile>& Tiles() const { return _tiles; }
    uvec2 FrameSize() const { return _size; }
    double RaysPerSecondOfLastFrame() const { return _raysPerSec; }  // CpuRenderer.cpp:490-494
    VrtContext* Handle() const { return _ctx; }

private:
    void Check(int status, const char* what);
    std::shared_ptr<VoxelMap> _map;
    VrtContext* _ctx = nullptr;
    VrtGBuffer* _gbuffer = nullptr;  // created on first use
    std::vector<uint32_t> _rgba;
    int _device = -1;
    std::vector<VrtTile> _tiles;
    std::vector<uint8_t> _payload;
    uint64_t _paletteEncoded[256] = {};
    bool _paletteValid = false;
    uvec2 _size;
    double _raysPerSec = 0;
};

}  // namespace vrt_host
