// Stand-in for LibGlimpsw/OGL/QuickGL.h: just enough surface for Renderer.h / GBuffer.h /
// CpuRenderer.cpp to compile.  A Buffer is host memory, shaders and textures do nothing.  OUR code.
#pragma once
#include <glad/glad.h>
#include <glm/glm.hpp>

#include <cstdint>
#include <cstdlib>
#include <memory>
#include <string_view>

namespace ogl {
struct Texture2D {
    uint32_t Width, Height;
    Texture2D(uint32_t w, uint32_t h, uint32_t, GLenum) : Width(w), Height(h) {}
};
struct TextureCube {};
struct Buffer {
    void* Data;
    size_t Size;
    Buffer(size_t size, GLenum) : Data(std::calloc(1, size)), Size(size) {}
    ~Buffer() { std::free(Data); }
    template <typename T>
    T* Map(GLenum) { return (T*)Data; }
};
struct Shader {
    template <typename T>
    void SetUniform(std::string_view, const T&) {}
    void DispatchCompute(uint32_t, uint32_t, uint32_t) {}
    void DispatchFullscreen() {}
};
}  // namespace ogl
