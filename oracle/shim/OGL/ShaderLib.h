#pragma once
#include "QuickGL.h"
namespace ogl {
struct ShaderLib {
    std::shared_ptr<Shader> LoadComp(std::string_view) { return std::make_shared<Shader>(); }
    std::shared_ptr<Shader> LoadFrag(std::string_view) { return std::make_shared<Shader>(); }
};
}  // namespace ogl
