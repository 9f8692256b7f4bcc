// Stand-in for LibGlimpsw/Common/Camera.h without the ImGui input handling: the members and the two
// matrix getters GBuffer::SetCamera uses (Camera.h:17-34 semantics: perspective(radians(FOV), aspect,
// near, far); view = rotation only when translateToView is false).  OUR code.
#pragma once
#include <glm/glm.hpp>

namespace glim {
struct Camera {
    glm::dvec3 Position{}, ViewPosition{};
    glm::vec2 Euler{};  // yaw, pitch
    float FieldOfView = 90.0f, AspectRatio = 1.0f, MoveSpeed = 10.0f, NearZ = 0.01f, FarZ = 1000.0f;
    glm::mat4 GetViewMatrix(bool translateToView = true) {
        glm::mat4 m = glm::eulerAngleXY(-Euler.y, Euler.x);
        if (translateToView) m = glm::translate(m, glm::vec3(-ViewPosition));
        return m;
    }
    glm::mat4 GetProjMatrix() { return glm::perspective(glm::radians(FieldOfView), AspectRatio, NearZ, FarZ); }
};
}  // namespace glim
