// Stand-in for LibGlimpsw/Common/SettingStore.h (ImGui-bound settings + TimeStat): no-ops.  OUR code.
#pragma once
#include <imgui.h>

#include <cstdint>
#include <string_view>

namespace glim {
struct SettingStore {
    template <typename... A>
    bool Combo(A&&...) { return false; }
    template <typename... A>
    bool Slider(A&&...) { return false; }
};
struct TimeStat {
    void Begin() {}
    void End() {}
    void GetElapsedMs(double& mean, double& dev) const { mean = 1.0, dev = 0.0; }
    void Draw(std::string_view) const {}
};
}  // namespace glim
