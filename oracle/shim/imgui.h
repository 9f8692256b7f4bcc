// Stub of the few Dear ImGui entry points the reference's renderer TU names (DrawSettings only).
// OUR code; the UI is out of scope.  Test infrastructure for oracle/_ref.
#pragma once
namespace ImGui {
inline void SeparatorText(const char*) {}
inline void PushItemWidth(float) {}
inline void PopItemWidth() {}
inline void Separator() {}
inline void Text(const char*, ...) {}
}  // namespace ImGui
