// Minimal stand-in for the parts of GLM that the reference's CPU renderer / voxel map touch.
// GLM itself is a vcpkg dependency that is absent from this image; this header is
// OUR code (written from the GLSL/GLM public semantics), used only to compile the reference's own
// CpuRenderer.cpp / VoxelMap.cpp from where they lie for oracle/_ref.  Test infrastructure.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>

namespace glm {

template <int N, typename T>
struct vec;

template <typename T>
struct vec<2, T> {
    T x, y;
    constexpr vec() : x(0), y(0) {}
    constexpr vec(T s) : x(s), y(s) {}
    template <typename A, typename B>
    constexpr vec(A a, B b) : x((T)a), y((T)b) {}
    template <typename U>
    constexpr vec(const vec<2, U>& v) : x((T)v.x), y((T)v.y) {}
    constexpr T& operator[](int i) { return i == 0 ? x : y; }
    constexpr const T& operator[](int i) const { return i == 0 ? x : y; }
};
template <typename T>
struct vec<3, T> {
    T x, y, z;
    constexpr vec() : x(0), y(0), z(0) {}
    constexpr vec(T s) : x(s), y(s), z(s) {}
    template <typename A, typename B, typename C>
    constexpr vec(A a, B b, C c) : x((T)a), y((T)b), z((T)c) {}
    template <typename U>
    constexpr vec(const vec<3, U>& v) : x((T)v.x), y((T)v.y), z((T)v.z) {}
    template <typename U>
    constexpr vec(const vec<4, U>& v);  // (GLM's conversion constructors are implicit unless GLM_FORCE_EXPLICIT_CTOR: Voxelize.cpp:104,126 rely on it)
    constexpr T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    constexpr const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <typename T>
struct vec<4, T> {
    T x, y, z, w;
    constexpr vec() : x(0), y(0), z(0), w(0) {}
    constexpr vec(T s) : x(s), y(s), z(s), w(s) {}
    template <typename A, typename B, typename C, typename D>
    constexpr vec(A a, B b, C c, D d) : x((T)a), y((T)b), z((T)c), w((T)d) {}
    template <typename U, typename D>
    constexpr vec(const vec<3, U>& v, D d) : x((T)v.x), y((T)v.y), z((T)v.z), w((T)d) {}
    template <typename U>
    constexpr vec(const vec<4, U>& v) : x((T)v.x), y((T)v.y), z((T)v.z), w((T)v.w) {}
    constexpr T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    constexpr const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};
template <typename T>
template <typename U>
constexpr vec<3, T>::vec(const vec<4, U>& v) : x((T)v.x), y((T)v.y), z((T)v.z) {}

using vec2 = vec<2, float>;
using vec3 = vec<3, float>;
using vec4 = vec<4, float>;
using dvec2 = vec<2, double>;
using dvec3 = vec<3, double>;
using dvec4 = vec<4, double>;
using ivec2 = vec<2, int32_t>;
using ivec3 = vec<3, int32_t>;
using ivec4 = vec<4, int32_t>;
using uvec2 = vec<2, uint32_t>;
using uvec3 = vec<3, uint32_t>;
using uvec4 = vec<4, uint32_t>;
using uint32 = uint32_t;
using int32 = int32_t;
using uint = unsigned int;
using bvec2 = vec<2, bool>;
using bvec3 = vec<3, bool>;

// component-wise application helpers
template <int N, typename T, typename F>
constexpr vec<N, T> map1(const vec<N, T>& a, F f) {
    vec<N, T> r;
    for (int i = 0; i < N; i++) r[i] = f(a[i]);
    return r;
}
template <int N, typename T, typename F>
constexpr vec<N, T> map2(const vec<N, T>& a, const vec<N, T>& b, F f) {
    vec<N, T> r;
    for (int i = 0; i < N; i++) r[i] = f(a[i], b[i]);
    return r;
}

#define GLMS_BINOP(op)                                                                                                  \
    template <int N, typename T>                                                                                        \
    constexpr vec<N, T> operator op(const vec<N, T>& a, const vec<N, T>& b) {                                           \
        vec<N, T> r;                                                                                                    \
        for (int i = 0; i < N; i++) r[i] = (T)(a[i] op b[i]);                                                           \
        return r;                                                                                                       \
    }                                                                                                                   \
    template <int N, typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>                      \
    constexpr vec<N, T> operator op(const vec<N, T>& a, S s) {                                                          \
        vec<N, T> r;                                                                                                    \
        for (int i = 0; i < N; i++) r[i] = (T)(a[i] op (T)s);                                                           \
        return r;                                                                                                       \
    }                                                                                                                   \
    template <int N, typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>                      \
    constexpr vec<N, T> operator op(S s, const vec<N, T>& a) {                                                          \
        vec<N, T> r;                                                                                                    \
        for (int i = 0; i < N; i++) r[i] = (T)((T)s op a[i]);                                                           \
        return r;                                                                                                       \
    }                                                                                                                   \
    template <int N, typename T>                                                                                        \
    constexpr vec<N, T>& operator op##=(vec<N, T>& a, const vec<N, T>& b) {                                             \
        for (int i = 0; i < N; i++) a[i] = (T)(a[i] op b[i]);                                                           \
        return a;                                                                                                       \
    }                                                                                                                   \
    template <int N, typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>                      \
    constexpr vec<N, T>& operator op##=(vec<N, T>& a, S s) {                                                            \
        for (int i = 0; i < N; i++) a[i] = (T)(a[i] op (T)s);                                                           \
        return a;                                                                                                       \
    }
GLMS_BINOP(+)
GLMS_BINOP(-)
GLMS_BINOP(*)
GLMS_BINOP(/)
GLMS_BINOP(&)
GLMS_BINOP(|)
GLMS_BINOP(^)
GLMS_BINOP(>>)
GLMS_BINOP(<<)
#undef GLMS_BINOP

template <int N, typename T>
constexpr vec<N, T> operator-(const vec<N, T>& a) {
    vec<N, T> r;
    for (int i = 0; i < N; i++) r[i] = -a[i];
    return r;
}
template <int N, typename T>
constexpr vec<N, T> operator~(const vec<N, T>& a) {
    vec<N, T> r;
    for (int i = 0; i < N; i++) r[i] = ~a[i];
    return r;
}
template <int N, typename T>
constexpr bool operator==(const vec<N, T>& a, const vec<N, T>& b) {
    for (int i = 0; i < N; i++)
        if (!(a[i] == b[i])) return false;
    return true;
}
template <int N, typename T>
constexpr bool operator!=(const vec<N, T>& a, const vec<N, T>& b) {
    return !(a == b);
}

// ---- common functions (GLSL semantics) ----
template <typename T, typename = std::enable_if_t<std::is_floating_point_v<T>>>
inline T floor(T x) { return std::floor(x); }
template <typename T, typename = std::enable_if_t<std::is_floating_point_v<T>>>
inline T fract(T x) { return x - std::floor(x); }
template <typename T, typename = std::enable_if_t<std::is_floating_point_v<T>>>
inline T round(T x) { return std::round(x); }
template <typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>>
constexpr T min(T a, T b) { return (b < a) ? b : a; }
template <typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>>
constexpr T max(T a, T b) { return (a < b) ? b : a; }
template <typename T, typename = std::enable_if_t<std::is_arithmetic_v<T>>>
constexpr T clamp(T x, T lo, T hi) { return min(max(x, lo), hi); }
template <typename T, typename = std::enable_if_t<std::is_floating_point_v<T>>>
inline T sign(T x) { return (T)((T(0) < x) - (x < T(0))); }
template <typename T, typename = std::enable_if_t<std::is_floating_point_v<T>>>
inline T step(T edge, T x) { return x < edge ? T(0) : T(1); }
template <typename T, typename = std::enable_if_t<std::is_floating_point_v<T>>>
inline T mix(T a, T b, T t) { return a * (T(1) - t) + b * t; }
template <typename T, typename = std::enable_if_t<std::is_floating_point_v<T>>>
inline T mix(T a, T b, bool t) { return t ? b : a; }
template <typename T>
constexpr T radians(T deg) { return deg * T(0.01745329251994329576923690768489); }
template <typename T>
constexpr T pi() { return T(3.14159265358979323846264338327950288); }
template <typename T>
constexpr T two_pi() { return T(6.28318530717958647692528676655900576); }

template <int N, typename T>
inline vec<N, T> floor(const vec<N, T>& a) { return map1(a, [](T v) { return (T)std::floor(v); }); }
template <int N, typename T>
inline vec<N, T> fract(const vec<N, T>& a) { return map1(a, [](T v) { return (T)(v - std::floor(v)); }); }
template <int N, typename T>
inline vec<N, T> round(const vec<N, T>& a) { return map1(a, [](T v) { return (T)std::round(v); }); }
template <int N, typename T>
inline vec<N, T> sign(const vec<N, T>& a) { return map1(a, [](T v) { return sign(v); }); }
template <int N, typename T>
constexpr vec<N, T> min(const vec<N, T>& a, const vec<N, T>& b) { return map2(a, b, [](T p, T q) { return min(p, q); }); }
template <int N, typename T>
constexpr vec<N, T> max(const vec<N, T>& a, const vec<N, T>& b) { return map2(a, b, [](T p, T q) { return max(p, q); }); }
template <int N, typename T>
constexpr vec<N, T> clamp(const vec<N, T>& a, T lo, T hi) { return map1(a, [lo, hi](T v) { return clamp(v, lo, hi); }); }
template <int N, typename T>
inline vec<N, T> step(T edge, const vec<N, T>& a) { return map1(a, [edge](T v) { return step(edge, v); }); }
template <int N, typename T>
inline vec<N, bool> greaterThanEqual(const vec<N, T>& a, const vec<N, T>& b) {
    vec<N, bool> r;
    for (int i = 0; i < N; i++) r[i] = a[i] >= b[i];
    return r;
}
template <int N, typename T>
inline vec<N, T> mix(const vec<N, T>& a, const vec<N, T>& b, const vec<N, bool>& t) {
    vec<N, T> r;
    for (int i = 0; i < N; i++) r[i] = t[i] ? b[i] : a[i];
    return r;
}
template <int N, typename T>
inline vec<N, T> mix(const vec<N, T>& a, const vec<N, T>& b, const vec<N, T>& t) {
    vec<N, T> r;
    for (int i = 0; i < N; i++) r[i] = a[i] * (T(1) - t[i]) + b[i] * t[i];
    return r;
}
template <int N, typename T>
inline T dot(const vec<N, T>& a, const vec<N, T>& b) {
    T s = 0;
    for (int i = 0; i < N; i++) s += a[i] * b[i];
    return s;
}

// IEEE binary32 -> binary16, round to nearest even (GLSL packHalf2x16)
// geometric functions as GLM defines them (func_geometric.inl): cross by components, normalize = v * inversesqrt(dot(v, v))
template <typename T>
inline vec<3, T> cross(const vec<3, T>& x, const vec<3, T>& y) {
    return vec<3, T>(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
template <int N, typename T>
inline T length(const vec<N, T>& v) { return std::sqrt(dot(v, v)); }
template <int N, typename T>
inline vec<N, T> normalize(const vec<N, T>& v) { return v * (T(1) / std::sqrt(dot(v, v))); }
template <int N, typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>
constexpr vec<N, T> max(const vec<N, T>& a, S s) { return map1(a, [s](T v) { return max(v, (T)s); }); }
template <int N, typename T, typename S, typename = std::enable_if_t<std::is_arithmetic_v<S>>>
constexpr vec<N, T> min(const vec<N, T>& a, S s) { return map1(a, [s](T v) { return min(v, (T)s); }); }

inline uint16_t glms_f32_to_f16(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u, ax = x & 0x7FFFFFFFu;
    if (ax > 0x7F800000u) return (uint16_t)(sign | 0x7E00u);
    if (ax >= 0x47800000u) return (uint16_t)(sign | 0x7C00u);
    if (ax >= 0x38800000u) {
        uint32_t mant = ax & 0x7FFFFFu, h = (((ax >> 23) - 112) << 10) | (mant >> 13), rem = mant & 0x1FFFu;
        if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) h++;
        return (uint16_t)(sign | h);
    }
    if (ax < 0x33000000u) return (uint16_t)sign;
    uint32_t mant = (ax & 0x7FFFFFu) | 0x800000u, shift = 126 - (ax >> 23), h = mant >> shift, rem = mant & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) h++;
    return (uint16_t)(sign | h);
}
inline uint32_t packHalf2x16(const vec2& v) { return (uint32_t)glms_f32_to_f16(v.x) | ((uint32_t)glms_f32_to_f16(v.y) << 16); }

}  // namespace glm

#include "mat4x4.hpp"
