// Column-major 4x4 matrix stand-in (see glm.hpp in this directory).  m[col][row], like glm::mat4.
#pragma once
#include "glm.hpp"

namespace glm {

template <typename T>
struct tmat4 {
    vec<4, T> c[4];
    constexpr tmat4() : c{vec<4, T>(1, 0, 0, 0), vec<4, T>(0, 1, 0, 0), vec<4, T>(0, 0, 1, 0), vec<4, T>(0, 0, 0, 1)} {}
    constexpr explicit tmat4(T d) : c{vec<4, T>(d, 0, 0, 0), vec<4, T>(0, d, 0, 0), vec<4, T>(0, 0, d, 0), vec<4, T>(0, 0, 0, d)} {}
    constexpr vec<4, T>& operator[](int i) { return c[i]; }
    constexpr const vec<4, T>& operator[](int i) const { return c[i]; }
};
using mat4 = tmat4<float>;
using dmat4 = tmat4<double>;

template <typename T>
__attribute__((optimize("fp-contract=off"))) inline tmat4<T> operator*(const tmat4<T>& a, const tmat4<T>& b) {
    tmat4<T> r(T(0));
    for (int col = 0; col < 4; col++)
        for (int row = 0; row < 4; row++) {
            T acc = 0;
            for (int k = 0; k < 4; k++) acc += a[k][row] * b[col][k];
            r[col][row] = acc;
        }
    return r;
}
template <typename T>
inline vec<4, T> operator*(const tmat4<T>& m, const vec<4, T>& v) {
    vec<4, T> r;
    for (int row = 0; row < 4; row++) r[row] = m[0][row] * v.x + m[1][row] * v.y + m[2][row] * v.z + m[3][row] * v.w;
    return r;
}
template <typename T>
__attribute__((optimize("fp-contract=off"))) inline tmat4<T> translate(const tmat4<T>& m, const vec<3, T>& v) {
    tmat4<T> r = m;
    for (int row = 0; row < 4; row++) r[3][row] = ((m[0][row] * v.x + m[1][row] * v.y) + m[2][row] * v.z) + m[3][row];
    return r;
}
template <typename T>
inline tmat4<T> scale(const tmat4<T>& m, const vec<3, T>& v) {
    tmat4<T> r = m;
    r[0] = m[0] * v.x;
    r[1] = m[1] * v.y;
    r[2] = m[2] * v.z;
    return r;
}
// glm::inverse(mat4) as GLM's func_matrix.inl defines it: cofactor expansion through 18 2x2 sub-determinants in the matrix's own
// precision, one reciprocal of the determinant, and a final scaling.  Evaluated one rounding at a time (VRT_NOFMA: GCC would
// otherwise contract a*b - c*d into FMAs at -march=native) so that the adapter (voxelrt_b200/host) and scenes/camera.py, which
// restate the same operation sequence, agree with it bit for bit.
#if defined(__GNUC__) && !defined(__clang__)
#define VRT_NOFMA __attribute__((optimize("fp-contract=off")))
#else
#define VRT_NOFMA
#endif
template <typename T>
VRT_NOFMA inline tmat4<T> inverse(const tmat4<T>& m) {
    const T c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3], c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    const T c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3], c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3], c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    const T c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2], c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2], c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    const T c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3], c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3], c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    const T c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2], c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2], c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    const T c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1], c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1], c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    const T f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    const T f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    const T v0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, v1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
    const T v2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, v3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
    tmat4<T> inv(T(0));
    for (int i = 0; i < 4; i++) {
        const T sa = (i & 1) ? T(-1) : T(1), sb = -sa;
        inv[0][i] = ((v1[i] * f0[i] - v2[i] * f1[i]) + v3[i] * f2[i]) * sa;
        inv[1][i] = ((v0[i] * f0[i] - v2[i] * f3[i]) + v3[i] * f4[i]) * sb;
        inv[2][i] = ((v0[i] * f1[i] - v1[i] * f3[i]) + v3[i] * f5[i]) * sa;
        inv[3][i] = ((v0[i] * f2[i] - v1[i] * f4[i]) + v2[i] * f5[i]) * sb;
    }
    const T d0 = m[0][0] * inv[0][0], d1 = m[0][1] * inv[1][0], d2 = m[0][2] * inv[2][0], d3 = m[0][3] * inv[3][0];
    const T one_over_det = T(1) / ((d0 + d1) + (d2 + d3));
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) inv[c][r] = inv[c][r] * one_over_det;
    return inv;
}
template <typename T>
inline tmat4<T> perspective(T fovy, T aspect, T zNear, T zFar) {  // right-handed, depth -1..1
    T t = std::tan(fovy / T(2));
    tmat4<T> r(T(0));
    r[0][0] = T(1) / (aspect * t);
    r[1][1] = T(1) / t;
    r[2][2] = -(zFar + zNear) / (zFar - zNear);
    r[2][3] = -T(1);
    r[3][2] = -(T(2) * zFar * zNear) / (zFar - zNear);
    return r;
}
template <typename T>
inline tmat4<T> eulerAngleXY(T ax, T ay) {
    T cx = std::cos(ax), sx = std::sin(ax), cy = std::cos(ay), sy = std::sin(ay);
    tmat4<T> r;
    r[0] = vec<4, T>(cy, -sx * -sy, cx * -sy, 0);
    r[1] = vec<4, T>(0, cx, sx, 0);
    r[2] = vec<4, T>(sy, -sx * cy, cx * cy, 0);
    r[3] = vec<4, T>(0, 0, 0, 1);
    return r;
}

}  // namespace glm
