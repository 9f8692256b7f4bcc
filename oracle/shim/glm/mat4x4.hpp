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
inline tmat4<T> operator*(const tmat4<T>& a, const tmat4<T>& b) {
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
inline tmat4<T> translate(const tmat4<T>& m, const vec<3, T>& v) {
    tmat4<T> r = m;
    r[3] = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3];
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
template <typename T>
inline tmat4<T> inverse(const tmat4<T>& m) {  // Gauss-Jordan in double; inputs are camera matrices
    double a[4][8];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            a[r][c] = (double)m[c][r];
            a[r][4 + c] = r == c ? 1.0 : 0.0;
        }
    for (int i = 0; i < 4; i++) {
        int p = i;
        for (int r = i + 1; r < 4; r++)
            if (std::fabs(a[r][i]) > std::fabs(a[p][i])) p = r;
        for (int c = 0; c < 8; c++) std::swap(a[i][c], a[p][c]);
        double d = a[i][i];
        for (int c = 0; c < 8; c++) a[i][c] /= d;
        for (int r = 0; r < 4; r++)
            if (r != i) {
                double f = a[r][i];
                for (int c = 0; c < 8; c++) a[r][c] -= f * a[i][c];
            }
    }
    tmat4<T> out(T(0));
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) out[c][r] = (T)a[r][4 + c];
    return out;
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
