"""Camera matrices as the reference builds them (input definition, float32 throughout).

Follows LibGlimpsw/Common/Camera.h:17-34 (perspective FOV 90, near 0.01, far 1000, view =
rotation only), VoxelRT/GBuffer.h:51 (CurrentProj = P * V) and GBuffer.h:133-139
(GetInverseProjScreenMat).  glm is absent from this image; these are restatements of the glm
formulas (column-major).  The matrices are computed ONCE here and handed, bit-identical, to
both the oracle and the CUDA path, so they can never cause a CPU/GPU divergence.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def perspective(fovy_rad, aspect, near, far):
    """glm::perspective (RH, depth -1..1). Returned as m[col][row]."""
    t = f32(np.tan(f32(fovy_rad) / f32(2)))
    m = np.zeros((4, 4), f32)
    m[0][0] = f32(1) / (f32(aspect) * t)
    m[1][1] = f32(1) / t
    m[2][2] = -(f32(far) + f32(near)) / (f32(far) - f32(near))
    m[2][3] = -f32(1)
    m[3][2] = -(f32(2) * f32(far) * f32(near)) / (f32(far) - f32(near))
    return m


def euler_angle_xy(ax, ay):
    """glm::eulerAngleXY(angleX, angleY) = Rx(ax) * Ry(ay), column-major."""
    cx, sx, cy, sy = (f32(np.cos(f32(ax))), f32(np.sin(f32(ax))), f32(np.cos(f32(ay))), f32(np.sin(f32(ay))))
    m = np.zeros((4, 4), f32)
    m[0] = [cy, -sx * -sy, cx * -sy, 0]
    m[1] = [0, cx, sx, 0]
    m[2] = [sy, -sx * cy, cx * cy, 0]
    m[3] = [0, 0, 0, 1]
    return m


def matmul(a, b):
    """glm a*b for column-major storage m[col][row]: (a*b)[c][r] = sum_k a[k][r] * b[c][k]."""
    out = np.zeros((4, 4), f32)
    for c in range(4):
        for r in range(4):
            acc = f32(0)
            for k in range(4):
                acc = f32(acc + a[k][r] * b[c][k])
            out[c][r] = acc
    return out


def translate(m, v):
    """glm::translate(m, v): m * T(v)."""
    out = m.copy()
    out[3] = ((m[0] * f32(v[0]) + m[1] * f32(v[1])).astype(f32) + m[2] * f32(v[2])).astype(f32) + m[3]
    return out.astype(f32)


def scale(m, v):
    out = m.copy()
    out[0] = m[0] * f32(v[0])
    out[1] = m[1] * f32(v[1])
    out[2] = m[2] * f32(v[2])
    return out.astype(f32)


def inverse(m):
    """glm::inverse(mat4): GLM's float cofactor expansion, one float32 rounding per operation — the same operation sequence as
    oracle/shim/glm/mat4x4.hpp (what the reference's GetInverseProjScreenMat runs on in oracle/_ref) and the adapter's Inverse()
    (voxelrt_b200/host/b200_renderer.cpp); tests pin the three against each other bit for bit."""
    m = np.asarray(m, f32)

    def det2(a, b, c, d):  # a*b - c*d
        return f32(f32(a * b) - f32(c * d))

    c00, c02, c03 = det2(m[2][2], m[3][3], m[3][2], m[2][3]), det2(m[1][2], m[3][3], m[3][2], m[1][3]), det2(m[1][2], m[2][3], m[2][2], m[1][3])
    c04, c06, c07 = det2(m[2][1], m[3][3], m[3][1], m[2][3]), det2(m[1][1], m[3][3], m[3][1], m[1][3]), det2(m[1][1], m[2][3], m[2][1], m[1][3])
    c08, c10, c11 = det2(m[2][1], m[3][2], m[3][1], m[2][2]), det2(m[1][1], m[3][2], m[3][1], m[1][2]), det2(m[1][1], m[2][2], m[2][1], m[1][2])
    c12, c14, c15 = det2(m[2][0], m[3][3], m[3][0], m[2][3]), det2(m[1][0], m[3][3], m[3][0], m[1][3]), det2(m[1][0], m[2][3], m[2][0], m[1][3])
    c16, c18, c19 = det2(m[2][0], m[3][2], m[3][0], m[2][2]), det2(m[1][0], m[3][2], m[3][0], m[1][2]), det2(m[1][0], m[2][2], m[2][0], m[1][2])
    c20, c22, c23 = det2(m[2][0], m[3][1], m[3][0], m[2][1]), det2(m[1][0], m[3][1], m[3][0], m[1][1]), det2(m[1][0], m[2][1], m[2][0], m[1][1])
    f0, f1, f2 = (c00, c00, c02, c03), (c04, c04, c06, c07), (c08, c08, c10, c11)
    f3, f4, f5 = (c12, c12, c14, c15), (c16, c16, c18, c19), (c20, c20, c22, c23)
    v0 = (m[1][0], m[0][0], m[0][0], m[0][0])
    v1 = (m[1][1], m[0][1], m[0][1], m[0][1])
    v2 = (m[1][2], m[0][2], m[0][2], m[0][2])
    v3 = (m[1][3], m[0][3], m[0][3], m[0][3])

    def comb(a, fa, b, fb, c, fc, s):  # ((a*fa - b*fb) + c*fc) * s
        return f32(f32(f32(f32(a * fa) - f32(b * fb)) + f32(c * fc)) * s)

    inv = np.zeros((4, 4), f32)
    for i in range(4):
        sa = f32(-1.0 if i & 1 else 1.0)
        sb = f32(-sa)
        inv[0][i] = comb(v1[i], f0[i], v2[i], f1[i], v3[i], f2[i], sa)
        inv[1][i] = comb(v0[i], f0[i], v2[i], f3[i], v3[i], f4[i], sb)
        inv[2][i] = comb(v0[i], f1[i], v1[i], f3[i], v3[i], f5[i], sa)
        inv[3][i] = comb(v0[i], f2[i], v1[i], f4[i], v2[i], f5[i], sb)
    d0, d1, d2, d3 = f32(m[0][0] * inv[0][0]), f32(m[0][1] * inv[1][0]), f32(m[0][2] * inv[2][0]), f32(m[0][3] * inv[3][0])
    with np.errstate(divide="ignore", invalid="ignore"):
        one_over_det = f32(f32(1) / f32(f32(d0 + d1) + f32(d2 + d3)))
        return (inv * one_over_det).astype(f32)


def inverse_proj_screen(proj_view, width, height):
    """GBuffer::GetInverseProjScreenMat (GBuffer.h:133-139)."""
    inv = inverse(proj_view)
    inv = translate(inv, (-1.0, -1.0, 0.0))
    inv = scale(inv, (f32(2.0) / f32(width), f32(2.0) / f32(height), 1.0))
    inv = translate(inv, (0.5, 0.5, 0.0))
    return inv


class Camera:
    """Reference defaults: Main.cpp:76-78 (pos (512,128,512), yaw 1.52, pitch -0.5)."""

    def __init__(self, pos=(512.0, 128.0, 512.0), yaw=1.52, pitch=-0.5, fov_deg=90.0, near=0.01, far=1000.0):
        self.pos = np.asarray(pos, np.float64)
        self.yaw, self.pitch = yaw, pitch
        self.fov_deg, self.near, self.far = fov_deg, near, far

    def matrices(self, width, height):
        """-> (proj[16], inv_proj_screen[16], world_origin[3] int, origin_frac[3] f32)."""
        aspect = f32(width) / f32(height)
        p = perspective(np.radians(f32(self.fov_deg)), aspect, self.near, self.far)
        v = euler_angle_xy(-self.pitch, self.yaw)  # Camera.h:42 destRotation = eulerAngleXY(-Euler.y, Euler.x)
        pv = matmul(p, v)
        inv = inverse_proj_screen(pv, width, height)
        wo = np.floor(self.pos).astype(np.int32)  # CpuRenderer.cpp:447
        frac = (self.pos - np.floor(self.pos)).astype(f32)  # :448
        return pv.reshape(16).copy(), inv.reshape(16).copy(), wo, frac


def orbit_cameras(n, seed=1, center=(384.0, 110.0, 384.0), radius=260.0, height=(100.0, 200.0)):
    """Seeded pose set used next to the default camera so nothing is tuned to one view."""
    rng = np.random.default_rng(seed)
    cams = []
    for i in range(n):
        ang = 2 * np.pi * (i + rng.random() * 0.5) / n
        r = radius * (0.35 + 0.65 * rng.random())
        y = height[0] + (height[1] - height[0]) * rng.random()
        pos = (center[0] + r * np.cos(ang), y, center[2] + r * np.sin(ang))
        # look roughly toward the centre, slightly down
        yaw = float(np.arctan2(center[0] - pos[0], -(center[2] - pos[2])))
        pitch = float(-0.25 - 0.5 * rng.random())
        cams.append(Camera(pos=pos, yaw=yaw, pitch=pitch))
    return cams
