// Capsule test of the reference's brush for the bench's edit workload (INPUT generation for BASELINE configs[4], `--edit-mode brush`; not the
// product path, not the oracle): which voxels of one 8^3 brick lie inside sdCapsule(p, a, b, r) < 0 (src/VoxelRT/Brush.cpp:4-8,19-20), in the
// reference's arithmetic — voxel centres p = conv2f(X) + 0.5, simd::dot as a chain of FMAs (SIMD.h:109-111), clamp as x86 max / min
// (a NaN quotient — zero-length capsule — becomes 0), the projection pa - ba * h unfused, and simd::length = approx_sqrt = rsqrt14(x) * x
// (SIMD.h:126, SIMD_AVX512.h:136-138; a voxel centre exactly ON the axis gives inf * 0 = NaN and is NOT inside).  tests/test_ref_brush_pin.py
// runs the reference's own BrushSession::Dispatch on the same strokes and demands the same voxels, allocation masks and dirty bricks.
// Built with -ffp-contract=off: every fused operation is written as fmaf.
#include <cmath>
#include <cstdint>

#include "../voxelrt_b200/csrc/x86_approx14.h"

extern "C" __attribute__((visibility("default"))) int brush_brick_mask(const int32_t a[3], const int32_t b[3], float radius, int32_t bx, int32_t by, int32_t bz,
                                                                        uint8_t* inside /* 512, index x | z << 3 | y << 6 */) {
    const float fa[3] = {(float)a[0], (float)a[1], (float)a[2]};
    const float ba[3] = {(float)b[0] - fa[0], (float)b[1] - fa[1], (float)b[2] - fa[2]};
    const float bb = fmaf(ba[0], ba[0], fmaf(ba[1], ba[1], ba[2] * ba[2]));
    int count = 0;
    for (int y = 0; y < 8; y++)
        for (int z = 0; z < 8; z++)
            for (int x = 0; x < 8; x++) {
                const float p[3] = {(float)(bx * 8 + x) + 0.5f, (float)(by * 8 + y) + 0.5f, (float)(bz * 8 + z) + 0.5f};
                const float pa[3] = {p[0] - fa[0], p[1] - fa[1], p[2] - fa[2]};
                const float q = fmaf(pa[0], ba[0], fmaf(pa[1], ba[1], pa[2] * ba[2])) / bb;
                const float m = (q > 0.0f) ? q : 0.0f;  // _mm512_max_ps(q, 0): the second operand unless q > 0 (NaN included)
                const float h = (m < 1.0f) ? m : 1.0f;  // _mm512_min_ps(m, 1)
                const float d[3] = {pa[0] - ba[0] * h, pa[1] - ba[1] * h, pa[2] - ba[2] * h};
                const float dd = fmaf(d[0], d[0], fmaf(d[1], d[1], d[2] * d[2]));
                const float len = vrt_x86::rsqrt14(dd) * dd;
                const bool in = (len - radius) < 0.0f;
                inside[x | (z << 3) | (y << 6)] = in ? 1 : 0;
                count += in;
            }
    return count;
}
