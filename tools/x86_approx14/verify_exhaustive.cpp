// Compares the rcp14 / rsqrt14 emulations — the product's (voxelrt_b200/csrc/x86_approx14.h, host side of the same functions the
// kernels call) and the oracle's (orc_x86_rcp14 / orc_x86_rsqrt14 in liboracle.so) — with the INSTRUCTIONS on all 2^32 inputs.
// Needs an AVX-512 host.  Build + run: tests/test_x86_approx14.py, or
//   g++ -O2 -fopenmp -mavx512f -mavx512vl verify_exhaustive.cpp -L../../oracle -l:liboracle.so -Wl,-rpath,../../oracle
#include <immintrin.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../voxelrt_b200/csrc/x86_approx14.h"
extern "C" float orc_x86_rsqrt14(float);
extern "C" float orc_x86_rcp14(float);

static float hw_rsq(float x) { return _mm_cvtss_f32(_mm_rsqrt14_ss(_mm_setzero_ps(), _mm_set_ss(x))); }
static float hw_rcp(float x) { return _mm_cvtss_f32(_mm_rcp14_ss(_mm_setzero_ps(), _mm_set_ss(x))); }

int main(int argc, char** argv) {
    // stride 1 = exhaustive; a larger odd stride samples (quick mode)
    const unsigned long long stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1ull;
    unsigned long long bad[5] = {0, 0, 0, 0, 0}, n = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad[:5], n)
    for (long long hi = 0; hi < 65536; hi++)
        for (unsigned long long lo = (unsigned long long)hi % stride; lo < 65536; lo += stride) {
            const uint32_t u = ((uint32_t)hi << 16) | (uint32_t)lo;
            const float x = vrt_x86::u2f(u);
            const uint32_t rc = vrt_x86::f2u(hw_rcp(x)), rs = vrt_x86::f2u(hw_rsq(x));
            bad[0] += rc != vrt_x86::f2u(vrt_x86::rcp14(x));
            bad[1] += rs != vrt_x86::f2u(vrt_x86::rsqrt14(x));
            if (u >= 0x00800000u && u < 0x7F800000u) bad[2] += rs != vrt_x86::f2u(vrt_x86::rsqrt14_pos_normal(x));
            bad[3] += rc != vrt_x86::f2u(orc_x86_rcp14(x));
            bad[4] += rs != vrt_x86::f2u(orc_x86_rsqrt14(x));
            n++;
        }
    printf("inputs %llu mismatches product rcp14 %llu rsqrt14 %llu rsqrt14_pos_normal %llu oracle rcp14 %llu rsqrt14 %llu\n", n, bad[0], bad[1], bad[2],
           bad[3], bad[4]);
    return (bad[0] | bad[1] | bad[2] | bad[3] | bad[4]) ? 1 : 0;
}
