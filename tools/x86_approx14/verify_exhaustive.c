#include <immintrin.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <omp.h>
#include "tables.inc"
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
static float hw_rsq(float x){ return _mm_cvtss_f32(_mm_rsqrt14_ss(_mm_setzero_ps(), _mm_set_ss(x))); }
static float hw_rcp(float x){ return _mm_cvtss_f32(_mm_rcp14_ss(_mm_setzero_ps(), _mm_set_ss(x))); }

static uint32_t emu_rcp14(uint32_t u){
    uint32_t sign=u&0x80000000u, a=u&0x7FFFFFFFu;
    if(a>0x7F800000u) return u|0x00400000u;        // NaN -> quiet
    if(a==0x7F800000u) return sign;                  // inf -> 0
    if(a==0) return sign|0x7F800000u;                // 0 -> inf
    int e=(int)(a>>23); uint32_t m=a&0x7FFFFFu;
    if(e==0){ // denormal: normalise
        int sh=__builtin_clz(m)-8; m=(m<<sh)&0x7FFFFFu; e=1-sh;
    }
    // x = 1.m * 2^(e-127); result = r * 2^-(e-127), r in (0.5,1]
    uint32_t rb; // bits of r
    if(m==0) rb=0x3F800000u; else { uint32_t idx=m>>17, t=(m>>7)&0x3FFu; rb=((kRcp14[idx][0]-kRcp14[idx][1]*t)>>9)<<7; }
    int re=(int)(rb>>23)-(e-127); uint32_t rm=rb&0x7FFFFFu;
    if(re>=255) return sign|0x7F800000u;
    if(re<=0){ // denormal result
        uint32_t full=rm|0x800000u; int sh=1-re; if(sh>24) return sign; return sign|(full>>sh);
    }
    return sign|((uint32_t)re<<23)|rm;
}
static uint32_t emu_rsqrt14(uint32_t u){
    uint32_t sign=u&0x80000000u, a=u&0x7FFFFFFFu;
    if(a>0x7F800000u) return u|0x00400000u;
    if(a==0) return sign|0x7F800000u;                // +-0 -> +-inf
    if(sign) return 0xFFC00000u;                     // negative -> indefinite
    if(a==0x7F800000u) return 0;                     // +inf -> 0
    int e=(int)(a>>23); uint32_t m=a&0x7FFFFFu;
    if(e==0){ int sh=__builtin_clz(m)-8; m=(m<<sh)&0x7FFFFFu; e=1-sh; }
    int E=e-127; int odd=E&1; int k=(E-odd)/2;      // x = (1.m * 2^odd) * 4^k
    uint32_t rb;
    if(m==0&&!odd) rb=0x3F800000u; else { uint32_t idx=(odd<<5)|(m>>18), t=(m>>8)&0x3FFu; rb=((kRsqrt14[idx][0]-kRsqrt14[idx][1]*t)>>9)<<7; }
    int re=(int)(rb>>23)-k;
    return ((uint32_t)re<<23)|(rb&0x7FFFFFu);
}
int main(){
    unsigned long long bad_rcp=0,bad_rsq=0; uint32_t first_rcp=0,first_rsq=0;
    #pragma omp parallel for schedule(static) reduction(+:bad_rcp,bad_rsq)
    for(long long hi=0;hi<65536;hi++){
        for(uint32_t lo=0;lo<65536;lo++){
            uint32_t u=((uint32_t)hi<<16)|lo; float x=u2f(u);
            uint32_t a=f2u(hw_rcp(x)), b=emu_rcp14(u);
            if(a!=b){ if(!bad_rcp){
                #pragma omp critical
                { if(!first_rcp){ first_rcp=u?u:1; printf("rcp mismatch x=%08x hw=%08x emu=%08x\n",u,a,b);} } }
                bad_rcp++; }
            a=f2u(hw_rsq(x)); b=emu_rsqrt14(u);
            if(a!=b){ if(!bad_rsq){
                #pragma omp critical
                { if(!first_rsq){ first_rsq=u?u:1; printf("rsq mismatch x=%08x hw=%08x emu=%08x\n",u,a,b);} } }
                bad_rsq++; }
        }
    }
    printf("mismatches over all 2^32 inputs: rcp14 %llu, rsqrt14 %llu\n",bad_rcp,bad_rsq);
    return 0;
}
