#include <immintrin.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
static float rsq(float x){ return _mm_cvtss_f32(_mm_rsqrt14_ss(_mm_setzero_ps(), _mm_set_ss(x))); }
static float rcp(float x){ return _mm_cvtss_f32(_mm_rcp14_ss(_mm_setzero_ps(), _mm_set_ss(x))); }
int main(int argc,char**argv){
  int which=atoi(argv[1]); // 0 rcp, 1 rsq [1,2), 2 rsq [2,4)
  uint32_t base = which==2?0x40000000u:0x3F800000u;
  // check dependence only on m>>7 (except m==0)
  int bad=0;
  static uint32_t T[65536];
  for(uint32_t i=0;i<65536;i++){
    uint32_t r0=f2u(which?rsq(u2f(base|(i<<7)|1)):rcp(u2f(base|(i<<7)|1)));
    T[i]=r0;
    for(uint32_t l=0;l<128;l++){ if(i==0&&l==0) continue; uint32_t r=f2u(which?rsq(u2f(base|(i<<7)|l)):rcp(u2f(base|(i<<7)|l))); if(r!=r0) bad++; }
  }
  printf("which=%d bad=%d  m=0 -> %08x, T[0]=%08x T[65535]=%08x\n",which,bad,f2u(which?rsq(u2f(base)):rcp(u2f(base))),T[0],T[65535]);
  FILE*f=fopen(which==0?"rcp.bin":which==1?"rsq1.bin":"rsq2.bin","wb"); fwrite(T,4,65536,f); fclose(f);
  return 0;
}
