#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cstring>
#include <cuda_runtime.h>
#include "mrhbm_dev.cuh"
int main(){
  // host check of the 32-bit emulation against real doubles over many h values and bytes
  uint64_t bad=0; uint32_t x=12345;
  for (long it=0; it<30000000; it++){ x = x*1664525u+1013904223u; uint32_t h=x; uint32_t b=(x>>7)&0xff;
    volatile double prod=(double)h*16777619.0; double m=prod - floor(prod/4294967296.0)*4294967296.0; uint32_t want=((uint32_t)m)^b;
    uint32_t got=mrhbm::fnv_lua_step(h,b); if(got!=want){ if(bad<5) printf("h=%u b=%u want=%u got=%u\n",h,b,want,got); bad++; } }
  // edge values
  uint32_t edges[]={0,1,2,255,256,0x7fffffff,0x80000000,0xffffffff,536870911,536870912,536870913,0x1fffffff,0x3fffffff};
  for (uint32_t h: edges){ volatile double prod=(double)h*16777619.0; double m=prod - floor(prod/4294967296.0)*4294967296.0; uint32_t want=((uint32_t)m)^7; if (mrhbm::fnv_lua_step(h,7)!=want){printf("edge h=%u\n",h);bad++;} }
  printf("bad=%llu\n",(unsigned long long)bad); return bad!=0; }
