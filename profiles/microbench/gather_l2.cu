// gather_l2.cu -- how fast can a B200 read random 64-byte buckets (4 x 16 B per thread) of a 32 MB table that
// should live in L2?  Decides the flavour of the loads in the combiner's global-table walk (round 2: the walk
// cost 12 ms for 2.8e8 lookups with ld.relaxed.gpu + L2 cache hints -- 1.5 TB/s, far below what L2 serves).
// Variants: weak .ca (default), weak .cg, ld.relaxed.gpu, ld.volatile, relaxed + evict_last hint, and 32-byte
// accesses (2 x 16 B).  Reports G lookups/s and GB/s of useful bytes.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  return x ^ (x >> 16);
}
enum { WEAK_CA, WEAK_CG, RELAXED, VOLATILE_, RELAXED_HINT };
template <int MODE>
__device__ __forceinline__ uint4 ld16(const uint4* p, uint64_t pol) {
  uint4 v;
  if (MODE == WEAK_CA) asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  if (MODE == WEAK_CG) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  if (MODE == RELAXED) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  if (MODE == VOLATILE_) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  if (MODE == RELAXED_HINT)
    asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol) : "memory");
  return v;
}
template <int MODE, int VEC /* 16-byte loads per lookup */>
__global__ void __launch_bounds__(1024, 1) k_gather(const uint4* tab, uint32_t log_buckets, int iters, uint32_t* out) {
  uint64_t pol = 0;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
  for (int it = 0; it < iters; it++) {
    x = hash32(x + it);
    const uint4* e = tab + (size_t)(x >> (32 - log_buckets)) * 4;
    uint4 v[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k++) v[k] = ld16<MODE>(e + k, pol);
#pragma unroll
    for (int k = 0; k < VEC; k++) acc += v[k].x ^ v[k].w;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const uint32_t log_buckets = 19;  // 2^19 buckets x 64 B = 32 MB
  uint4* tab;
  uint32_t* out;
  cudaMalloc(&tab, (size_t)64 << log_buckets);
  cudaMemset(tab, 1, (size_t)64 << log_buckets);
  cudaMalloc(&out, (size_t)sms * 1024 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 512;
  printf("# random 64-byte bucket reads from a 32 MB table, %d CTAs x 1024 threads x %d lookups\n", sms, iters);
  auto run = [&](const char* name, int vec, auto launch) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0);
      launch();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    const double n = (double)sms * 1024 * iters;
    printf("%-28s %7.3f ms  %7.2f G lookups/s  %7.1f GB/s  %s\n", name, best, n / best * 1e-6, n * vec * 16 / best * 1e-6,
           cudaGetLastError() == cudaSuccess ? "" : "ERR");
  };
  run("weak .ca 4x16B", 4, [&] { k_gather<WEAK_CA, 4><<<sms, 1024>>>(tab, log_buckets, iters, out); });
  run("weak .cg 4x16B", 4, [&] { k_gather<WEAK_CG, 4><<<sms, 1024>>>(tab, log_buckets, iters, out); });
  run("relaxed.gpu 4x16B", 4, [&] { k_gather<RELAXED, 4><<<sms, 1024>>>(tab, log_buckets, iters, out); });
  run("volatile 4x16B", 4, [&] { k_gather<VOLATILE_, 4><<<sms, 1024>>>(tab, log_buckets, iters, out); });
  run("relaxed.gpu+evict_last 4x16B", 4, [&] { k_gather<RELAXED_HINT, 4><<<sms, 1024>>>(tab, log_buckets, iters, out); });
  run("weak .cg 2x16B", 2, [&] { k_gather<WEAK_CG, 2><<<sms, 1024>>>(tab, log_buckets, iters, out); });
  run("relaxed.gpu 2x16B", 2, [&] { k_gather<RELAXED, 2><<<sms, 1024>>>(tab, log_buckets, iters, out); });
  run("weak .cg 1x16B", 1, [&] { k_gather<WEAK_CG, 1><<<sms, 1024>>>(tab, log_buckets, iters, out); });
  run("relaxed.gpu 1x16B", 1, [&] { k_gather<RELAXED, 1><<<sms, 1024>>>(tab, log_buckets, iters, out); });
  return 0;
}
