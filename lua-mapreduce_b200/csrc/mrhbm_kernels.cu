// mrhbm_kernels.cu -- hand-written sm_100a kernels of the in-HBM shuffle:
//   k_hist / k_exscan / k_scatter : hash-partition into (partition, sub-bin) bins
//                                   (replaces partitionfn + per-partition spill files,
//                                    mapreduce/job.lua:203-221)
//   k_sort_reduce / k_big_bins    : per-bin shared-memory sort by key + segmented sum
//                                   (replaces keys_sorted + heap merge + reducer loop,
//                                    mapreduce/utils.lua:123-128,206-271, job.lua:264-284)
//   k_compact, k_checksum_*       : result gathering and parity properties
//   k_gen_*                       : synthetic device-side mapfn
// HBM-bound integer work: no tensor cores.  Grids are multiples of the SM count.
#include "mrhbm_kernels.h"

#include "mrhbm_dev.cuh"

namespace mrhbm {

static int g_sm_count = 148;
constexpr int kSortThreads = 512;
constexpr int kIdxBits = 12;  // kCapBytes/16 = 4096 records at most
constexpr uint32_t kIdxMask = (1u << kIdxBits) - 1;
constexpr int kDigitBits = 64 - kIdxBits;

// ============================================================================
// synthetic device-side mapfn
// ============================================================================
__global__ void k_gen_u64(uint4* dst, uint64_t seed, uint64_t start, uint64_t n) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t k = splitmix64(seed + start + i);
    uint32_t v = (uint32_t)(splitmix64(seed + (1ull << 40) + start + i) >> 32);
    stg_stream(dst + i, make_uint4((uint32_t)k, (uint32_t)(k >> 32), v, 0u));
  }
}

__device__ __forceinline__ int rank_to_key_dev(uint64_t rank, unsigned char* out /*28, zeroed*/) {
  unsigned char tmp[16];
  int n = 0;
  uint64_t r = rank;
  while (r > 0) {
    r -= 1;
    tmp[n++] = (unsigned char)('a' + (r % 26));
    r /= 26;
  }
  int o = 0;
  while (n > 0) out[o++] = tmp[--n];
  uint64_t h = splitmix64(rank ^ 0xA5A5A5A5A5A5A5A5ull);
  unsigned l = (unsigned)(h % 8);
  if (((h >> 8) % 64) == 0) l = 8 + (unsigned)((h >> 16) % 15);
  for (unsigned j = 0; j < l; j++) out[o++] = (unsigned char)('A' + (splitmix64(h + j) % 26));
  return o;
}

__global__ void k_gen_zipf32(uint4* dst, uint64_t seed, uint64_t start, uint64_t n,
                             const uint64_t* __restrict__ table, uint64_t V) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t u = splitmix64(seed + (1ull << 41) + start + i);
    uint64_t lo = 0, hi = V;
    while (lo < hi) {
      uint64_t mid = (lo + hi) >> 1;
      if (__ldg(table + mid) < u)
        lo = mid + 1;
      else
        hi = mid;
    }
    uint64_t rank = lo + 1 > V ? V : lo + 1;
    union {
      unsigned char b[32];
      uint4 v[2];
    } rec;
    rec.v[0] = make_uint4(0, 0, 0, 0);
    rec.v[1] = make_uint4(0, 0, 0, 0);
    rank_to_key_dev(rank, rec.b);
    rec.v[1].w = 1u;  // value
    stg_stream(dst + 2 * i, rec.v[0]);
    stg_stream(dst + 2 * i + 1, rec.v[1]);
  }
}

// ============================================================================
// histogram + scatter
// ============================================================================
template <int RB>
__device__ __forceinline__ void load_rec(const uint4* p, uint32_t* w) {
#pragma unroll
  for (int v = 0; v < Rec<RB>::kVec; v++) {
    uint4 x = ldg_stream(p + v);
    w[4 * v + 0] = x.x;
    w[4 * v + 1] = x.y;
    w[4 * v + 2] = x.z;
    w[4 * v + 3] = x.w;
  }
}

template <int RB>
__global__ void __launch_bounds__(256) k_hist(const uint4* __restrict__ recs, uint64_t n, BinParams bp,
                                              uint32_t* __restrict__ hist) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t w[Rec<RB>::kWords];
    load_rec<RB>(recs + i * Rec<RB>::kVec, w);
    uint32_t bin = bin_of<RB>(w, bp, nullptr);
    atomicAdd(hist + bin, 1u);  // RED: no return value
  }
}

template <int RB>
__global__ void __launch_bounds__(256) k_scatter(const uint4* __restrict__ recs, uint64_t n, BinParams bp,
                                                 uint32_t* __restrict__ cursor, uint4* __restrict__ mid) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t w[Rec<RB>::kWords];
    load_rec<RB>(recs + i * Rec<RB>::kVec, w);
    uint32_t bin = bin_of<RB>(w, bp, nullptr);
    uint32_t pos = atomicAdd(cursor + bin, 1u);
    uint4* d = mid + (uint64_t)pos * Rec<RB>::kVec;
#pragma unroll
    for (int v = 0; v < Rec<RB>::kVec; v++)
      stg_stream(d + v, make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]));
  }
}

// single-CTA exclusive scan (B <= a few million): out_excl[0..n], optional copy into
// out_copy (scatter cursors), optional list of entries larger than cap
__global__ void __launch_bounds__(1024) k_exscan(const uint32_t* __restrict__ in, uint32_t n,
                                                 uint32_t* __restrict__ out_excl,
                                                 uint32_t* __restrict__ out_copy, uint32_t cap,
                                                 uint32_t* __restrict__ big_list, uint32_t* nbig,
                                                 uint32_t* total) {
  __shared__ uint32_t warp_sums[32];
  const uint32_t T = blockDim.x, tid = threadIdx.x;
  uint32_t per = (n + T - 1) / T;
  uint32_t b0 = tid * per, b1 = min(n, b0 + per);
  uint32_t s = 0;
  for (uint32_t i = b0; i < b1; i++) s += in[i];
  uint32_t incl = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if ((tid & 31) >= (uint32_t)d) incl += t;
  }
  if ((tid & 31) == 31) warp_sums[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    uint32_t w = tid < (T >> 5) ? warp_sums[tid] : 0u, wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
      if (tid >= (uint32_t)d) wi += t;
    }
    warp_sums[tid] = wi - w;  // exclusive
  }
  __syncthreads();
  uint32_t run = warp_sums[tid >> 5] + incl - s;
  for (uint32_t i = b0; i < b1; i++) {
    uint32_t c = in[i];
    out_excl[i] = run;
    if (out_copy) out_copy[i] = run;
    if (big_list && c > cap) big_list[atomicAdd(nbig, 1u)] = i;
    run += c;
  }
  if (tid == T - 1) {
    out_excl[n] = run;
    if (total) *total = run;
  }
}

// ============================================================================
// per-bin sort + segmented reduce in shared memory
// ============================================================================
struct SortSmem {
  uint4* rec;        // cap records
  uint64_t* comp;    // cap composite words (digit << kIdxBits | index)
  uint16_t* permA;   // cap
  uint16_t* permB;   // cap
  uint64_t* red;     // 64 words of reduction scratch
};
__host__ __device__ inline size_t sort_smem_bytes(int rb) {
  size_t cap = kCapBytes / rb;
  return (size_t)kCapBytes + cap * 8 + cap * 2 * 2 + 64 * 8;
}
__device__ __forceinline__ SortSmem carve(unsigned char* base, int rb) {
  size_t cap = kCapBytes / rb;
  SortSmem s;
  s.rec = (uint4*)base;
  s.comp = (uint64_t*)(base + kCapBytes);
  s.permA = (uint16_t*)(base + kCapBytes + cap * 8);
  s.permB = s.permA + cap;
  s.red = (uint64_t*)(base + kCapBytes + cap * 8 + cap * 4);
  return s;
}

// ascending bitonic sort of comp[0..n2), n2 a power of two >= 64
__device__ __forceinline__ void bitonic_sort(uint64_t* comp, uint32_t n2) {
  const uint32_t tid = threadIdx.x, T = blockDim.x, half = n2 >> 1;
  for (uint32_t k = 2; k <= n2; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t t = tid; t < half; t += T) {
        uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // insert a 0 bit at log2(j)
        uint32_t l = i | j;
        uint64_t a = comp[i], b = comp[l];
        bool up = (i & k) == 0;
        if ((a > b) == up) {
          comp[i] = b;
          comp[l] = a;
        }
      }
      __syncthreads();
    }
  }
}

// in-place exclusive scan of a[0..n) (n <= 4096) by the whole CTA; returns the total
__device__ __forceinline__ uint32_t block_exscan(uint32_t* a, uint32_t n, uint32_t* scratch33) {
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  uint32_t per = (n + T - 1) / T;
  uint32_t b0 = min(n, tid * per), b1 = min(n, b0 + per);
  uint32_t s = 0;
  for (uint32_t i = b0; i < b1; i++) s += a[i];
  uint32_t incl = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if ((tid & 31) >= (uint32_t)d) incl += t;
  }
  if ((tid & 31) == 31) scratch33[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    uint32_t w = tid < (T >> 5) ? scratch33[tid] : 0u, wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
      if (tid >= (uint32_t)d) wi += t;
    }
    scratch33[tid] = wi - w;
    if (tid == 31) scratch33[32] = wi;
  }
  __syncthreads();
  uint32_t run = scratch33[tid >> 5] + incl - s;
  for (uint32_t i = b0; i < b1; i++) {
    uint32_t c = a[i];
    a[i] = run;
    run += c;
  }
  uint32_t total = scratch33[32];
  __syncthreads();
  return total;
}

enum { MODE_FINAL = 0, MODE_PARTIAL = 1 };
struct ChunkOut {
  void* keys;          // FINAL: key slots; PARTIAL: AoS records (uint4)
  uint64_t* sums;      // FINAL only
  uint64_t base;       // element offset into the destination
  uint32_t* err_flags;
};

// Sorts cnt (<= cap) records of one bin by key, sums the values of equal keys and writes
// the groups in ascending key order.  Returns the number of groups.
template <int RB, int MODE, bool NC>
__device__ uint32_t process_chunk(const SortSmem& sm, const uint4* __restrict__ src, uint32_t cnt,
                                  const ChunkOut& out) {
  using R = Rec<RB>;
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  const uint32_t* recw = (const uint32_t*)sm.rec;
  // 1. coalesced load of the bin
  for (uint32_t v = tid; v < cnt * R::kVec; v += T) sm.rec[v] = NC ? ldg_stream(src + v) : src[v];
  __syncthreads();
  // 2. range of the 64-bit key prefix
  uint64_t pmin = ~0ull, pmax = 0;
  for (uint32_t i = tid; i < cnt; i += T) {
    uint64_t p = key_prefix64<RB>(recw + i * R::kWords);
    pmin = p < pmin ? p : pmin;
    pmax = p > pmax ? p : pmax;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    uint64_t a = __shfl_xor_sync(0xffffffffu, pmin, d), b = __shfl_xor_sync(0xffffffffu, pmax, d);
    pmin = a < pmin ? a : pmin;
    pmax = b > pmax ? b : pmax;
  }
  if ((tid & 31) == 0) {
    sm.red[tid >> 5] = pmin;
    sm.red[32 + (tid >> 5)] = pmax;
  }
  __syncthreads();
  pmin = ~0ull;
  pmax = 0;
  for (uint32_t w = 0; w < (T >> 5); w++) {
    uint64_t a = sm.red[w], b = sm.red[32 + w];
    pmin = a < pmin ? a : pmin;
    pmax = b > pmax ? b : pmax;
  }
  __syncthreads();
  uint64_t range = pmax - pmin;
  int bits = range ? 64 - __clzll((long long)range) : 0;
  int drop = bits > kDigitBits ? bits - kDigitBits : 0;
  uint32_t n2 = 64;
  while (n2 < cnt) n2 <<= 1;
  // 3. composite words: most significant kDigitBits of the normalised prefix + index
  for (uint32_t i = tid; i < n2; i += T) {
    uint64_t c = ~0ull;
    if (i < cnt) c = (((key_prefix64<RB>(recw + i * R::kWords) - pmin) >> drop) << kIdxBits) | i;
    sm.comp[i] = c;
  }
  __syncthreads();
  bitonic_sort(sm.comp, n2);
  // 4. permutation + do equal digits hide different keys?
  int tie = 0;
  for (uint32_t j = tid; j < cnt; j += T) {
    uint64_t c = sm.comp[j];
    sm.permA[j] = (uint16_t)(c & kIdxMask);
    if (j > 0 && (R::kKeyWords > 2 || drop > 0)) {
      uint64_t p = sm.comp[j - 1];
      if ((c >> kIdxBits) == (p >> kIdxBits) &&
          !key_eq<RB>(recw + (uint32_t)(c & kIdxMask) * R::kWords,
                      recw + (uint32_t)(p & kIdxMask) * R::kWords))
        tie = 1;
    }
  }
  uint16_t* perm = sm.permA;
  uint16_t* other = sm.permB;
  if (__syncthreads_or(tie)) {
    // 5. rare: LSD passes over kDigitBits-wide chunks of the whole key; the previous rank
    //    in the low bits makes every pass stable
    constexpr int kKeyBits = R::kKeyWords * 32;
    constexpr int kChunks = (kKeyBits + kDigitBits - 1) / kDigitBits;
    for (int c = kChunks - 1; c >= 0; c--) {
      int bitpos = c * kDigitBits;
      int nb = kKeyBits - bitpos < kDigitBits ? kKeyBits - bitpos : kDigitBits;
      for (uint32_t j = tid; j < n2; j += T) {
        uint64_t w = ~0ull;
        if (j < cnt) w = (key_bits<RB>(recw + (uint32_t)perm[j] * R::kWords, bitpos, nb) << kIdxBits) | j;
        sm.comp[j] = w;
      }
      __syncthreads();
      bitonic_sort(sm.comp, n2);
      for (uint32_t j = tid; j < cnt; j += T) other[j] = perm[sm.comp[j] & kIdxMask];
      __syncthreads();
      uint16_t* t = perm;
      perm = other;
      other = t;
    }
  }
  // 6. group heads and their output slots
  uint32_t* slot = (uint32_t*)sm.comp;  // comp is free now
  for (uint32_t j = tid; j < cnt; j += T) {
    uint32_t head = 1;
    if (j > 0) head = !key_eq<RB>(recw + (uint32_t)perm[j] * R::kWords, recw + (uint32_t)perm[j - 1] * R::kWords);
    other[j] = (uint16_t)head;
    slot[j] = head;
  }
  __syncthreads();
  uint32_t groups = block_exscan(slot, cnt, (uint32_t*)sm.red);
  // 7. segmented sum by the head thread and write-out
  for (uint32_t j = tid; j < cnt; j += T) {
    if (!other[j]) continue;
    const uint32_t* r = recw + (uint32_t)perm[j] * R::kWords;
    uint64_t s = rec_value<RB>(r);
    for (uint32_t k = j + 1; k < cnt && !other[k]; k++) s += rec_value<RB>(recw + (uint32_t)perm[k] * R::kWords);
    uint64_t o = out.base + slot[j];
    if (MODE == MODE_FINAL) {
      if constexpr (R::kU64) {
        ((uint64_t*)out.keys)[o] = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
      } else {
        uint32_t* d = (uint32_t*)out.keys + o * R::kKeyWords;
#pragma unroll
        for (int w = 0; w < R::kKeyWords; w++) d[w] = r[w];
      }
      out.sums[o] = s;
    } else {
      uint32_t* d = (uint32_t*)out.keys + o * R::kWords;
#pragma unroll
      for (int w = 0; w < R::kKeyWords; w++) d[w] = r[w];
      if constexpr (R::kU64) {
        d[2] = (uint32_t)s;
        d[3] = (uint32_t)(s >> 32);
      } else {
        if (s > 0xffffffffull) atomicOr(out.err_flags, (uint32_t)ERRF_OVERFLOW);
        d[R::kKeyWords] = (uint32_t)s;
      }
    }
  }
  __syncthreads();
  return groups;
}

template <int RB>
__global__ void __launch_bounds__(kSortThreads, 2)
    k_sort_reduce(ShuffleBuffers b, uint32_t B, uint32_t cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_bin;
  SortSmem sm = carve(smem_raw, RB);
  for (;;) {
    if (threadIdx.x == 0) s_bin = atomicAdd(b.counters + CNT_TICKET, 1u);
    __syncthreads();
    uint32_t bin = s_bin;
    __syncthreads();
    if (bin >= B) break;
    uint32_t off = b.bin_off[bin], cnt = b.bin_off[bin + 1] - off;
    if (cnt > cap) continue;  // k_big_bins
    if (cnt == 0) {
      if (threadIdx.x == 0) b.ucount[bin] = 0;
      continue;
    }
    ChunkOut out{b.out_keys, b.out_sums, off, b.counters + CNT_ERR};
    uint32_t g = process_chunk<RB, MODE_FINAL, true>(sm, (const uint4*)b.mid + (uint64_t)off * Rec<RB>::kVec, cnt, out);
    if (threadIdx.x == 0) b.ucount[bin] = g;
  }
}

// One CTA per oversized bin (hot keys): chunk-wise in-place reduce until the bin fits.
template <int RB>
__global__ void __launch_bounds__(kSortThreads, 2) k_big_bins(ShuffleBuffers b, uint32_t cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SortSmem sm = carve(smem_raw, RB);
  uint32_t bin = b.big_list[blockIdx.x];
  uint32_t off = b.bin_off[bin], n = b.bin_off[bin + 1] - off;
  uint4* base = (uint4*)b.mid + (uint64_t)off * Rec<RB>::kVec;
  while (n > cap) {
    uint32_t w = 0;
    for (uint32_t c = 0; c < n; c += cap) {
      uint32_t m = n - c < cap ? n - c : cap;
      ChunkOut out{base, nullptr, w, b.counters + CNT_ERR};
      // groups of a chunk never outnumber the records consumed so far: w + g <= c + m
      w += process_chunk<RB, MODE_PARTIAL, false>(sm, base + (uint64_t)c * Rec<RB>::kVec, m, out);
      __threadfence_block();
    }
    if (w == n) {  // nothing merged: more distinct keys than one CTA can sort
      if (threadIdx.x == 0) {
        atomicOr(b.counters + CNT_ERR, (uint32_t)ERRF_SKEW);
        b.ucount[bin] = 0;
      }
      return;
    }
    n = w;
  }
  ChunkOut out{b.out_keys, b.out_sums, off, b.counters + CNT_ERR};
  uint32_t g = process_chunk<RB, MODE_FINAL, false>(sm, base, n, out);
  if (threadIdx.x == 0) b.ucount[bin] = g;
}

// ============================================================================
// result gathering + parity properties
// ============================================================================
template <int RB>
__global__ void __launch_bounds__(256) k_compact(ShuffleBuffers b, uint32_t B, uint32_t* __restrict__ dkeys,
                                                 uint64_t* __restrict__ dsums) {
  constexpr int KW = Rec<RB>::kKeyWords;
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t* skeys = (const uint32_t*)b.out_keys;
  for (uint32_t bin = warp; bin < B; bin += nwarps) {
    uint32_t len = b.ucount[bin];
    uint64_t so = b.bin_off[bin], d0 = b.uoff[bin];
    for (uint32_t i = lane; i < len * KW; i += 32) dkeys[d0 * KW + i] = skeys[so * KW + i];
    for (uint32_t i = lane; i < len; i += 32) dsums[d0 + i] = b.out_sums[so + i];
  }
}

__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

template <int RB>
__global__ void __launch_bounds__(256) k_checksum_in(const uint4* __restrict__ recs, uint64_t n,
                                                     unsigned long long* acc) {
  uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t w[Rec<RB>::kWords];
    load_rec<RB>(recs + i * Rec<RB>::kVec, w);
    uint64_t f1, f2, v = rec_value<RB>(w);
    key_mix2<RB>(w, f1, f2);
    a0 += f1 * v;
    a1 += f2 * v;
    a2 += v;
    a3 += 1;
  }
  a0 = warp_sum64(a0);
  a1 = warp_sum64(a1);
  a2 = warp_sum64(a2);
  a3 = warp_sum64(a3);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(acc + 0, a0);
    atomicAdd(acc + 1, a1);
    atomicAdd(acc + 2, a2);
    atomicAdd(acc + 3, a3);
  }
}

template <int RB>
__global__ void __launch_bounds__(256) k_checksum_out(ShuffleBuffers b, uint32_t B, BinParams bp,
                                                      unsigned long long* acc) {
  constexpr int KW = Rec<RB>::kKeyWords;
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t* keys = (const uint32_t*)b.out_keys;
  uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, bad_order = 0, bad_part = 0;
  for (uint32_t bin = warp; bin < B; bin += nwarps) {
    uint32_t len = b.ucount[bin];
    uint64_t so = b.bin_off[bin];
    for (uint32_t i = lane; i < len; i += 32) {
      uint32_t w[Rec<RB>::kWords], p[Rec<RB>::kWords];
#pragma unroll
      for (int k = 0; k < KW; k++) w[k] = keys[(so + i) * KW + k];
      uint64_t s = b.out_sums[so + i], f1, f2;
      key_mix2<RB>(w, f1, f2);
      a0 += f1 * s;
      a1 += f2 * s;
      a2 += s;
      a3 += 1;
      if (i > 0) {
#pragma unroll
        for (int k = 0; k < KW; k++) p[k] = keys[(so + i - 1) * KW + k];
        if (key_cmp<RB>(p, w) >= 0) bad_order++;
      }
      uint32_t pid;
      uint32_t mybin = bin_of<RB>(w, bp, &pid);
      if (mybin != bin) bad_part++;
    }
  }
  a0 = warp_sum64(a0);
  a1 = warp_sum64(a1);
  a2 = warp_sum64(a2);
  a3 = warp_sum64(a3);
  bad_order = warp_sum64(bad_order);
  bad_part = warp_sum64(bad_part);
  if (lane == 0) {
    atomicAdd(acc + 0, a0);
    atomicAdd(acc + 1, a1);
    atomicAdd(acc + 2, a2);
    atomicAdd(acc + 3, a3);
    atomicAdd(acc + 4, bad_order);
    atomicAdd(acc + 5, bad_part);
  }
}

// ============================================================================
// launchers
// ============================================================================
static inline int stream_grid(uint64_t n, int threads, int ctas_per_sm) {
  uint64_t need = (n + threads - 1) / threads;
  uint64_t cap = (uint64_t)g_sm_count * ctas_per_sm;
  return (int)(need < cap ? (need ? need : 1) : cap);
}

cudaError_t kernels_configure() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return e;
  g_sm_count = prop.multiProcessorCount;
#define CFG(RB)                                                                                   \
  e = cudaFuncSetAttribute(k_sort_reduce<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                           (int)sort_smem_bytes(RB));                                             \
  if (e != cudaSuccess) return e;                                                                 \
  e = cudaFuncSetAttribute(k_big_bins<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                           (int)sort_smem_bytes(RB));                                             \
  if (e != cudaSuccess) return e;
  CFG(16) CFG(32) CFG(64) CFG(128)
#undef CFG
  return cudaSuccess;
}

#define DISPATCH_RB(rb, CALL) \
  switch (rb) {               \
    case 16: { constexpr int RB = 16; CALL; } break;   \
    case 32: { constexpr int RB = 32; CALL; } break;   \
    case 64: { constexpr int RB = 64; CALL; } break;   \
    default: { constexpr int RB = 128; CALL; } break;  \
  }

int launch_gen_u64(void* dst, uint64_t seed, uint64_t start, uint64_t n, cudaStream_t s) {
  if (!n) return 0;
  k_gen_u64<<<stream_grid(n, 256, 8), 256, 0, s>>>((uint4*)dst, seed, start, n);
  return 1;
}
int launch_gen_zipf32(void* dst, uint64_t seed, uint64_t start, uint64_t n, const uint64_t* d_table,
                      uint64_t V, cudaStream_t s) {
  if (!n) return 0;
  k_gen_zipf32<<<stream_grid(n, 256, 8), 256, 0, s>>>((uint4*)dst, seed, start, n, d_table, V);
  return 1;
}
int launch_hist(int rb, const void* recs, uint64_t n, uint32_t P, uint32_t S, uint32_t partitioner,
                uint32_t ordered, uint32_t* hist, cudaStream_t s) {
  if (!n) return 0;
  BinParams bp{P, S, partitioner, ordered};
  DISPATCH_RB(rb, (k_hist<RB><<<stream_grid(n, 256, 8), 256, 0, s>>>((const uint4*)recs, n, bp, hist)));
  return 1;
}
int launch_exscan(const uint32_t* in, uint32_t n, uint32_t* out_excl, uint32_t* out_copy, uint32_t cap,
                  uint32_t* big_list, uint32_t* nbig, uint32_t* total, cudaStream_t s) {
  k_exscan<<<1, 1024, 0, s>>>(in, n, out_excl, out_copy, cap, big_list, nbig, total);
  return 1;
}
int launch_scatter(int rb, const void* recs, uint64_t n, uint32_t P, uint32_t S, uint32_t partitioner,
                   uint32_t ordered, uint32_t* cursor, void* mid, cudaStream_t s) {
  if (!n) return 0;
  BinParams bp{P, S, partitioner, ordered};
  DISPATCH_RB(rb, (k_scatter<RB><<<stream_grid(n, 256, 8), 256, 0, s>>>((const uint4*)recs, n, bp, cursor,
                                                                       (uint4*)mid)));
  return 1;
}
int launch_sort_reduce(int rb, const ShuffleBuffers& b, uint32_t B, uint32_t cap, int sm_count,
                       cudaStream_t s) {
  int grid = (int)(B < (uint32_t)(2 * sm_count) ? B : (uint32_t)(2 * sm_count));
  if (grid < 1) grid = 1;
  DISPATCH_RB(rb, (k_sort_reduce<RB><<<grid, kSortThreads, sort_smem_bytes(RB), s>>>(b, B, cap)));
  return 1;
}
int launch_big_bins(int rb, const ShuffleBuffers& b, uint32_t nbig, uint32_t cap, cudaStream_t s) {
  if (!nbig) return 0;
  DISPATCH_RB(rb, (k_big_bins<RB><<<nbig, kSortThreads, sort_smem_bytes(RB), s>>>(b, cap)));
  return 1;
}
int launch_compact(int rb, const ShuffleBuffers& b, uint32_t B, void* dst_keys, uint64_t* dst_sums,
                   cudaStream_t s) {
  int grid = g_sm_count * 8;
  DISPATCH_RB(rb, (k_compact<RB><<<grid, 256, 0, s>>>(b, B, (uint32_t*)dst_keys, dst_sums)));
  return 1;
}
int launch_checksum_in(int rb, const void* recs, uint64_t n, uint64_t* acc4, cudaStream_t s) {
  if (!n) return 0;
  DISPATCH_RB(rb, (k_checksum_in<RB><<<stream_grid(n, 256, 8), 256, 0, s>>>((const uint4*)recs, n,
                                                                           (unsigned long long*)acc4)));
  return 1;
}
int launch_checksum_out(int rb, const ShuffleBuffers& b, uint32_t B, uint32_t P, uint32_t S,
                        uint32_t partitioner, uint32_t ordered, uint64_t* acc6, cudaStream_t s) {
  BinParams bp{P, S, partitioner, ordered};
  int grid = g_sm_count * 8;
  DISPATCH_RB(rb, (k_checksum_out<RB><<<grid, 256, 0, s>>>(b, B, bp, (unsigned long long*)acc6)));
  return 1;
}

}  // namespace mrhbm
