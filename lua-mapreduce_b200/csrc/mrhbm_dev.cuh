// mrhbm_dev.cuh -- device-side record model, key ordering, partitioners.
//
// Record layouts (little endian words), see include/mrhbm.h:
//   RB=16      : w[0..1] = u64 key, w[2..3] = u64 value (emitters store u32 value, zero pad;
//                in-place partial sums of the big-bin path use all 64 bits)
//   RB=32/64/128: w[0..KW) = key bytes zero padded (no NUL inside), w[KW] = u32 value
// Key order = the reference's keys_sorted / merge order (mapreduce/utils.lua:123-128,214):
// C-locale bytewise, shorter first == big-endian word compare of the zero padded slot;
// u64 keys are the reference's 8-byte big-endian strings == numeric order.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "mrhbm_kernels.h"

namespace mrhbm {

template <int RB>
struct Rec {
  static constexpr int kWords = RB / 4;
  static constexpr bool kU64 = (RB == 16);
  static constexpr int kKeyWords = kU64 ? 2 : kWords - 1;  // 32-bit words of key
  static constexpr int kKeyBytes = kKeyWords * 4;
  static constexpr int kVec = RB / 16;  // uint4 per record
};

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  return mix64(x + 0x9E3779B97F4A7C15ull);
}
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// i-th most significant 32-bit word of the key, as an unsigned big-endian number
template <int RB>
__device__ __forceinline__ uint32_t be_word(const uint32_t* r, int i) {
  if constexpr (Rec<RB>::kU64) {
    return i == 0 ? r[1] : (i == 1 ? r[0] : 0u);
  } else {
    return i < Rec<RB>::kKeyWords ? bswap32(r[i]) : 0u;
  }
}
template <int RB>
__device__ __forceinline__ uint64_t key_prefix64(const uint32_t* r) {
  return ((uint64_t)be_word<RB>(r, 0) << 32) | be_word<RB>(r, 1);
}
// nbits (<=51) of the key starting `bitpos` bits below the most significant bit
template <int RB>
__device__ __forceinline__ uint64_t key_bits(const uint32_t* r, int bitpos, int nbits) {
  int wi = bitpos >> 5, sh = bitpos & 31;
  uint64_t hi = ((uint64_t)be_word<RB>(r, wi) << 32) | be_word<RB>(r, wi + 1);
  uint64_t lo = (uint64_t)be_word<RB>(r, wi + 2) << 32;
  uint64_t win = sh ? ((hi << sh) | (lo >> (64 - sh))) : hi;
  return win >> (64 - nbits);
}
template <int RB>
__device__ __forceinline__ int key_cmp(const uint32_t* a, const uint32_t* b) {
  if constexpr (Rec<RB>::kU64) {
    uint64_t x = (uint64_t)a[0] | ((uint64_t)a[1] << 32), y = (uint64_t)b[0] | ((uint64_t)b[1] << 32);
    return (x > y) - (x < y);
  } else {
#pragma unroll
    for (int i = 0; i < Rec<RB>::kKeyWords; i++) {
      uint32_t x = be_word<RB>(a, i), y = be_word<RB>(b, i);
      if (x != y) return x < y ? -1 : 1;
    }
    return 0;
  }
}
template <int RB>
__device__ __forceinline__ bool key_eq(const uint32_t* a, const uint32_t* b) {
  bool e = true;
#pragma unroll
  for (int i = 0; i < Rec<RB>::kKeyWords; i++) e &= (a[i] == b[i]);
  return e;
}
template <int RB>
__device__ __forceinline__ uint64_t rec_value(const uint32_t* r) {
  if constexpr (Rec<RB>::kU64)
    return (uint64_t)r[2] | ((uint64_t)r[3] << 32);
  else
    return r[Rec<RB>::kKeyWords];
}

// ---- partitioners ---------------------------------------------------------
// examples/WordCount/partitionfn.lua:8-16 evaluated exactly as Lua 5.2 does, i.e. in IEEE
// doubles: h*16777619 reaches ~2^56, so the product is rounded to 53 significant bits
// (round-to-nearest-even) BEFORE "% 2^32".  Emulated with integers so no FP64 is issued.
// Only (h * 16777619 rounded to 53 bits) mod 2^32 is needed, and the rounding only touches the
// low `sh` bits, so everything is computed from the 32-bit halves of the product:
// sh = bitlen(product) - 53 = bitlen(high word) - 21.
__host__ __device__ __forceinline__ uint32_t fnv_lua_step(uint32_t h, uint32_t byte) {
  uint32_t lo = h * 16777619u;
#ifdef __CUDA_ARCH__
  uint32_t hi = __umulhi(h, 16777619u);
  int sh = 11 - __clz((int)hi);  // (32 - clz) - 21
#else
  uint32_t hi = (uint32_t)(((uint64_t)h * 16777619ull) >> 32);
  int sh = hi ? 11 - __builtin_clz(hi) : -21;
#endif
  if (sh > 0) {
    uint32_t half = 1u << (sh - 1), rem = lo & ((1u << sh) - 1u);
    uint32_t q = lo >> sh;  // bit 0 of q is bit sh of the product: the ties-to-even parity
    if (rem > half || (rem == half && (q & 1u))) q++;
    lo = q << sh;           // a carry out of bit 31 vanishes mod 2^32, as it must
  }
  return lo ^ byte;
}
// h over the key bytes (u64 keys: the 8 big-endian bytes; strings: up to the first NUL)
template <int RB>
__device__ __forceinline__ uint32_t fnv_lua_hash(const uint32_t* r) {
  uint32_t h = 2166136261u;
  if constexpr (Rec<RB>::kU64) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint32_t w = i < 4 ? r[1] : r[0];
      h = fnv_lua_step(h, (w >> (24 - 8 * (i & 3))) & 0xff);
    }
  } else {
    // The word loop is unrolled (the record lives in registers), the byte loop is NOT: fully unrolled,
    // 4 records x 28 bytes of fnv_lua_step overflowed the instruction cache (ncu on k_hist<32>: 43 % of
    // the stall samples were "no instruction").  Key bytes 0x00 / 0x01 are stored escaped (01 01 / 01 02,
    // mrhbm_emit_str); the reference's partitionfn sees the original bytes, so they are restored here.
    bool esc = false;
#pragma unroll
    for (int i = 0; i < Rec<RB>::kKeyWords; i++) {
      uint32_t w = r[i];
      if (w == 0) break;  // keys hold no NUL: a zero word is past the end (short keys leave early)
#pragma unroll 1
      do {
        const uint32_t b = w & 0xffu;
        if (esc) {
          h = fnv_lua_step(h, b - 1u);
          esc = false;
        } else if (b == 1u) {
          esc = true;
        } else {
          h = fnv_lua_step(h, b);
        }
        w >>= 8;
      } while (w);
      if ((r[i] >> 24) == 0) break;
    }
  }
  return h;
}
// 64-bit word hash over the little-endian u32 words of the key bytes while non-zero
// (length-exact for NUL-free keys, independent of the slot width)
template <int RB, bool EARLY = true>
__device__ __forceinline__ uint64_t word_hash(const uint32_t* r) {
  uint64_t h = 0x9E3779B97F4A7C15ull;
  bool live = true;
#pragma unroll
  for (int i = 0; i < Rec<RB>::kKeyWords; i++) {
    uint32_t w;
    if constexpr (Rec<RB>::kU64)
      w = bswap32(i == 0 ? r[1] : r[0]);  // LE load of the 8 big-endian key bytes
    else
      w = r[i];
    if (EARLY) {
      if (w == 0) break;  // short keys leave early
    } else {
      live = live && (w != 0);
      if (!live) continue;
    }
    h = (h ^ w) * 0xBF58476D1CE4E5B9ull;
    h ^= h >> 29;
  }
  h ^= h >> 32;
  h *= 0x94D049BB133111EBull;
  return h;
}

// floor(x * m / 2^64) for a 32-bit m: two 32x32 multiplies instead of the four of __umul64hi
__host__ __device__ __forceinline__ uint32_t mulhi_u64_u32(uint64_t x, uint32_t m) {
  const uint64_t lo = (uint64_t)(uint32_t)x * m, hi = (uint64_t)(uint32_t)(x >> 32) * m;
  return (uint32_t)((hi + (lo >> 32)) >> 32);
}

// partition id, owning rank and bin (= slot(pid)*S + sub) of a record held in registers / smem words
// WORLD: 0 = bp.world read at run time, 1 = one GPU, 2 = several (folded by the caller's template)
template <int RB, int WORLD = 0>
__device__ __forceinline__ uint32_t bin_of(const uint32_t* r, const BinParams& bp, uint32_t* pid_out,
                                           uint32_t* dest_out = nullptr) {
  uint32_t pid;
  uint64_t h = 0;  // uniform 64-bit hash; feeds the hash sub-bin (only computed further when needed)
  const bool mul = bp.partitioner == 1u && Rec<RB>::kU64;
  if (mul) {  // MULHASH
    uint64_t key = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
    h = key * 0x9E3779B97F4A7C15ull;
    pid = mulhi_u64_u32(h, bp.P);
  } else if (bp.partitioner == 0u) {  // FNV_LUA
    pid = fnv_lua_hash<RB>(r) % bp.P;
  } else {  // WORDHASH
    h = word_hash<RB>(r);
    pid = mulhi_u64_u32(h, bp.P);
  }
  uint32_t sub = 0;
  if (bp.S > 1) {
    uint64_t src;
    if (bp.ordered) {
      src = key_prefix64<RB>(r);
    } else if (bp.partitioner == 0u) {
      src = word_hash<RB>(r);
    } else {
      src = mix64(h * (uint64_t)bp.P + 0x632BE59BD9B4E019ull);  // independent of pid
    }
    sub = mulhi_u64_u32(src, bp.S);
  }
  if (pid_out) *pid_out = pid;
  // partition p is owned by rank p % world and sits in slot pbase[rank] + p / world
  uint32_t slot = pid, dest = 0;
  if (WORLD == 2 || (WORLD == 0 && bp.world > 1)) {
    uint32_t q;
    if (bp.wshift != 0xffffffffu) {  // power-of-two world: no integer division
      dest = pid & (bp.world - 1u);
      q = pid >> bp.wshift;
    } else {
      dest = pid % bp.world;
      q = pid / bp.world;
    }
    slot = bp.pbase[dest] + q;
  }
  if (dest_out) *dest_out = dest;
  return slot * bp.S + sub;
}

// checksum mixers over the whole key slot (input and result share the slot format)
template <int RB>
__device__ __forceinline__ void key_mix2(const uint32_t* r, uint64_t& f1, uint64_t& f2) {
  uint64_t a = 0x243F6A8885A308D3ull, b = 0x13198A2E03707344ull;
#pragma unroll
  for (int i = 0; i < Rec<RB>::kKeyWords; i++) {
    a = mix64(a ^ r[i]);
    b = (b ^ r[i]) * 0x9FB21C651E98DF25ull;
    b ^= b >> 28;
  }
  f1 = a;
  f2 = mix64(b);
}

// ---- shared-memory hash tables of whole records (k_combine, k_agg_bins) --------------------
// 128-bit volatile shared loads: one or two per probe instead of a word-by-word compare (a
// record-strided word access hits only 4 bank groups)
__device__ __forceinline__ uint4 lds128_volatile(const uint4* p) {
  uint4 v;
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
// does table entry e hold the key of record w?  (the value/state word is ignored)
template <int RB>
__device__ __forceinline__ bool entry_key_eq_scalar(const uint32_t* e, const uint32_t* w) {
  bool eq = true;
#pragma unroll
  for (int k = 0; k < Rec<RB>::kKeyWords; k++) eq &= (((volatile const uint32_t*)e)[k] == w[k]);
  return eq;
}
template <int RB>
__device__ __forceinline__ bool entry_key_eq(const uint32_t* e, const uint32_t* w) {
  constexpr int KW = Rec<RB>::kKeyWords;
  bool eq = true;
#pragma unroll
  for (int v = 0; v < Rec<RB>::kVec; v++) {
    uint4 x = lds128_volatile((const uint4*)e + v);
    uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (4 * v + k < KW) eq &= (xs[k] == w[4 * v + k]);
    if (!eq) break;
  }
  return eq;
}

// ---- streaming global access ------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

}  // namespace mrhbm
