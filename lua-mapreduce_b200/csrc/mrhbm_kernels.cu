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

#include <algorithm>
#include <type_traits>

#include "mrhbm_dev.cuh"
#include "mrhbm_sort.cuh"

namespace mrhbm {

static int g_sm_count = 148;
static uint32_t g_tune = 0;  // MRHBM_TUNE measurement hooks (set once by kernels_set_tune)

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
// device-side WordCount mapfn: text -> (word, 1) records
// ============================================================================
// Lua's %s in the C locale == isspace: ' ', \t \n \v \f \r (examples/WordCount/mapfn.lua:5)
__device__ __forceinline__ bool tok_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }
__device__ __forceinline__ bool tok_start(const unsigned char* t, uint64_t i) {
  return !tok_space(t[i]) && (i == 0 || tok_space(t[i - 1]));
}
// one thread per byte, one count per 256-byte block
__global__ void __launch_bounds__(256) k_tok_count(const unsigned char* __restrict__ t, uint64_t len,
                                                   uint32_t* __restrict__ block_counts) {
  uint64_t nb = (len + 255) / 256;
  for (uint64_t blk = blockIdx.x; blk < nb; blk += gridDim.x) {
    uint64_t i = blk * 256 + threadIdx.x;
    int st = i < len && tok_start(t, i);
    int c = __syncthreads_count(st);
    if (threadIdx.x == 0) block_counts[blk] = (uint32_t)c;
  }
}
template <int RB>
__global__ void __launch_bounds__(256) k_tok_emit(const unsigned char* __restrict__ t, uint64_t len,
                                                  const uint32_t* __restrict__ block_off, uint4* __restrict__ recs,
                                                  uint32_t* __restrict__ flags) {
  constexpr int KB = Rec<RB>::kKeyBytes;
  __shared__ uint32_t wcount[8];
  uint64_t nb = (len + 255) / 256;
  for (uint64_t blk = blockIdx.x; blk < nb; blk += gridDim.x) {
    uint64_t i = blk * 256 + threadIdx.x;
    bool st = i < len && tok_start(t, i);
    // rank of this word start inside the block: ballot prefix within the warp + warp totals
    uint32_t m = __ballot_sync(0xffffffffu, st);
    uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wcount[warp] = __popc(m);
    __syncthreads();
    uint32_t before = __popc(m & ((1u << lane) - 1u));
    for (uint32_t w = 0; w < warp; w++) before += wcount[w];
    __syncthreads();
    if (!st) continue;
    union {
      unsigned char b[RB];
      uint4 v[RB / 16];
    } rec;
#pragma unroll
    for (int v = 0; v < RB / 16; v++) rec.v[v] = make_uint4(0, 0, 0, 0);
    int l = 0;  // bytes written to the key slot
    uint64_t p = i;
    while (p < len && !tok_space(t[p]) && l < KB) {
      const unsigned char ch = t[p++];
      if (ch <= 1) {  // 0x00 / 0x01 inside a word travel escaped, like mrhbm_emit_str does it
        rec.b[l++] = 1;
        if (l < KB) rec.b[l++] = (unsigned char)(ch + 1);
        else l = KB + 1;
      } else {
        rec.b[l++] = ch;
      }
    }
    if (l >= KB || (p < len && !tok_space(t[p]))) {  // does not fit (one zero byte must remain)
      atomicOr(flags, (uint32_t)ERRF_KEYLEN);
      l = KB - 1;
      rec.b[KB - 1] = 0;
    }
    uint32_t one = 1u;
    rec.b[KB] = (unsigned char)one;  // little-endian u32 value 1 (remaining value bytes are zero)
    uint4* d = recs + ((uint64_t)block_off[blk] + before) * (RB / 16);
#pragma unroll
    for (int v = 0; v < RB / 16; v++) d[v] = rec.v[v];
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

// U records per thread per iteration: all loads first, then all atomics, then all stores, so
// that each thread keeps U independent memory chains in flight (the loop is latency-bound:
// load -> L2 atomic -> store).
template <int RB>
struct Unroll {
  static constexpr int U = RB == 16 ? 8 : RB == 32 ? 4 : RB == 64 ? 2 : 1;
};

template <int RB>
__global__ void __launch_bounds__(256) k_hist(const uint4* __restrict__ recs, uint64_t n, BinParams bp,
                                              uint32_t* __restrict__ hist) {
  constexpr int U = Unroll<RB>::U;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < n; base += stride * U) {
    uint32_t w[U][Rec<RB>::kWords];
#pragma unroll
    for (int k = 0; k < U; k++)
      if (base + k * stride < n) load_rec<RB>(recs + (base + k * stride) * Rec<RB>::kVec, w[k]);
#pragma unroll
    for (int k = 0; k < U; k++) {
      if (base + k * stride < n) {
        uint32_t bin = bin_of<RB>(w[k], bp, nullptr);
        atomicAdd(hist + (((((size_t)bin) << bp.rep_shift) | (blockIdx.x & ((1u << bp.rep_shift) - 1u))) << bp.ctr_shift), 1u);  // RED
      }
    }
  }
}

template <int RB>
__global__ void __launch_bounds__(256) k_scatter(const uint4* __restrict__ recs, uint64_t n, BinParams bp,
                                                 uint32_t* __restrict__ cursor, uint4* __restrict__ mid) {
  constexpr int U = Unroll<RB>::U;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; base < n; base += stride * U) {
    uint32_t w[U][Rec<RB>::kWords];
    uint32_t pos[U];
#pragma unroll
    for (int k = 0; k < U; k++)
      if (base + k * stride < n) load_rec<RB>(recs + (base + k * stride) * Rec<RB>::kVec, w[k]);
#pragma unroll
    for (int k = 0; k < U; k++) {
      if (base + k * stride < n) {
        uint32_t bin = bin_of<RB>(w[k], bp, nullptr);
        pos[k] = atomicAdd(cursor + (((((size_t)bin) << bp.rep_shift) | (blockIdx.x & ((1u << bp.rep_shift) - 1u))) << bp.ctr_shift), 1u);
      }
    }
#pragma unroll
    for (int k = 0; k < U; k++) {
      if (base + k * stride < n) {
        uint4* d = mid + (uint64_t)pos[k] * Rec<RB>::kVec;
#pragma unroll
        for (int v = 0; v < Rec<RB>::kVec; v++)
          stg_stream(d + v, make_uint4(w[k][4 * v], w[k][4 * v + 1], w[k][4 * v + 2], w[k][4 * v + 3]));
      }
    }
  }
}

// Strided sample of the u64 keys, histogram of their top 8 bits: decides BEFORE any record moves whether
// key-ordered sub-bins (sub = mulhi(key, S)) would be balanced.  Sequential or clustered keys show up as
// overfull buckets and the shuffle uses hash sub-bins from the start instead of paying a discarded attempt.
__global__ void __launch_bounds__(256) k_sample_u64(const uint4* __restrict__ recs, uint64_t n, uint32_t nsample,
                                                    uint32_t* __restrict__ hist256) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < nsample; i += gridDim.x * 256) {
    const uint64_t idx = (uint64_t)(((unsigned __int128)i * n) / nsample);
    const uint4 r = __ldg(recs + idx);
    atomicAdd(&h[r.y >> 24], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(hist256 + threadIdx.x, h[threadIdx.x]);
}

// ============================================================================
// two-level coalesced split (many bins)
// ============================================================================
// A per-pair scatter pays one L2 round trip (cursor claim) and one scattered store per pair.
// k_split_tma instead partitions a TILE of records inside shared memory (one shared-memory atomic
// per pair -- measured 4.9 cycles per warp on B200, profiles/microbench/smem_rank.cu), claims global
// space once per (tile, bin) and copies the tile out bin by bin, so stores are contiguous runs.
// With F fine bins per coarse region it runs twice:
//   level 1: source -> coarse regions (local memory), level 2: coarse region -> its F fine bins.
// On several GPUs level 2 IS the exchange: a coarse region belongs to the rank that owns its
// partitions; every rank's level 1 fills its own copy of ALL regions, and level 2 of the owner pulls
// region y from every rank's buffer (peer-mapped memory over NVLink) with the same 40 KB bulk copies
// that feed it on one GPU, tile by tile, while it splits the previous tile -- transfer and split
// overlap inside one kernel, and the sort then runs on local data only.  Replaces
// mapreduce/job.lua:203-221 (partitionfn + spill) and the GridFS / scp transport (job.lua:255-260,
// fs.lua:143-160).
constexpr int kSplitMaxBins = 1024;

struct SplitArgs {
  const uint4* src;       // level 1: records
  uint64_t n;             // level 1: number of records
  // level 2 source: region y = blockIdx.y of this rank as filled by source rank z = blockIdx.z:
  // seg_stride records at peer[z] + (region_first + y) * seg_stride, holding
  // seg_counts[z * seg_zstride + ((region_first + y) << ctr_shift)] records
  unsigned long long peer[8];  // every rank's region buffer as mapped into this process (one GPU: peer[0] = l1)
  const uint32_t* seg_counts;
  uint64_t seg_zstride;
  uint64_t seg_stride;
  uint32_t region_first;
  uint4* dst;             // destination regions
  uint64_t dst_stride;    // records per destination region (optimistic layout)
  uint32_t* cursor;       // destination fill levels (index << ctr_shift)
  uint32_t capacity;      // records a destination region can take
  uint32_t F;             // fine bins per coarse region (a power of two)
  uint32_t logF;
  uint32_t nbins;         // bins this level distinguishes inside one tile (<= kSplitMaxBins)
  uint32_t level;         // 1 or 2
  uint32_t ctr_shift;
  uint32_t* err_flags;
  // exact layout (after k_hist + k_exscan): bin f starts at base_off[f << rep_shift]; nullptr = the
  // optimistic fixed-stride layout
  const uint32_t* base_off;
  uint32_t rep_shift;
  uint32_t B;
  // optimistic layout: rank d owns the coarse regions [rbase[d], rbase[d+1]) = its fine bins
  // [fbase[d], fbase[d+1]) in blocks of F (one GPU: rbase = {0, C1}, fbase = {0, B})
  uint32_t rbase[9], fbase[9];
  uint32_t ndest, me;
  unsigned long long* span;  // measurement hook or nullptr (SplitPlan::span)
  uint32_t* ticket;          // level 1, optimistic layout: {next tile - gridDim.x, CTAs that have left}, both 0 at launch
};

constexpr int kTmaSplitThreads = 512;
constexpr int kTmaTileBytes = kSplitTileBytes;
__host__ __device__ constexpr size_t tma_split_smem(int rb, int tile_bytes = kTmaTileBytes) {
  return 2 * (size_t)tile_bytes + (size_t)(tile_bytes / rb) * sizeof(uint32_t);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// one thread: arm the barrier with the byte count and start the bulk copy global -> shared
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of dst are done
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// One thread issues ONE cp.async.bulk per tile (global -> shared, completion on an mbarrier),
// double-buffered, so tile t+1 lands while tile t is split; the tile is not re-staged: a
// permutation says which raw record goes to which output position, and the copy-out reads the raw
// tile through it.
// The kernel is issue bound (profiles/README.md: 57 % issue utilisation, time follows the instruction count), so
// what is known at launch time is folded at compile time.  SPEC = 0: level, number of GPUs and layout are read from
// the arguments (exact layouts, the rare paths).  Otherwise the optimistic layout with bits 0-1 = level and
// bit 2 = several GPUs.
template <int RB, int TILE_BYTES = kTmaTileBytes, int MINB = 2, int THREADS_ = kTmaSplitThreads, int SPEC = 0>
__global__ void __launch_bounds__(THREADS_, MINB) k_split_tma(SplitArgs a, BinParams bp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using R = Rec<RB>;
  constexpr int THREADS = THREADS_;
  constexpr int T = TILE_BYTES / RB;              // records per tile
  constexpr int U = (T + THREADS - 1) / THREADS;     // records per thread
  constexpr int IPT = kSplitMaxBins / THREADS;       // bins per thread in the scan
  constexpr int WORLD = SPEC ? ((SPEC & 4) ? 2 : 1) : 0;
  constexpr bool DYN = SPEC != 0 && (SPEC & 3) == 1;  // tiles from a ticket counter (level 1, optimistic layout)
  static_assert(T < 65536 && kSplitMaxBins <= 65536, "perm packs (raw index, bin) into 16 + 16 bits");
  uint32_t* perm = (uint32_t*)(smem_raw + 2 * TILE_BYTES);  // output position -> raw index | bin << 16
  __shared__ uint32_t scnt[kSplitMaxBins], soff[kSplitMaxBins];
  // this tile: the slot (record index in a.dst, 32-bit wrap-around arithmetic) output position 0 would have in the bin's run
  __shared__ uint32_t sbase[kSplitMaxBins];
  __shared__ uint32_t wsum[THREADS / 32];
  __shared__ uint32_t s_rbase[9], s_fbase[9];  // (kernel-parameter arrays indexed by a register would be copied to local memory)
  __shared__ unsigned long long s_src[8];      // byte address and length (records) of the tile streams of this CTA
  __shared__ uint32_t s_n[8];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t s_next_len[2];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t level = SPEC ? (uint32_t)(SPEC & 3) : a.level;
  const bool exact = SPEC ? false : a.base_off != nullptr;
  const bool multi = SPEC ? (SPEC & 4) != 0 : a.ndest > 1;
  if (a.span && tid == 0 && level == 1) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMin(a.span + 0, t);  // first CTA start
    atomicMax(a.span + 2, t);  // last CTA start
  }
#pragma unroll
  for (int i = 0; i < 9; i++)
    if (tid == (uint32_t)i) {
      s_rbase[i] = a.rbase[i];
      s_fbase[i] = a.fbase[i];
    }
  const uint4* src = a.src;
  uint64_t n = a.n;
  uint32_t coarse = 0;
  if (level == 2) {
    coarse = blockIdx.y;
    if (exact) {  // the coarse region is the union of its fine bins
      uint32_t f0 = coarse * a.F, f1 = f0 + a.F < a.B ? f0 + a.F : a.B;
      uint32_t o0 = a.base_off[(size_t)f0 << a.rep_shift], o1 = a.base_off[(size_t)f1 << a.rep_shift];
      src += (size_t)o0 * R::kVec;
      n = o1 - o0;
    }
  }
  // The tile streams this CTA splits: one (the source range) or, for optimistic level 2, region `coarse` as every
  // rank filled it -- ndest streams, the remote ones read over NVLink.  The CTA takes tiles x, x + gridDim.x, ... of
  // EVERY stream, round-robin over the streams, so that its remote bulk copies fly while it splits local tiles
  // (with one CTA per source the ranks first did all their local tiles, then sat on the link: level 2 took
  // local time + link time).
  const bool region_streams = level == 2 && !exact;
  const uint32_t nsrc = (region_streams && multi) ? a.ndest : 1u;
#pragma unroll
  for (int z = 0; z < 8; z++) {
    if (tid == (uint32_t)z && (uint32_t)z < nsrc) {
      if (!region_streams) {
        s_src[0] = (unsigned long long)(uintptr_t)src;
        s_n[0] = (uint32_t)n;
      } else {
        s_src[z] = a.peer[z] + (unsigned long long)(a.region_first + coarse) * a.seg_stride * RB;
        const uint32_t c = a.seg_counts[(size_t)z * a.seg_zstride + ((size_t)(a.region_first + coarse) << a.ctr_shift)];
        s_n[z] = c < a.seg_stride ? c : (uint32_t)a.seg_stride;
      }
    }
  }
  uint32_t fine_base = a.fbase[0];  // first fine bin of this rank
  if (multi) {
#pragma unroll
    for (int z = 1; z < 8; z++)
      if (a.me == (uint32_t)z) fine_base = a.fbase[z];
  }
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t b = tid; b < kSplitMaxBins; b += THREADS) scnt[b] = 0;
  __syncthreads();
  uint32_t ntmax = 0;
  for (uint32_t z = 0; z < nsrc; z++) ntmax = max(ntmax, (s_n[z] + T - 1) / T);
  // (tiles x, x + gridDim.x, ...: a contiguous range of tiles per CTA measured 4 % slower)
  const uint32_t rounds = ntmax > blockIdx.x ? (ntmax - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
  // the CTA's steps in order: round r = 0, 1, ... (tile blockIdx.x + r * gridDim.x), inside a round the streams
  // zrot, zrot + 1, ... (mod nsrc); kept as counters, advanced without a division
  const uint32_t zrot = nsrc > 1 ? (blockIdx.x + a.me) % nsrc : 0u;
  uint32_t st_r = 0, st_zi = 0, st_z = zrot;
  const uint4* cur_p = nullptr;
  uint32_t cur_len = 0;
  // describes the step the counters stand on (cur_p, cur_len), skipping empty ones; false when the CTA is done
  auto settle = [&]() -> bool {
    while (st_r < rounds) {
      const uint64_t t0 = (uint64_t)(blockIdx.x + st_r * gridDim.x) * T;
      const uint32_t nz = s_n[st_z];
      if (t0 < nz) {
        cur_p = (const uint4*)(uintptr_t)s_src[st_z] + t0 * R::kVec;
        cur_len = (nz - t0) < (uint64_t)T ? (uint32_t)(nz - t0) : (uint32_t)T;
        return true;
      }
      if (++st_zi == nsrc) {
        st_zi = 0;
        st_r++;
      }
      if (++st_z == nsrc) st_z = 0;
    }
    cur_len = 0;
    return false;
  };
  auto advance = [&]() -> bool {
    if (++st_zi == nsrc) {
      st_zi = 0;
      st_r++;
    }
    if (++st_z == nsrc) st_z = 0;
    return settle();
  };
  bool more = DYN ? false : settle();
  if (!DYN && tid == 0 && more) bulk_load(smem_raw, cur_p, cur_len * RB, &mbar[0]);

  // one tile of tn records in `raw`; FULL: tn == T, no per-record bounds checks
  auto split_tile = [&](auto full_c, const uint4* raw, const uint32_t tn) {
    constexpr bool FULL = decltype(full_c)::value;
    uint32_t sub[U], rk[U];
#pragma unroll
    for (int k = 0; k < U; k++) {
      const uint32_t i = tid + k * THREADS;
      if (((k + 1) * THREADS <= T || i < (uint32_t)T) && (FULL || i < tn)) {
        uint32_t dest;
        const uint32_t fine = bin_of<RB, WORLD>((const uint32_t*)(raw + (size_t)i * R::kVec), bp, nullptr, &dest);
        if (level == 1) {
          if (multi)
            sub[k] = s_rbase[dest] + ((fine - s_fbase[dest]) >> a.logF);
          else
            sub[k] = fine >> a.logF;
        } else {
          sub[k] = (fine - fine_base) & (a.F - 1u);
        }
        rk[k] = atomicAdd(scnt + sub[k], 1u);
      }
    }
    __syncthreads();
    // exclusive scan of the per-bin counts; the thread that owns a bin also claims its global space
    uint32_t v[IPT], ex[IPT], g[IPT], s = 0;
#pragma unroll
    for (int i = 0; i < IPT; i++) {
      v[i] = scnt[IPT * tid + i];  // (bins >= a.nbins stay zero)
      scnt[IPT * tid + i] = 0;     // ready for the next tile: every count of this one has been taken
      s += v[i];
    }
#pragma unroll
    for (int i = 0; i < IPT; i++) {
      g[i] = 0;
      const uint32_t b = IPT * tid + i;
      if (v[i]) {
        const uint32_t dbin = level == 1 ? b : coarse * a.F + b;
        g[i] = atomicAdd(a.cursor + ((size_t)dbin << a.ctr_shift), v[i]);  // result needed only after the placement
      }
    }
    uint32_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= (uint32_t)d) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    {
      uint32_t ws = lane < THREADS / 32 ? wsum[lane] : 0u, wi = ws;
#pragma unroll
      for (int d = 1; d < THREADS / 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= (uint32_t)d) wi += t;
      }
      uint32_t e = __shfl_sync(0xffffffffu, wi - ws, warp) + incl - s;
#pragma unroll
      for (int i = 0; i < IPT; i++) {
        ex[i] = e;
        soff[IPT * tid + i] = e;
        e += v[i];
      }
    }
    __syncthreads();
    // which raw record goes to which output position
#pragma unroll
    for (int k = 0; k < U; k++) {
      const uint32_t i = tid + k * THREADS;
      if (((k + 1) * THREADS <= T || i < (uint32_t)T) && (FULL || i < tn)) perm[soff[sub[k]] + rk[k]] = i | (sub[k] << 16);
    }
#pragma unroll
    for (int i = 0; i < IPT; i++) {
      const uint32_t b = IPT * tid + i;
      if (v[i]) {
        const uint32_t dbin = level == 1 ? b : coarse * a.F + b;
        uint32_t gg = g[i], first;
        if (exact) {  // absolute start of the destination region, always large enough
          const uint32_t f = level == 1 ? b * a.F : dbin;
          first = a.base_off[(size_t)f << a.rep_shift];
        } else {
          first = dbin * (uint32_t)a.dst_stride;
          if (gg + v[i] > a.capacity) {
            // The run does not fit: the whole attempt is abandoned (ERRF_CAPACITY, the host redoes the shuffle with
            // the exact layout from the untouched input), so these records only have to land INSIDE the buffer:
            // at the start of the bin's region (the buffers carry one tile of slack behind the last region).
            atomicOr(a.err_flags, (uint32_t)ERRF_CAPACITY);
            gg = 0;
          }
        }
        sbase[b] = first + gg - ex[i];  // output position p of this tile goes to slot sbase[b] + p
      }
    }
    __syncthreads();
    // copy out: consecutive output positions of one bin are consecutive in global memory
#pragma unroll
    for (int k = 0; k < U; k++) {
      const uint32_t p = tid + k * THREADS;
      if (((k + 1) * THREADS <= T || p < (uint32_t)T) && (FULL || p < tn)) {
        const uint32_t pb = perm[p];
        const uint32_t slot = sbase[pb >> 16] + p;
        const uint4* r = raw + (size_t)(pb & 0xffffu) * R::kVec;
        uint4* d = a.dst + (size_t)slot * R::kVec;
#pragma unroll
        for (int vv = 0; vv < R::kVec; vv++) stg_stream(d + vv, r[vv]);
      }
    }
    __syncthreads();  // the raw tile may be overwritten, perm / soff / sbase reused
  };

  if (!DYN) {
    for (uint32_t it = 0; more; it++) {
      const uint32_t tn = cur_len;
      const uint4* raw = (const uint4*)(smem_raw + (it & 1) * TILE_BYTES);
      // the other buffer was last read by the copy-out of the previous iteration (barrier at its end)
      more = advance();
      if (tid == 0 && more) bulk_load(smem_raw + ((it & 1) ^ 1) * TILE_BYTES, cur_p, cur_len * RB, &mbar[(it & 1) ^ 1]);
      mbar_wait(&mbar[it & 1], (it >> 1) & 1);
      if (tn == (uint32_t)T)
        split_tile(std::true_type{}, raw, tn);
      else
        split_tile(std::false_type{}, raw, tn);
    }
  } else {
    // Level 1 of the optimistic layout: ONE stream, and the CTAs do not run at one speed (measured inside a shuffle:
    // with 132 tiles each the first CTA ended after 0.62 ms, the last after 0.80 ms).  After its first tile
    // (blockIdx.x) a CTA takes the next tile nobody has from a ticket counter; thread 0 fetches the ticket one
    // iteration before it needs it, so the L2 round trip is not waited for.
    const uint64_t nrec = s_n[0];
    const uint4* base = (const uint4*)(uintptr_t)s_src[0];
    auto tile_len = [&](uint32_t t) -> uint32_t {
      const uint64_t t0 = (uint64_t)t * T;
      return t0 < nrec ? ((nrec - t0) < (uint64_t)T ? (uint32_t)(nrec - t0) : (uint32_t)T) : 0u;
    };
    uint32_t tn = tile_len(blockIdx.x);
    uint32_t tk = 0;  // (thread 0) the tile to prefetch next
    if (tid == 0) {
      if (tn) bulk_load(smem_raw, base + (uint64_t)blockIdx.x * T * R::kVec, tn * RB, &mbar[0]);
      tk = gridDim.x + atomicAdd(a.ticket, 1u);
    }
    for (uint32_t it = 0; tn; it++) {
      const uint4* raw = (const uint4*)(smem_raw + (it & 1) * TILE_BYTES);
      if (tid == 0) {  // the other buffer was last read by the copy-out of the previous iteration (barrier at its end)
        const uint32_t len = tile_len(tk);
        s_next_len[(it & 1) ^ 1] = len;
        if (len) {
          bulk_load(smem_raw + ((it & 1) ^ 1) * TILE_BYTES, base + (uint64_t)tk * T * R::kVec, len * RB, &mbar[(it & 1) ^ 1]);
          tk = gridDim.x + atomicAdd(a.ticket, 1u);
        }
      }
      mbar_wait(&mbar[it & 1], (it >> 1) & 1);
      if (tn == (uint32_t)T)
        split_tile(std::true_type{}, raw, tn);
      else
        split_tile(std::false_type{}, raw, tn);
      tn = s_next_len[(it & 1) ^ 1];  // (written before the barriers of split_tile)
    }
    // the last CTA to leave resets the counters for the next launch on the stream
    if (tid == 0 && atomicAdd(a.ticket + 1, 1u) == gridDim.x - 1) {
      a.ticket[0] = 0;
      a.ticket[1] = 0;
    }
  }
  if (a.span && tid == 0 && level == 1) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMax(a.span + 1, t);  // last CTA end
    atomicMin(a.span + 3, t);  // first CTA end
  }
}

// single-CTA exclusive scan (B <= a few million): out_excl[0..n], optional copy into
// out_copy (scatter cursors), optional list of entries larger than cap
__device__ __forceinline__ void exscan_body(const uint32_t* __restrict__ in, uint32_t n,
                                            uint32_t* __restrict__ out_excl, uint32_t* __restrict__ out_copy,
                                            uint32_t* __restrict__ out_dense, uint32_t cap,
                                            uint32_t* __restrict__ big_list, uint32_t* nbig, uint32_t* total,
                                            uint32_t shift) {
  __shared__ uint32_t warp_sums[32];
  const uint32_t T = blockDim.x, tid = threadIdx.x;
  uint32_t per = (n + T - 1) / T;
  uint32_t b0 = tid * per, b1 = min(n, b0 + per);
  uint32_t s = 0;
  for (uint32_t i = b0; i < b1; i++) s += in[(size_t)i << shift];
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
    uint32_t c = in[(size_t)i << shift];
    out_excl[i] = run;
    if (out_copy) out_copy[(size_t)i << shift] = run;
    if (out_dense) out_dense[i] = c;
    if (big_list && c > cap) big_list[atomicAdd(nbig, 1u)] = i;
    run += c;
  }
  if (tid == T - 1) {
    out_excl[n] = run;
    if (total) *total = run;
  }
}
__global__ void __launch_bounds__(1024) k_exscan(const uint32_t* __restrict__ in, uint32_t n,
                                                 uint32_t* __restrict__ out_excl,
                                                 uint32_t* __restrict__ out_copy,
                                                 uint32_t* __restrict__ out_dense, uint32_t cap,
                                                 uint32_t* __restrict__ big_list, uint32_t* nbig,
                                                 uint32_t* total, uint32_t shift) {
  exscan_body(in, n, out_excl, out_copy, out_dense, cap, big_list, nbig, total, shift);
}
// Multi-CTA exclusive scan (bins in the tens of thousands made the single-CTA scan a visible slice of
// the shuffle): block b owns elements [1024 b, 1024 b + 1024), one per thread.  Pass 1 writes the block
// sums, pass 2 adds up the sums of the preceding blocks (<= 4096 of them), scans its own elements and
// writes the same outputs as exscan_body.
__global__ void __launch_bounds__(1024) k_exscan_partials(const uint32_t* __restrict__ in, uint32_t n, uint32_t shift,
                                                          uint32_t* __restrict__ partials) {
  __shared__ uint32_t warp_sums[32];
  const uint32_t tid = threadIdx.x, i = blockIdx.x * 1024u + tid;
  uint32_t v = i < n ? in[(size_t)i << shift] : 0u;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((tid & 31) == 0) warp_sums[tid >> 5] = v;
  __syncthreads();
  if (tid < 32) {
    uint32_t w = warp_sums[tid];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(0xffffffffu, w, d);
    if (tid == 0) partials[blockIdx.x] = w;
  }
}
__global__ void __launch_bounds__(1024) k_exscan_apply(const uint32_t* __restrict__ in, uint32_t n,
                                                       uint32_t* __restrict__ out_excl, uint32_t* __restrict__ out_copy,
                                                       uint32_t* __restrict__ out_dense, uint32_t cap,
                                                       uint32_t* __restrict__ big_list, uint32_t* nbig, uint32_t* total,
                                                       uint32_t shift, const uint32_t* __restrict__ partials) {
  __shared__ uint32_t warp_sums[32], warp_pre[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, i = blockIdx.x * 1024u + tid;
  uint32_t pre = 0;  // sum of the preceding blocks
  for (uint32_t b = tid; b < blockIdx.x; b += 1024u) pre += partials[b];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, d);
  const uint32_t c = i < n ? in[(size_t)i << shift] : 0u;
  uint32_t incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (uint32_t)d) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  if (lane == 0) warp_pre[warp] = pre;
  __syncthreads();
  uint32_t w = warp_sums[lane], wi = w, p = warp_pre[lane];
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
    if (lane >= (uint32_t)d) wi += t;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) p += __shfl_xor_sync(0xffffffffu, p, d);
  const uint32_t run = p + __shfl_sync(0xffffffffu, wi - w, warp) + incl - c;  // exclusive prefix of element i
  if (i < n) {
    out_excl[i] = run;
    if (out_copy) out_copy[(size_t)i << shift] = run;
    if (out_dense) out_dense[i] = c;
    if (big_list && c > cap) big_list[atomicAdd(nbig, 1u)] = i;
    if (i == n - 1) {
      out_excl[n] = run + c;
      if (total) *total = run + c;
    }
  }
}
// one CTA per source rank: exclusive scan of that rank's counts for this rank's bins
__global__ void __launch_bounds__(1024) k_exscan_rows(const uint32_t* __restrict__ all, uint32_t stride, uint32_t base,
                                                      uint32_t n, uint32_t* __restrict__ out, uint32_t* __restrict__ totals) {
  uint32_t r = blockIdx.x;
  exscan_body(all + (size_t)r * stride + base, n, out + (size_t)r * (n + 1), nullptr, nullptr, 0xffffffffu, nullptr, nullptr,
              totals + r, 0);
}

// ============================================================================
// map-side combiner (mapreduce/job.lua:92-96,198-202 with the built-in sum)
// ============================================================================
// Two hash tables of whole records.
//  * Every CTA keeps an open-addressing table in SHARED memory for its lifetime: hot keys (Zipf) are summed
//    there with shared-memory atomics and never travel further.
//  * Whatever does not find a place there -- the long tail -- is added to ONE open-addressing table in GLOBAL
//    memory sized to stay resident in the 126 MB L2 (2^21 16-byte entries = 32 MB): one L2 atomic per pair
//    instead of a partition pass, a sort pass and a reduce pass over that pair.  At its end every CTA flushes
//    its shared table into the global one; k_gtab_compact then emits one record per table entry.
// The result (one record per distinct key, in the common case) is what the partition/sort/reduce stages see.
// The global table does not have to be perfect: two entries for one key only mean that the sort+reduce stage
// adds them up, so reads of a slot another thread is publishing need no stronger ordering than "state first".
// A full table raises ERRF_SKEW and the host retries with a larger one (not L2 resident any more) or without
// the combiner; a u32 sum about to wrap raises ERRF_OVERFLOW (string records carry u32 values).
constexpr uint32_t kCombineLock = 0xffffffffu;
constexpr int kCombineThreads = 1024;
constexpr int kCombineSmem = 188 * 1024;  // (+ 38 KB of per-warp miss queues = 226 of the 227 KB)

// cheap 32-bit hash of the key words (one IMAD per word); the shared table takes its top bits, the global
// table a remix of it
template <int RB>
__device__ __forceinline__ uint32_t slot_hash(const uint32_t* w) {
  uint32_t h = 0x9E3779B9u;
#pragma unroll
  for (int k = 0; k < Rec<RB>::kKeyWords; k++) h = (h ^ w[k]) * 0x85EBCA6Bu + (h >> 15);
  h ^= h >> 16;
  h *= 0xC2B2AE35u;
  return h ^ (h >> 13);
}
// L2 residency: the global table (64 MB) only stays in the 126 MB L2 if the 32 GB of pairs streaming through the
// same L2 do not push it out (ncu, before: L2 hit rate 38 %, the table walk waiting ~10 us per round).  The pair
// stream is read with an evict_first policy, every table access carries evict_last.  Table reads are relaxed
// gpu-scope loads (L2 is the point of coherence; `volatile` would compile to system-scope loads).
struct L2Policy {
  uint64_t stream, keep;
};
__device__ __forceinline__ L2Policy l2_policies() {
  L2Policy p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.stream));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.keep));
  return p;
}
__device__ __forceinline__ uint4 ldg_stream_hint(const uint4* p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(p), "r"(bytes), "l"(pol) : "memory");
}
constexpr int kPrefetchTrips = 4;  // (3..6 measure the same) x 32 KB per CTA x 148 CTAs = 19 MB of the stream in L2 ahead of its use; 12 trips
                                    // (57 MB) pushed the 32 MB global table out of L2 again: 10.8 -> 12.9 ms
constexpr uint32_t kEpochTrips = 256;  // shared-table clean-up period: 256 trips = 262,144 pairs per CTA
__device__ __forceinline__ uint32_t ldv_u32(const uint32_t* p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ldv_u64(const unsigned long long* p, uint64_t pol) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ldv_v4(const uint4* p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(pol)
               : "memory");
  return v;
}
__device__ __forceinline__ void red_add_u32(uint32_t* p, uint32_t v, uint64_t pol) {
  asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v, uint64_t pol) {
  asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void stv_v4(uint4* p, const uint4& v, uint64_t pol) {
  asm volatile("st.relaxed.gpu.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w),
               "l"(pol)
               : "memory");
}

template <int RB>
__device__ __forceinline__ void load_rec_hint(const uint4* p, uint32_t* w, uint64_t pol) {
#pragma unroll
  for (int v = 0; v < Rec<RB>::kVec; v++) {
    uint4 x = ldg_stream_hint(p + v, pol);
    w[4 * v + 0] = x.x;
    w[4 * v + 1] = x.y;
    w[4 * v + 2] = x.z;
    w[4 * v + 3] = x.w;
  }
}

// ---- the global table --------------------------------------------------------------------------------------
// 2^glog SHORT entries of 16 bytes, open addressing with linear probing, ONE entry per probe:
//   string keys of up to 12 bytes: {key word 0, 1, 2, u32 state}        u64 keys: {u64 key, u64 state}
// state = 0 empty / all ones being written / sum + 1.  An entry is one aligned 16-byte vector: it is claimed with
// a CAS on the state, written (key + lock) with one vector store, published with an exchange of the state after a
// fence, and read with one vector load -- a reader that sees a published state sees its key.
// String keys longer than 12 bytes (rare in word counts) go to 2^(glog-4) LONG entries of one record slot each
// behind the short ones, walked entry by entry by the whole warp (state first, then the key).
// Why one 16-byte entry per probe: a random L2 read costs the SM per REQUEST, not per byte
// (profiles/microbench/gather_l2.cu: 143 G 16-byte reads/s, 72 G 32-byte, 36 G 64-byte reads/s on the chip), and
// a probe is never waited for: it is issued when 32 pairs have queued up and looked at when the next 32 have.
// CHECKED: use the returned old value to catch a u32 sum about to wrap (otherwise the add is fire-and-forget and
// the caller has bounded the sums: pairs x largest value < 2^32).
__host__ __device__ inline uint64_t gtab_bytes(int rb, uint32_t glog) {
  return ((uint64_t)16 << glog) + (rb == 16 ? 0ull : ((uint64_t)rb << (glog - 4)));
}
constexpr uint32_t kGtabMaxProbes = 192;  // linear probes before the table counts as full
constexpr int kGtabMaxLongProbes = 128;

// A pair on its way to the global table, 16 bytes: strings of up to 12 bytes {key word 0, 1, 2, u32 value},
// u64 keys {key lo, key hi, value lo, value hi} -- also the layout of a short table entry (value -> state).
template <int RB>
__device__ __forceinline__ uint32_t gtab_slot(const uint4& E, uint32_t glog) {
  uint32_t g = (E.x ^ 0x9E3779B9u) * 0x85EBCA6Bu;
  g = (g ^ (g >> 15) ^ E.y) * 0x2C1B3C6Du;
  if (!Rec<RB>::kU64) g = (g ^ (g >> 13) ^ E.z) * 0x297A2D39u;
  g ^= g >> 16;
  g *= 0xC2B2AE35u;
  g ^= g >> 15;
  return g >> (32 - glog);
}
// One probe of the short table: entry `e` as read into x.  Returns 0 = the pair is settled (added, or inserted into
// the empty entry), 1 = the entry holds another key (probe the next one), 2 = the entry is being written (look again).
template <int RB, bool CHECKED>
__device__ __forceinline__ int gtab_probe(uint4* e, const uint4& x, const uint4& E, uint32_t* __restrict__ flags, uint64_t pol) {
  if constexpr (Rec<RB>::kU64) {
    unsigned long long* st_p = (unsigned long long*)e + 1;
    const unsigned long long st = (unsigned long long)x.z | ((unsigned long long)x.w << 32), lock = ~0ull;
    const unsigned long long v = (unsigned long long)E.z | ((unsigned long long)E.w << 32);
    if (st == 0) {
      if (atomicCAS(st_p, 0ull, lock) != 0ull) return 2;
      stv_v4(e, make_uint4(E.x, E.y, 0xffffffffu, 0xffffffffu), pol);
      __threadfence();
      atomicExch(st_p, v + 1ull);
      return 0;
    }
    if (st == lock) return 2;
    if (x.x != E.x || x.y != E.y) return 1;
    if (v) red_add_u64(st_p, v, pol);
    return 0;
  } else {
    uint32_t* st_p = (uint32_t*)e + 3;
    const uint32_t v32 = E.w;
    if (x.w == 0) {
      if (atomicCAS(st_p, 0u, kCombineLock) != 0u) return 2;
      stv_v4(e, make_uint4(E.x, E.y, E.z, kCombineLock), pol);
      __threadfence();
      atomicExch(st_p, v32 + 1u);
      return 0;
    }
    if (x.w == kCombineLock) return 2;
    if (x.x != E.x || x.y != E.y || x.z != E.z) return 1;
    if (v32) {
      if (CHECKED) {
        const uint32_t old = atomicAdd(st_p, v32);
        if (old + v32 < old || old + v32 >= 0xfffffff0u) atomicOr(flags, (uint32_t)ERRF_OVERFLOW);
      } else {
        red_add_u32(st_p, v32, pol);
      }
    }
    return 0;
  }
}
// WARP-COLLECTIVE: string keys longer than 12 bytes, one record slot per entry, state first, then the key
template <int RB, bool CHECKED>
__device__ __forceinline__ void gtab_add_long(bool active, uint32_t* __restrict__ gtab, uint32_t glog, const uint32_t* w, uint32_t v32,
                                              uint32_t* __restrict__ flags, uint64_t pol) {
  using R = Rec<RB>;
  constexpr int W = R::kWords, KW = R::kKeyWords;
  uint32_t* ltab = gtab + ((size_t)4 << glog);
  const uint32_t lmask = (1u << (glog - 4)) - 1u;
  uint32_t slot = slot_hash<RB>(w) >> (32 - (glog - 4));
  bool ldone = !active;
  if (!ldone && v32 >= 0xfffffff0u) {
    atomicOr(flags, (uint32_t)ERRF_OVERFLOW);
    ldone = true;
  }
#pragma unroll 1
  for (int probe = 0; probe < kGtabMaxLongProbes; probe++) {
    if (__all_sync(0xffffffffu, ldone)) break;
    if (!ldone) {
      uint32_t* e = ltab + (size_t)slot * W;
      uint32_t st = ldv_u32(e + KW, pol);
      if (st == 0) {
        if (atomicCAS(e + KW, 0u, kCombineLock) == 0u) {
#pragma unroll
          for (int i = 0; i < R::kVec; i++)  // the last vector rewrites the lock word with itself
            stv_v4((uint4*)e + i, make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], 4 * i + 3 == KW ? kCombineLock : w[4 * i + 3]), pol);
          __threadfence();
          atomicExch(e + KW, v32 + 1u);
          ldone = true;
        }
        st = kCombineLock;
      }
      if (!ldone) {
        while (st == kCombineLock) st = ldv_u32(e + KW, pol);
        bool eq = true;
#pragma unroll
        for (int i = 0; i < R::kVec; i++) {
          const uint4 y = ldv_v4((const uint4*)e + i, pol);
          eq = eq && y.x == w[4 * i] && y.y == w[4 * i + 1] && y.z == w[4 * i + 2] && (4 * i + 3 == KW || y.w == w[4 * i + 3]);
        }
        if (eq) {
          if (v32) {
            if (CHECKED) {
              const uint32_t old = atomicAdd(e + KW, v32);
              if (old + v32 < old || old + v32 >= 0xfffffff0u) atomicOr(flags, (uint32_t)ERRF_OVERFLOW);
            } else {
              red_add_u32(e + KW, v32, pol);
            }
          }
          ldone = true;
        }
      }
      slot = (slot + 1) & lmask;
    }
  }
  if (!ldone) atomicOr(flags, (uint32_t)ERRF_SKEW);
}

// flags[0] |= ERRF_*, flags[2] = max over the values seen (saturated to u32).
// vcap: largest value the shared table accepts -- chosen by the host so that one CTA's sum of such values cannot
// wrap 32 bits (pairs per CTA x vcap < 2^32), which keeps saturation tests out of the per-pair path.
template <int RB, bool CHECKED>
__global__ void __launch_bounds__(kCombineThreads, 1)
    k_combine(const uint4* __restrict__ recs, uint64_t n, uint32_t pf_trips, uint32_t vcap, uint32_t* __restrict__ gtab,
              uint32_t glog, uint32_t* __restrict__ flags, uint32_t tune) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using R = Rec<RB>;
  constexpr int W = R::kWords, KW = R::kKeyWords;
  if ((tune & 64u) && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMin((unsigned long long*)(flags + 8), t);
  }
  constexpr int KC = KW < 3 ? KW : 3;  // key words of a shared-table entry: only keys of up to 12 bytes live there
  // Shared table, ARRAY PER FIELD (32 lanes probing 32 random entries hit 32 random banks, not the four a record
  // stride allows): tag[e] = 0 empty / 1 being written / the key's 32-bit hash with bit 1 set; val[e] = u32 partial
  // sum; key[k][e] = key word k < KC.  A probe reads two neighbouring tags; only a tag hit touches the key words.
  // Longer keys (rare in a word count) bypass it: 20 bytes per entry instead of 36 buys 1.8 times the entries.
  // Every kEpochTrips trips the CTA stops at a barrier and evicts the entries that gathered less than two pairs
  // since they were admitted (into the global table, like a miss): the table is filled first come first served,
  // and without that the keys that happened to arrive first would keep out warmer keys that arrived later.
  // (compile-time sizes: table addresses become immediates -- the kernel is issue bound and short of registers)
  constexpr uint32_t entries = (uint32_t)(kCombineSmem / (R::kU64 ? 16 : 20));  // tag + value + 2 (u64) or 3 key words
  uint32_t* tag = (uint32_t*)smem_raw;
  uint32_t* val = tag + entries;
  uint32_t* key = val + entries;
  // Pairs that found no place in the shared table queue up per warp as 16-byte compact entries (a ring of 64, with
  // the number of the probe they are at).  When 32 are there the warp ISSUES their probes -- one 16-byte read per
  // lane, into registers -- and goes on with the stream; it LOOKS at them when the next 32 are ready: a match is
  // added to with a fire-and-forget red, an empty entry is claimed, anything else (another key: next probe; an
  // entry being written: same probe) goes back into the ring.  Nothing ever waits for the table: the L2 round trip,
  // microseconds with 32 warps per SM and nothing else to run, is spent on the stream (ncu, synchronous walk:
  // 20-35 % of the stall samples waited for these reads, the walk taking as long as its unluckiest lane).
  __shared__ uint4 queue[kCombineThreads / 32][64];
  __shared__ uint8_t queue_probe[kCombineThreads / 32][64];
  __shared__ uint32_t long_queue[kCombineThreads / 32][32];  // indices of pairs with long keys, walked 32 at a time
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  const L2Policy pol = l2_policies();
  uint4* q = queue[tid >> 5];
  uint8_t* qp = queue_probe[tid >> 5];
  uint32_t* lq = long_queue[tid >> 5];
  uint32_t lq_n = 0;
  uint32_t q_tail = 0, q_pend = 0;  // ring: [.. in flight ..][.. q_pend pending ..] tail
  uint32_t fly_base = 0, fly_n = 0; // the batch in flight: fly_n entries from fly_base (0 = none)
  uint4 fx;                         // ... this lane's probed table entry
  uint32_t vmax = 0;
  for (uint32_t i = tid; i < entries * (KC + 2); i += blockDim.x) tag[i] = 0;
  __syncthreads();
  const uint32_t gmask = (1u << glog) - 1u;
  auto append = [&](bool p, const uint4& E, uint32_t probe) {  // collective: the entries of the lanes with p go to the tail
    const uint32_t m = __ballot_sync(0xffffffffu, p);
    if (p) {
      const uint32_t at = (q_tail + __popc(m & ((1u << lane) - 1u))) & 63u;
      q[at] = E;
      qp[at] = (uint8_t)probe;
    }
    q_tail = (q_tail + __popc(m)) & 63u;
    q_pend += __popc(m);
    __syncwarp();
  };
  auto look = [&]() {  // collective: the batch in flight
    const bool mine = lane < fly_n;
    uint4 E = make_uint4(0, 0, 0, 0);
    uint32_t probe = 0;
    int r = 0;
    if (mine) {
      E = q[(fly_base + lane) & 63u];
      probe = qp[(fly_base + lane) & 63u];
      const uint32_t slot = (gtab_slot<RB>(E, glog) + probe) & gmask;
      r = gtab_probe<RB, CHECKED>((uint4*)gtab + slot, fx, E, flags, pol.keep);
      if (r == 1 && ++probe >= kGtabMaxProbes) {
        atomicOr(flags, (uint32_t)ERRF_SKEW);  // table full: the host retries with a larger one
        r = 0;
      }
    }
    fly_n = 0;
    __syncwarp();
    append(r != 0, E, probe);  // (never more than the batch just freed: the ring cannot overflow here)
  };
  auto issue = [&]() {  // collective: the oldest min(32, q_pend) pending entries
    fly_n = q_pend < 32u ? q_pend : 32u;
    fly_base = (q_tail - q_pend) & 63u;
    if (lane < fly_n) {
      const uint4 E = q[(fly_base + lane) & 63u];
      const uint32_t slot = (gtab_slot<RB>(E, glog) + qp[(fly_base + lane) & 63u]) & gmask;
      fx = ldv_v4((const uint4*)gtab + slot, pol.keep);
    }
    q_pend -= fly_n;
  };
  auto push = [&](bool p, const uint4& E) {  // collective: new entries; makes room first, starts a batch when 32 wait
    const uint32_t k = __popc(__ballot_sync(0xffffffffu, p));
    while (fly_n + q_pend + k > 64) {  // (rare: the ring holds the batch in flight + what came back + what is new)
      if (fly_n) look();
      else issue();
    }
    append(p, E, 0u);
    if (q_pend >= 32) {
      if (fly_n) look();
      if (q_pend >= 32) issue();
    }
  };
  auto long_walk = [&]() {  // collective
    uint32_t w[W];
#pragma unroll
    for (int k = 0; k < W; k++) w[k] = 0;
    const bool mine = lane < lq_n;
    if (mine) load_rec_hint<RB>(recs + (size_t)lq[lane] * R::kVec, w, pol.stream);
    gtab_add_long<RB, CHECKED>(mine, gtab, glog, w, (uint32_t)rec_value<RB>(w), flags, pol.keep);
    lq_n = 0;
    __syncwarp();
  };
  // (between two barriers) entries with a sum below `below` leave the shared table for the global one
  auto evict = [&](uint32_t below) {
    const uint32_t e_round = (entries + 31) / 32 * 32;
    for (uint32_t e = tid; e < e_round; e += blockDim.x) {
      const bool out = e < entries && tag[e] > 1u && val[e] < below;
      uint4 E = make_uint4(0, 0, 0, 0);
      if (out) {
        const uint32_t k0 = key[e], k1 = key[entries + e], k2 = KC > 2 ? key[2u * entries + e] : 0u;
        E = R::kU64 ? make_uint4(k0, k1, val[e], 0u) : make_uint4(k0, k1, k2, val[e]);
        tag[e] = 0;
      }
      if (__any_sync(0xffffffffu, out)) push(out, E);
    }
  };
  // every CTA streams ONE contiguous slice of the pairs
  const uint64_t slice = ((n + gridDim.x - 1) / gridDim.x + 31) / 32 * 32;
  const uint64_t slice_lo = slice * blockIdx.x < n ? slice * blockIdx.x : n;
  const uint64_t slice_hi = slice_lo + slice < n ? slice_lo + slice : n;  // (this CTA's pairs: [slice_lo, slice_hi))
  const uint32_t ntrips = (uint32_t)((slice_hi - slice_lo + blockDim.x - 1) / blockDim.x);  // (the same for every thread: barriers inside)
  // the pair of the NEXT trip is requested before this one is processed.  Index arithmetic of the loop: running
  // 64-bit pointers and 32-bit trip bounds (the kernel is issue bound; the per-trip 64-bit index, bounds checks and
  // prefetch address were 15 % of its instructions)
  const uint32_t span = (uint32_t)(slice_hi - slice_lo);                                    // pairs of this CTA (< 2^32)
  const uint32_t my_trips = tid < span ? (span - tid + blockDim.x - 1) / blockDim.x : 0u;   // trips in which this thread has a pair
  // the warp's 32 pairs of trip t + pf_trips lie inside the slice for t < pf_ok
  const uint32_t wbase = (tid & ~31u) + 32u;
  const uint32_t pf_full = span >= wbase ? (span - wbase) / blockDim.x + 1u : 0u;          // trips whose 32 pairs are all there
  const uint32_t pf_ok = pf_full > pf_trips ? pf_full - pf_trips : 0u;
  const uint4* p_next = recs + (slice_lo + tid) * R::kVec;                                  // this thread's pair of the next trip
  const uint4* p_pf = recs + (slice_lo + (tid & ~31u) + (uint64_t)pf_trips * blockDim.x) * R::kVec;  // the warp's prefetch target
  const size_t p_step = (size_t)blockDim.x * R::kVec;
  uint32_t i32 = (uint32_t)slice_lo + tid;  // index of the pair in hand (all pairs of one shuffle: < 2^32)
  uint32_t wn[W];
  if (my_trips) load_rec_hint<RB>(p_next, wn, pol.stream);
  for (uint32_t trip = 0; trip < ntrips; trip++, i32 += blockDim.x) {
    if (trip % kEpochTrips == kEpochTrips - 1) {
      __syncthreads();
      evict(2u);
      __syncthreads();
    }
    uint32_t w[W];
#pragma unroll
    for (int k = 0; k < W; k++) w[k] = wn[k];
    p_next += p_step;
    if (trip + 1 < my_trips) load_rec_hint<RB>(p_next, wn, pol.stream);
    // ... and the warp's 32 pairs of kPrefetchTrips trips ahead are pulled from DRAM into L2 by the bulk-copy engine:
    // one trip of work (~0.4 us) does not cover a DRAM access under load (1.5-2 us), an L2 hit it does.  (One 32 KB
    // request per CTA by a single thread instead of 1 KB per warp saves 10 % of the instructions and LOSES 25 %:
    // 10.8 -> 13.6 ms, the warps of a CTA are not in step.)
    if (lane == 0 && trip < pf_ok) bulk_prefetch_l2(p_pf, 32u * RB, pol.stream);
    p_pf += p_step;
    bool need = trip < my_trips;
    const uint64_t v = need ? rec_value<RB>(w) : 0ull;
    vmax = max(vmax, (uint32_t)(v > 0xffffffffull ? 0xffffffffull : v));
    // Straight-line, predicated code: a probe loop that lanes leave at different trips falls apart into fragments
    // that do not reconverge inside it, and every fragment then waits for shared memory on its own.
    bool is_short = true;
    if (!R::kU64) {
#pragma unroll
      for (int k = 3; k < KW; k++) is_short = is_short && w[k] == 0;
    }
    if (need && is_short && v != 0 && v <= vcap) {  // (a zero would still have to create its key: left to the global table)
      uint32_t h = 0x9E3779B9u;  // slot_hash over the KC words a short key has
#pragma unroll
      for (int k = 0; k < KC; k++) h = (h ^ w[k]) * 0x85EBCA6Bu + (h >> 15);
      h ^= h >> 16;
      h *= 0xC2B2AE35u;
      h ^= h >> 13;
      const uint32_t tg = h | 2u;
      const uint32_t s0 = __umulhi(h, entries), s1 = s0 + 1 == entries ? 0u : s0 + 1;
      const uint32_t t0 = ((volatile uint32_t*)tag)[s0], t1 = ((volatile uint32_t*)tag)[s1];
      if (t0 == tg || t1 == tg) {
        const uint32_t sl = t0 == tg ? s0 : s1;
        uint32_t x[KC];
#pragma unroll
        for (int k = 0; k < KC; k++) x[k] = ((volatile uint32_t*)key)[(uint32_t)k * entries + sl];
        bool eq = true;
#pragma unroll
        for (int k = 0; k < KC; k++) eq = eq && x[k] == w[k];
        if (eq) {
          atomicAdd(val + sl, (uint32_t)v);
          need = false;
        }  // (same hash, other key: the pair goes to the global table, which is allowed to hold a key twice)
      } else if (t0 == 0 || t1 == 0) {  // claim the empty slot (the table fills early in the kernel's life; then rare)
        const uint32_t sl = t0 == 0 ? s0 : s1;
        if (atomicCAS(tag + sl, 0u, 1u) == 0u) {
#pragma unroll
          for (int k = 0; k < KC; k++) key[(uint32_t)k * entries + sl] = w[k];
          val[sl] = (uint32_t)v;
          __threadfence_block();
          atomicExch(tag + sl, tg);
          need = false;
        }
      }
    }
    // misses: long string keys (rare) are remembered by index and walk their table 32 at a time, all lanes busy;
    // everything else queues up as a compact entry
    if (!R::kU64) {
      if (need && v >= 0xfffffff0ull) {  // (its state word, value + 1, would read "being written")
        atomicOr(flags, (uint32_t)ERRF_OVERFLOW);
        need = false;
      }
      const uint32_t ml = __ballot_sync(0xffffffffu, need && !is_short);
      if (ml) {
        if (lq_n + __popc(ml) > 32) long_walk();
        if (need && !is_short) lq[lq_n + __popc(ml & ((1u << lane) - 1u))] = i32;
        lq_n += __popc(ml);
        __syncwarp();
      }
    }
    const bool p = need && is_short;
    if (__any_sync(0xffffffffu, p)) push(p, R::kU64 ? make_uint4(w[0], w[1], w[2], w[3]) : make_uint4(w[0], w[1], w[2], (uint32_t)v));
  }
  __syncthreads();
  evict(0xffffffffu);  // the shared table's entries take the same road as the misses
  if (!R::kU64 && lq_n) long_walk();
  // drain: whatever is in flight or pending, batch by batch, until nothing comes back
  while (fly_n || q_pend) {
    if (fly_n) look();
    if (q_pend) issue();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  if (lane == 0 && vmax) atomicMax(flags + 2, vmax);
  if ((tune & 64u) && threadIdx.x == 0) {  // measurement hook: when does the first CTA end, when the last?
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMax((unsigned long long*)(flags + 8) + 1, t);
    atomicMin((unsigned long long*)(flags + 8) + 2, t);
  }
}

// one record per non-empty entry of the global table (short entries first, then the long ones), appended to out
template <int RB>
__global__ void __launch_bounds__(256) k_gtab_compact(const uint32_t* __restrict__ gtab, uint32_t glog, uint4* __restrict__ out,
                                                      uint32_t* __restrict__ count) {
  using R = Rec<RB>;
  constexpr int W = R::kWords, KW = R::kKeyWords;
  const uint32_t nshort = 1u << glog, nlong = R::kU64 ? 0u : 1u << (glog - 4), lane = threadIdx.x & 31;
  const uint32_t* ltab = gtab + ((size_t)4 << glog);
  for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < nshort + nlong; base += gridDim.x * blockDim.x) {
    const uint32_t s = base + lane;  // (nshort and nlong are multiples of 32: a warp never straddles the two tables)
    uint32_t w[W];
#pragma unroll
    for (int k = 0; k < W; k++) w[k] = 0;
    bool has;
    if (s < nshort) {
      const uint4 x = ((const uint4*)gtab)[s];
      if constexpr (R::kU64) {
        unsigned long long st = (unsigned long long)x.z | ((unsigned long long)x.w << 32);
        has = st != 0;
        st -= 1ull;
        w[0] = x.x;
        w[1] = x.y;
        w[2] = (uint32_t)st;
        w[3] = (uint32_t)(st >> 32);
      } else {
        has = x.w != 0;
        w[0] = x.x;
        w[1] = x.y;
        w[2] = x.z;
        w[KW] = x.w - 1u;
      }
    } else {
      const uint4* e = (const uint4*)(ltab + (size_t)(s - nshort) * W);
#pragma unroll
      for (int i = 0; i < R::kVec; i++) {
        const uint4 x = e[i];
        w[4 * i] = x.x;
        w[4 * i + 1] = x.y;
        w[4 * i + 2] = x.z;
        w[4 * i + 3] = x.w;
      }
      has = w[KW] != 0;
      w[KW] -= 1u;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, has);
    if (!m) continue;
    uint32_t pos = 0;
    if (lane == (uint32_t)(__ffs(m) - 1)) pos = atomicAdd(count, (uint32_t)__popc(m));
    pos = __shfl_sync(0xffffffffu, pos, __ffs(m) - 1) + __popc(m & ((1u << lane) - 1u));
    if (has) {
      uint4* d = out + (size_t)pos * R::kVec;
#pragma unroll
      for (int i = 0; i < R::kVec; i++) d[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
  }
}

// pairs a rank reduces = the fill levels (clamped to the region capacity) of its regions in every rank's
// level-1 cursors; acc2[0] = all of them, acc2[1] = the ones that were already local (source == me)
__global__ void __launch_bounds__(256) k_region_totals(const uint32_t* __restrict__ counts, uint64_t zstride, uint32_t G,
                                                       uint32_t me, uint32_t first, uint32_t nreg, uint32_t shift,
                                                       uint32_t clamp, unsigned long long* __restrict__ acc2) {
  __shared__ unsigned long long s_tot[8], s_own[8];
  unsigned long long tot = 0, own = 0;
  for (uint64_t i = threadIdx.x; i < (uint64_t)nreg * G; i += blockDim.x) {
    const uint32_t z = (uint32_t)(i / nreg), r = (uint32_t)(i % nreg);
    uint32_t c = counts[(size_t)z * zstride + ((size_t)(first + r) << shift)];
    c = c < clamp ? c : clamp;
    tot += c;
    if (z == me) own += c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
    own += __shfl_xor_sync(0xffffffffu, own, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_tot[threadIdx.x >> 5] = tot;
    s_own[threadIdx.x >> 5] = own;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tot = own = 0;
    for (int w = 0; w < 8; w++) {
      tot += s_tot[w];
      own += s_own[w];
    }
    acc2[0] = tot;
    acc2[1] = own;
  }
}

// global per-bin totals from the all-gathered counts: this rank's bins go to tot[], and every
// rank counts the bins (of ALL ranks) above cap, so that all ranks take the same decision
__global__ void k_sum_src(const uint32_t* __restrict__ all, uint32_t world, uint32_t stride, uint32_t base,
                          uint32_t n, uint32_t* __restrict__ tot, uint32_t cap, uint32_t* __restrict__ nover) {
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < stride; b += gridDim.x * blockDim.x) {
    uint32_t t = 0;
    for (uint32_t s = 0; s < world; s++) t += all[(size_t)s * stride + b];
    if (b >= base && b < base + n) tot[b - base] = t;
    if (t > cap) atomicAdd(nover, 1u);
  }
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
    uint64_t so = out_start(b, bin), d0 = b.uoff[bin];
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
                                                      uint32_t bin_base, unsigned long long* acc) {
  constexpr int KW = Rec<RB>::kKeyWords;
  uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t* keys = (const uint32_t*)b.out_keys;
  uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, bad_order = 0, bad_part = 0;
  for (uint32_t bin = warp; bin < B; bin += nwarps) {
    uint32_t len = b.ucount[bin];
    uint64_t so = out_start(b, bin);
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
        int oc = key_cmp<RB>(p, w);
        if (oc > 0 || (oc == 0 && !b.no_reduce)) bad_order++;
      }
      uint32_t pid;
      uint32_t mybin = bin_of<RB>(w, bp, &pid);
      if (mybin != bin + bin_base) bad_part++;
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

void kernels_set_tune(uint32_t bits) { g_tune = bits; }

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
  if (e != cudaSuccess) return e;                                                                 \

  CFG(16) CFG(32) CFG(64) CFG(128)
#undef CFG
#define CFGS(M, X)                                                                                                       \
  e = cudaFuncSetAttribute(k_sort_reduce_u64<M, X>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem_bytes(16)); \
  if (e != cudaSuccess) return e;
  CFGS(false, false) CFGS(false, true) CFGS(true, false) CFGS(true, true)
#undef CFGS
#define CFGC1(RB, CH)                                                                                      \
  e = cudaFuncSetAttribute(k_combine<RB, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCombineSmem);  \
  if (e != cudaSuccess) return e;
#define CFGC(RB) CFGC1(RB, false) CFGC1(RB, true)
  CFGC(16) CFGC(32) CFGC(64) CFGC(128)
#undef CFGC1
#undef CFGC
#define CFGT1(RB, SPEC)                                                                                  \
  e = cudaFuncSetAttribute(k_split_tma<RB, kTmaTileBytes, 2, kTmaSplitThreads, SPEC>,                    \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tma_split_smem(RB));        \
  if (e != cudaSuccess) return e;
#define CFGT(RB) CFGT1(RB, 0) CFGT1(RB, 1) CFGT1(RB, 2) CFGT1(RB, 5) CFGT1(RB, 6)
  CFGT(16) CFGT(32) CFGT(64) CFGT(128)
#undef CFGT
#undef CFGT1
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
int launch_hist(int rb, const void* recs, uint64_t n, const BinParams& bp, uint32_t* hist, cudaStream_t s) {
  if (!n) return 0;
  DISPATCH_RB(rb, (k_hist<RB><<<stream_grid(n, 256, 8), 256, 0, s>>>((const uint4*)recs, n, bp, hist)));
  return 1;
}
int launch_exscan(const uint32_t* in, uint32_t n, uint32_t* out_excl, uint32_t* out_copy, uint32_t* out_dense,
                  uint32_t cap, uint32_t* big_list, uint32_t* nbig, uint32_t* total, uint32_t shift,
                  cudaStream_t s, uint32_t* scratch) {
  const uint32_t nb = (n + 1023u) / 1024u;
  if (scratch && nb >= 4 && nb <= kScanScratchWords) {
    k_exscan_partials<<<nb, 1024, 0, s>>>(in, n, shift, scratch);
    k_exscan_apply<<<nb, 1024, 0, s>>>(in, n, out_excl, out_copy, out_dense, cap, big_list, nbig, total, shift, scratch);
    return 2;
  }
  k_exscan<<<1, 1024, 0, s>>>(in, n, out_excl, out_copy, out_dense, cap, big_list, nbig, total, shift);
  return 1;
}
int launch_tok_count(const unsigned char* text, uint64_t len, uint32_t* block_counts, cudaStream_t s) {
  if (!len) return 0;
  uint64_t nb = (len + 255) / 256;
  k_tok_count<<<(int)std::min<uint64_t>(nb, (uint64_t)g_sm_count * 16), 256, 0, s>>>(text, len, block_counts);
  return 1;
}
int launch_tok_emit(int rb, const unsigned char* text, uint64_t len, const uint32_t* block_off, void* recs,
                    uint32_t* flags, cudaStream_t s) {
  if (!len || rb == 16) return 0;
  uint64_t nb = (len + 255) / 256;
  int grid = (int)std::min<uint64_t>(nb, (uint64_t)g_sm_count * 16);
  switch (rb) {
    case 32: k_tok_emit<32><<<grid, 256, 0, s>>>(text, len, block_off, (uint4*)recs, flags); break;
    case 64: k_tok_emit<64><<<grid, 256, 0, s>>>(text, len, block_off, (uint4*)recs, flags); break;
    default: k_tok_emit<128><<<grid, 256, 0, s>>>(text, len, block_off, (uint4*)recs, flags); break;
  }
  return 1;
}
uint64_t gtab_bytes_host(int rb, uint32_t glog) { return gtab_bytes(rb, glog); }
int launch_combine(int rb, const void* recs, uint64_t n, uint32_t* gtab, uint32_t glog, uint32_t* flags, bool checked,
                   int sm_count, cudaStream_t s) {
  if (!n) return 0;
  const uint32_t pf_trips = (g_tune >> 8) & 0xffu ? (g_tune >> 8) & 0xffu : (uint32_t)kPrefetchTrips;  // (MRHBM_TUNE bits 8-15: override)
  // the shared table only takes values whose per-CTA sum cannot wrap 32 bits
  const uint64_t per_cta = (n + sm_count - 1) / sm_count + kCombineThreads;
  const uint32_t vcap = (uint32_t)std::min<uint64_t>(0xffffull, 0xffffffefull / per_cta);
  if (checked) {
    DISPATCH_RB(rb, (k_combine<RB, true><<<sm_count, kCombineThreads, kCombineSmem, s>>>((const uint4*)recs, n, pf_trips, vcap, gtab, glog, flags, g_tune)));
  } else {
    DISPATCH_RB(rb, (k_combine<RB, false><<<sm_count, kCombineThreads, kCombineSmem, s>>>((const uint4*)recs, n, pf_trips, vcap, gtab, glog, flags, g_tune)));
  }
  return 1;
}
int launch_gtab_compact(int rb, const uint32_t* gtab, uint32_t glog, void* out, uint32_t* count, cudaStream_t s) {
  const uint32_t slots = (1u << glog) + (rb == 16 ? 0u : 1u << (glog - 4));
  int grid = (int)std::min<uint32_t>((slots + 255) / 256, (uint32_t)g_sm_count * 8);
  DISPATCH_RB(rb, (k_gtab_compact<RB><<<grid, 256, 0, s>>>(gtab, glog, (uint4*)out, count)));
  return 1;
}
int launch_sum_src(const uint32_t* all, uint32_t world, uint32_t stride, uint32_t base, uint32_t n,
                   uint32_t* tot, uint32_t cap, uint32_t* nover, cudaStream_t s) {
  if (!stride) return 0;
  k_sum_src<<<(stride + 255) / 256 < 1024 ? (stride + 255) / 256 : 1024, 256, 0, s>>>(all, world, stride, base, n, tot,
                                                                                      cap, nover);
  return 1;
}
int launch_region_totals(const uint32_t* counts, uint64_t zstride, uint32_t G, uint32_t me, uint32_t first, uint32_t nreg,
                         uint32_t shift, uint32_t clamp, uint64_t* acc2, cudaStream_t s) {
  k_region_totals<<<1, 256, 0, s>>>(counts, zstride, G, me, first, nreg, shift, clamp, (unsigned long long*)acc2);
  return 1;
}
int launch_exscan_rows(const uint32_t* all, uint32_t world, uint32_t stride, uint32_t base, uint32_t n, uint32_t* out,
                       uint32_t* totals, cudaStream_t s) {
  k_exscan_rows<<<world, 1024, 0, s>>>(all, stride, base, n, out, totals);
  return 1;
}
int launch_sample_u64(const void* recs, uint64_t n, uint32_t nsample, uint32_t* hist256, cudaStream_t s) {
  if (!n || !nsample) return 0;
  k_sample_u64<<<(nsample + 255) / 256 < 256 ? (nsample + 255) / 256 : 256, 256, 0, s>>>((const uint4*)recs, n, nsample, hist256);
  return 1;
}
int launch_scatter(int rb, const void* recs, uint64_t n, const BinParams& bp, uint32_t* cursor, void* mid,
                   cudaStream_t s) {
  if (!n) return 0;
  DISPATCH_RB(rb, (k_scatter<RB><<<stream_grid(n, 256, 8), 256, 0, s>>>((const uint4*)recs, n, bp, cursor, (uint4*)mid)));
  return 1;
}
// level 1 (source -> coarse regions, possibly on peers) and level 2 (coarse regions -> fine bins) of the split
static SplitArgs split_args(const BinParams& bp, const SplitPlan& pl) {
  SplitArgs a{};
  a.base_off = pl.base_off;
  a.rep_shift = bp.rep_shift;
  a.B = pl.B;
  a.F = pl.F;
  a.logF = 0;
  while ((1u << a.logF) < pl.F) a.logF++;
  a.ctr_shift = bp.ctr_shift;
  a.err_flags = pl.err_flags;
  a.span = pl.span;
  a.ticket = pl.ticket;
  a.ndest = pl.base_off ? 1u : pl.ndest;
  a.me = pl.base_off ? 0u : pl.me;
  for (int d = 0; d < 8; d++) a.peer[d] = pl.base_off ? 0ull : pl.peer[d];
  for (int d = 0; d <= 8; d++) {
    a.rbase[d] = pl.base_off ? (d ? pl.C1 : 0u) : pl.rbase[d];
    a.fbase[d] = pl.base_off ? (d ? pl.B : 0u) : pl.fbase[d];
  }
  return a;
}
// the instantiation that has the launch-time facts folded in (k_split_tma's SPEC)
template <int RB>
static void launch_split_rb(dim3 grid, const SplitArgs& a, const BinParams& bp, cudaStream_t s) {
  const int spec = a.base_off ? 0 : (int)a.level | (a.ndest > 1 ? 4 : 0);
  const size_t smem = tma_split_smem(RB);
  switch (spec) {
    case 1: k_split_tma<RB, kTmaTileBytes, 2, kTmaSplitThreads, 1><<<grid, kTmaSplitThreads, smem, s>>>(a, bp); break;
    case 2: k_split_tma<RB, kTmaTileBytes, 2, kTmaSplitThreads, 2><<<grid, kTmaSplitThreads, smem, s>>>(a, bp); break;
    case 5: k_split_tma<RB, kTmaTileBytes, 2, kTmaSplitThreads, 5><<<grid, kTmaSplitThreads, smem, s>>>(a, bp); break;
    case 6: k_split_tma<RB, kTmaTileBytes, 2, kTmaSplitThreads, 6><<<grid, kTmaSplitThreads, smem, s>>>(a, bp); break;
    default: k_split_tma<RB, kTmaTileBytes, 2, kTmaSplitThreads, 0><<<grid, kTmaSplitThreads, smem, s>>>(a, bp); break;
  }
}
static void launch_split(int rb, dim3 grid, const SplitArgs& a, const BinParams& bp, cudaStream_t s) {
  DISPATCH_RB(rb, (launch_split_rb<RB>(grid, a, bp, s)));
}
int launch_split_l1(int rb, const void* recs, uint64_t n, const BinParams& bp, const SplitPlan& pl, cudaStream_t s) {
  if (!n) return 0;
  SplitArgs a = split_args(bp, pl);
  const int ctas = 2 * g_sm_count;
  a.src = (const uint4*)recs;
  a.n = n;
  a.dst = (uint4*)pl.l1;
  a.dst_stride = pl.sub_stride;
  a.cursor = pl.cursor1;
  a.capacity = (uint32_t)std::min<uint64_t>(pl.sub_stride, 0xffffffffull);
  a.nbins = pl.C1;
  a.level = 1;
  launch_split(rb, dim3(ctas), a, bp, s);
  return 1;
}
int launch_split_l2(int rb, const BinParams& bp, const SplitPlan& pl, cudaStream_t s) {
  SplitArgs a = split_args(bp, pl);
  const int ctas = 2 * g_sm_count;
  const uint32_t nsub = pl.base_off ? 1u : pl.ndest;
  const uint32_t regions = pl.base_off ? pl.C1 : pl.C1_local;
  if (!regions) return 0;
  a.src = (const uint4*)pl.l1;
  a.seg_counts = pl.l1_counts;
  a.seg_zstride = pl.l1_zstride;
  a.seg_stride = pl.sub_stride;
  a.region_first = a.rbase[a.me];
  a.dst = (uint4*)pl.mid;
  a.dst_stride = pl.cap;
  a.cursor = pl.cursor;
  a.capacity = pl.cap;
  a.nbins = pl.F;
  a.level = 2;
  // x CTAs per region (each takes every x-th tile of all the region's streams), chosen so that all CTAs fill whole waves of `ctas` resident CTAs
  // (C1 = 220, x = 2 ran 440 CTAs = 1.49 waves on 296 slots: 26 % of the second wave idle)
  int x = 1;
  double best = 0;
  const uint64_t region_recs = pl.base_off ? (uint64_t)pl.F * pl.cap : pl.sub_stride;
  const uint64_t tiles = nsub * ((region_recs * rb + kTmaTileBytes - 1) / kTmaTileBytes);
  for (int cand = 1; cand <= 16 && (uint64_t)cand <= std::max<uint64_t>(1, tiles / 4); cand++) {
    uint64_t total = (uint64_t)cand * regions, waves = (total + ctas - 1) / ctas;
    double eff = (double)total / (double)(waves * ctas);
    if (eff > best + 0.02) {
      best = eff;
      x = cand;
    }
  }
  launch_split(rb, dim3(x, regions), a, bp, s);
  return 1;
}
int launch_sort_reduce(int rb, const ShuffleBuffers& b, uint32_t B, uint32_t cap, int sm_count,
                       cudaStream_t s) {
  int grid = (int)(B < (uint32_t)(2 * sm_count) ? B : (uint32_t)(2 * sm_count));
  if (grid < 1) grid = 1;
  // u64 keys: the register-pipelined variant -- the key range of a bin from its index (key-ordered sub-bins) or from a
  // min / max over its records (hash sub-bins of clustered or sequential keys: 2.18 -> see profiles/README.md)
  if (rb == 16) {
    const bool multi = !(b.stride || b.nseg == 1), minmax = !(b.hint_S > 1);
    if (!multi && !minmax) k_sort_reduce_u64<false, false><<<grid, kSortThreads, sort_smem_bytes(16), s>>>(b, B, cap);
    if (!multi && minmax) k_sort_reduce_u64<false, true><<<grid, kSortThreads, sort_smem_bytes(16), s>>>(b, B, cap);
    if (multi && !minmax) k_sort_reduce_u64<true, false><<<grid, kSortThreads, sort_smem_bytes(16), s>>>(b, B, cap);
    if (multi && minmax) k_sort_reduce_u64<true, true><<<grid, kSortThreads, sort_smem_bytes(16), s>>>(b, B, cap);
    return 1;
  }
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
int launch_checksum_out(int rb, const ShuffleBuffers& b, uint32_t B, const BinParams& bp, uint32_t bin_base,
                        uint64_t* acc6, cudaStream_t s) {
  int grid = g_sm_count * 8;
  DISPATCH_RB(rb, (k_checksum_out<RB><<<grid, 256, 0, s>>>(b, B, bp, bin_base, (unsigned long long*)acc6)));
  return 1;
}

}  // namespace mrhbm
