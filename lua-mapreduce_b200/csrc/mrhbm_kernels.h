// mrhbm_kernels.h -- host-callable launchers of the sm_100a kernels (internal, C++).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mrhbm {

struct BinParams;

constexpr int kCapBytes = 32 * 1024;  // record bytes one CTA sorts in shared memory (2 CTAs per SM)
inline uint32_t cap_records(int rb) { return (uint32_t)(kCapBytes / rb); }

struct ShuffleBuffers {
  // all device pointers
  uint32_t* hist;      // [B]
  uint32_t* bin_off;   // [B+1] exclusive scan of hist
  uint32_t* cursor;    // [B]   scatter cursors (start = bin_off)
  uint32_t* ucount;    // [B]   groups per bin after sort+reduce
  uint32_t* uoff;      // [B+1] exclusive scan of ucount
  uint32_t* big_list;  // [B]   bins larger than cap
  uint32_t* counters;  // [8]: 0 = nbig, 1 = ticket, 2 = error flags, 3 = total
  void* mid;           // scattered records, N * RB
  void* out_keys;      // N * key_bytes (runs at bin_off)
  uint64_t* out_sums;  // N
};
enum { CNT_NBIG = 0, CNT_TICKET = 1, CNT_ERR = 2, CNT_TOTAL = 3 };
enum { ERRF_SKEW = 1, ERRF_OVERFLOW = 2 };

// every launcher returns the number of kernels it launched
int launch_gen_u64(void* dst, uint64_t seed, uint64_t start, uint64_t n, cudaStream_t s);
int launch_gen_zipf32(void* dst, uint64_t seed, uint64_t start, uint64_t n, const uint64_t* d_table,
                      uint64_t V, cudaStream_t s);
int launch_hist(int rb, const void* recs, uint64_t n, uint32_t P, uint32_t S, uint32_t partitioner,
                uint32_t ordered, uint32_t ctr_shift, uint32_t* hist, cudaStream_t s);
int launch_exscan(const uint32_t* in, uint32_t n, uint32_t* out_excl, uint32_t* out_copy,
                  uint32_t cap, uint32_t* big_list, uint32_t* nbig, uint32_t* total, uint32_t shift,
                  cudaStream_t s);
int launch_scatter(int rb, const void* recs, uint64_t n, uint32_t P, uint32_t S, uint32_t partitioner,
                   uint32_t ordered, uint32_t ctr_shift, uint32_t* cursor, void* mid, cudaStream_t s);
int launch_sort_reduce(int rb, const ShuffleBuffers& b, uint32_t B, uint32_t cap, int sm_count,
                       cudaStream_t s);
int launch_big_bins(int rb, const ShuffleBuffers& b, uint32_t nbig, uint32_t cap, cudaStream_t s);
int launch_compact(int rb, const ShuffleBuffers& b, uint32_t B, void* dst_keys, uint64_t* dst_sums,
                   cudaStream_t s);
int launch_checksum_in(int rb, const void* recs, uint64_t n, uint64_t* acc4, cudaStream_t s);
int launch_checksum_out(int rb, const ShuffleBuffers& b, uint32_t B, uint32_t P, uint32_t S,
                        uint32_t partitioner, uint32_t ordered, uint64_t* acc6, cudaStream_t s);
cudaError_t kernels_configure();  // opt-in shared memory sizes; call once per device

}  // namespace mrhbm
