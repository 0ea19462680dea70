// mrhbm_kernels.h -- host-callable launchers of the sm_100a kernels (internal, C++).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mrhbm {

#ifdef __CUDACC__
#define MRHBM_HD __host__ __device__
#else
#define MRHBM_HD
#endif

struct BinParams {
  uint32_t P;            // partitions
  uint32_t S;            // sub-bins per partition
  uint32_t partitioner;  // MRHBM_PART_*
  uint32_t ordered;      // 1: sub-bin = top key bits (partition becomes one ascending run)
                         // 0: sub-bin = further hash bits (partition = S ascending runs)
  uint32_t ctr_shift;    // hist / cursor counters live at index bin << ctr_shift: the L2
                         // atomic unit serialises per 32 B sector, so counters are spread out
  // multi-GPU: partition p is owned by rank p % world and sits in slot pbase[p % world] + p / world,
  // so that the bins of one destination rank are contiguous (its all-to-all send region)
  uint32_t world;        // 1 = single GPU (slot == p)
  uint32_t wshift;       // log2(world) when world is a power of two, else 0xffffffff
  uint32_t pbase[9];
  uint32_t rep_shift;    // few, heavily hit bins: 2^rep_shift counter copies per bin (copy = CTA id),
                         // i.e. bin b owns the consecutive virtual bins [b << rep_shift, (b+1) << rep_shift)
};

MRHBM_HD inline uint32_t partition_slot(const BinParams& bp, uint32_t pid) {
  return bp.world > 1 ? bp.pbase[pid % bp.world] + pid / bp.world : pid;
}

constexpr int kSplitTileBytes = 40 * 1024;  // bytes of records k_split_tma partitions in shared memory at a time
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
  // Where the records of bin b live: nseg segments, segment s holds
  // seg_off[s][b+1]-seg_off[s][b] records at src + (seg_base[s] + seg_off[s][b]) records.
  // Single GPU: one segment == the scattered buffer (seg_off[0] = bin_off).  After the
  // all-to-all: one segment per source rank inside the receive buffer.
  const void* src;
  uint32_t nseg;
  const uint32_t* seg_off[8];
  uint64_t seg_base[8];
  // Optimistic single-pass layout (no histogram pass): bin b owns the fixed slots
  // [b*stride, (b+1)*stride) of mid / out and its fill level is its scatter cursor.
  uint32_t stride;     // 0 = exact layout through bin_off
  uint32_t out_stride; // optimistic layout: the groups of bin b start at out slot b * out_stride (<= stride: a
                       // duplicate-heavy bin holds many records but at most one CTA's worth of distinct keys)
  uint32_t ctr_shift;  // cursor[b << ctr_shift]
  // key-ordered sub-bins of u64 keys (sub = mulhi(key, S)): lets the sort kernel skip the
  // min/max pass.  hint_S = 0: unknown.
  uint32_t hint_S;
  uint64_t hint_q;     // floor(2^64 / S)
  uint32_t rep_shift;  // bin_off / seg_off are indexed by virtual bin = bin << rep_shift
  uint32_t no_reduce;  // MRHBM_RED_NONE: every pair is its own output row (sorted, grouped on the host)
  unsigned long long* span;  // measurement hook or nullptr: k_sort_reduce_u64's first CTA start, last CTA end, -, first CTA end
};
MRHBM_HD inline uint64_t bin_start(const ShuffleBuffers& b, uint32_t bin) {
  return b.stride ? (uint64_t)bin * b.stride : (uint64_t)b.bin_off[(size_t)bin << b.rep_shift];
}
MRHBM_HD inline uint64_t out_start(const ShuffleBuffers& b, uint32_t bin) {
  return b.stride ? (uint64_t)bin * b.out_stride : (uint64_t)b.bin_off[(size_t)bin << b.rep_shift];
}
MRHBM_HD inline uint32_t bin_count(const ShuffleBuffers& b, uint32_t bin) {
  if (b.stride) {
    uint32_t c = b.cursor[(size_t)bin << b.ctr_shift];
    return c < b.stride ? c : b.stride;
  }
  return b.bin_off[(size_t)(bin + 1) << b.rep_shift] - b.bin_off[(size_t)bin << b.rep_shift];
}
enum { CNT_NBIG = 0, CNT_TICKET = 1, CNT_ERR = 2, CNT_TOTAL = 3, CNT_GBIG = 4 };
enum { ERRF_SKEW = 1, ERRF_OVERFLOW = 2, ERRF_CAPACITY = 4, ERRF_KEYLEN = 8 };

// every launcher returns the number of kernels it launched
int launch_gen_u64(void* dst, uint64_t seed, uint64_t start, uint64_t n, cudaStream_t s);
int launch_gen_zipf32(void* dst, uint64_t seed, uint64_t start, uint64_t n, const uint64_t* d_table,
                      uint64_t V, cudaStream_t s);
int launch_hist(int rb, const void* recs, uint64_t n, const BinParams& bp, uint32_t* hist, cudaStream_t s);
// exclusive scan of in[i << shift], i < n: out_excl[n+1]; optional copies: out_copy[i << shift] (scatter
// cursors), out_dense[i] (the counts, densely packed), entries > cap appended to big_list
// scratch: kScanScratchWords words for the multi-CTA variant (nullptr: single CTA)
constexpr uint32_t kScanScratchWords = 4096;
int launch_exscan(const uint32_t* in, uint32_t n, uint32_t* out_excl, uint32_t* out_copy,
                  uint32_t* out_dense, uint32_t cap, uint32_t* big_list, uint32_t* nbig, uint32_t* total,
                  uint32_t shift, cudaStream_t s, uint32_t* scratch = nullptr);
// histogram of the top 8 key bits of `nsample` evenly spaced u64 records (hist256 += counts)
int launch_sample_u64(const void* recs, uint64_t n, uint32_t nsample, uint32_t* hist256, cudaStream_t s);
int launch_scatter(int rb, const void* recs, uint64_t n, const BinParams& bp, uint32_t* cursor, void* mid,
                   cudaStream_t s);
// device-side tokeniser: word starts per 256-byte block, then (after an exclusive scan of the
// block counts) one record per word; *flags gets ERRF_KEYLEN when a word exceeds the key slot
int launch_tok_count(const unsigned char* text, uint64_t len, uint32_t* block_counts, cudaStream_t s);
int launch_tok_emit(int rb, const unsigned char* text, uint64_t len, const uint32_t* block_off, void* recs,
                    uint32_t* flags, cudaStream_t s);
inline uint64_t tok_blocks(uint64_t len) { return (len + 255) / 256; }
// map-side combine of one committed range into the global table gtab (2^glog 16-byte short entries + 2^(glog-4)
// record-sized long ones = gtab_bytes_host(rb, glog) bytes, zeroed by the caller); flags[0] |= ERRF_SKEW when the table is full, ERRF_OVERFLOW when a u32 sum would wrap (checked
// mode only: unchecked adds are fire-and-forget and the caller verifies afterwards that pairs x flags[2], the
// largest value seen, stays below 2^32)
int launch_combine(int rb, const void* recs, uint64_t n, uint32_t* gtab, uint32_t glog, uint32_t* flags, bool checked,
                   int sm_count, cudaStream_t s);
// one record per non-empty table entry appended to out; *count (zeroed by the caller) += entries
int launch_gtab_compact(int rb, const uint32_t* gtab, uint32_t glog, void* out, uint32_t* count, cudaStream_t s);
uint64_t gtab_bytes_host(int rb, uint32_t glog);
// tot[b] = sum over s < world of all[s * stride + base + b], b < n; *nover += bins (of all `stride`
// bins) whose global total exceeds cap
int launch_sum_src(const uint32_t* all, uint32_t world, uint32_t stride, uint32_t base, uint32_t n,
                   uint32_t* tot, uint32_t cap, uint32_t* nover, cudaStream_t s);
// row r < world: out[r*(n+1) ..] = exclusive scan of all[r*stride + base ..+n), totals[r] = its sum
// acc2[0] = sum over source ranks z < G and regions r in [first, first+nreg) of min(counts[z*zstride + (r << shift)], clamp),
// acc2[1] = the z == me part of it
int launch_region_totals(const uint32_t* counts, uint64_t zstride, uint32_t G, uint32_t me, uint32_t first, uint32_t nreg,
                         uint32_t shift, uint32_t clamp, uint64_t* acc2, cudaStream_t s);
int launch_exscan_rows(const uint32_t* all, uint32_t world, uint32_t stride, uint32_t base, uint32_t n,
                       uint32_t* out, uint32_t* totals, cudaStream_t s);
// Two-level coalesced split.  Optimistic layout (base_off == nullptr): level 1 sends every record to the
// coarse region of its bin inside this rank's region buffer l1 (all C1 regions of the job, sub_stride slots
// each); level 2 of the rank that owns region y pulls it from every rank's buffer (peer[z], NVLink) and splits
// it into fine bins of `cap` slots in mid.  C1 == B_local, F == 1 on one GPU = single level straight into mid.
// Exact layout (base_off = bin offsets after k_hist + k_exscan): both levels write absolute slots, one GPU.
struct SplitPlan {
  uint32_t B;                 // fine bins (of the whole job)
  uint32_t F;                 // fine bins per coarse region (power of two)
  uint32_t C1;                // coarse regions of the whole job (bins level 1 distinguishes, <= 1024)
  uint32_t C1_local;          // regions this rank owns (level 2 splits these)
  uint32_t cap;               // slots per fine bin (optimistic)
  uint64_t sub_stride;        // slots per coarse region in a rank's region buffer
  uint32_t* cursor1;          // level-1 fill levels, [C1 << ctr_shift]
  uint32_t* cursor;           // fine fill levels, [B_local << ctr_shift]
  void* l1;                   // this rank's region buffer: C1 * sub_stride records
  void* mid;                  // fine bins
  uint32_t* err_flags;
  const uint32_t* base_off;
  uint32_t ndest, me;         // ranks, this rank (1, 0 on a single GPU)
  unsigned long long peer[8]; // byte address of every rank's region buffer as mapped into this process
  uint32_t rbase[9], fbase[9];  // first region / first fine bin of every rank
  const uint32_t* l1_counts;  // level 2: fill of region r on rank z at l1_counts[z * l1_zstride + (r << ctr_shift)]
  uint64_t l1_zstride;
  uint32_t* ticket;           // two zeroed words: level 1 hands out its tiles from a counter (the kernel zeroes them again)
  unsigned long long* span;   // measurement hook (MRHBM_TUNE bit 6) or nullptr: level 1's first CTA start, last CTA end, last CTA start, first CTA end (%globaltimer ns)
};
int launch_split_l1(int rb, const void* recs, uint64_t n, const BinParams& bp, const SplitPlan& pl, cudaStream_t s);
int launch_split_l2(int rb, const BinParams& bp, const SplitPlan& pl, cudaStream_t s);
int launch_sort_reduce(int rb, const ShuffleBuffers& b, uint32_t B, uint32_t cap, int sm_count,
                       cudaStream_t s);
int launch_big_bins(int rb, const ShuffleBuffers& b, uint32_t nbig, uint32_t cap, cudaStream_t s);
int launch_compact(int rb, const ShuffleBuffers& b, uint32_t B, void* dst_keys, uint64_t* dst_sums,
                   cudaStream_t s);
int launch_checksum_in(int rb, const void* recs, uint64_t n, uint64_t* acc4, cudaStream_t s);
int launch_checksum_out(int rb, const ShuffleBuffers& b, uint32_t B, const BinParams& bp, uint32_t bin_base,
                        uint64_t* acc6, cudaStream_t s);
void kernels_set_tune(uint32_t bits);  // measurement hooks (MRHBM_TUNE), read once at mrhbm_init
cudaError_t kernels_configure();  // opt-in shared memory sizes; call once per device

}  // namespace mrhbm
