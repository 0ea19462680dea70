// mrhbm_api.cu -- host runtime behind the C ABI of include/mrhbm.h: HBM emit pool, map-job
// bookkeeping (commit = replace-by-job-id, abort = discard; mapreduce/job.lua:217-221,
// worker.lua:120-127), the shuffle driver, result iteration.  No CPU fallback anywhere.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mrhbm.h"
#include "mrhbm_comm.h"
#include "mrhbm_kernels.h"
#include "mrhbm_dev.cuh"  // fnv_lua_step, mulhi_u64_u32: the host side partitions the keys that do not fit a slot

using namespace mrhbm;

namespace {
enum { R_OPEN = 0, R_COMMITTED = 1, R_DEAD = 2 };
constexpr uint32_t kErrfLongKeys = 0x80000000u;  // rides on the error-flag all-gather of a multi-GPU shuffle: this rank holds long keys
constexpr size_t kMaxLongKey = 1u << 20;  // bytes; longer keys are refused (MRHBM_E_KEY)
struct LongPair {
  std::string esc;  // the key as a record slot would hold it: 0x00 / 0x01 escaped (order preserving, see mrhbm_emit_str)
  uint64_t value;
};
struct LongGroup {
  std::string esc;
  std::vector<uint64_t> values;  // built-in sum: one element
};

struct Range {
  uint64_t off, cnt;
  std::string job;
  int state;
  uint64_t owner;  // map handle serial
};
constexpr size_t kStageRecs = 1 << 16;
enum { EV_START, EV_CSTART, EV_COMBINE, EV_HIST, EV_PLAN, EV_GATH, EV_SCATTER, EV_EXCH, EV_SORT, EV_BIG, EV_END, EV_PROBE, EV_N };
constexpr uint32_t kSampleKeys = 1u << 18;      // keys sampled to decide between key-ordered and hash sub-bins
constexpr uint64_t kSplitMaxBinsHost = 1024;    // bins one level of k_split_tma distinguishes
}  // namespace

struct mrhbm_ctx {
  mrhbm_config cfg{};
  int dev = 0, rb = 16, kb = 8, sm_count = 148;
  cudaStream_t stream = nullptr;
  // HBM emit pool: committed + open map output, record granularity
  void* pool = nullptr;
  uint64_t pool_cap = 0, pool_used = 0;
  std::vector<Range> ranges;
  uint64_t map_serial = 0;
  // Keys that do not fit a record slot (rare: the reference takes any key length, utils.lua:104-110) never reach the
  // device: their pairs wait on the host, per committed map job, are partitioned and grouped at the barrier (on several
  // GPUs after an all-gather) and merged into the reduce-side iteration at their place in the key order.
  std::map<std::string, std::vector<LongPair>> long_jobs;  // committed pairs by job id (commit = replace-by-id)
  std::vector<std::vector<LongGroup>> long_groups;         // after a shuffle: per owned partition, ascending escaped key
  uint64_t long_group_count = 0;
  bool any_long = false;  // several GPUs: some rank holds such pairs (learnt with the error flags every shuffle gathers anyway)
  // shuffle state
  ShuffleBuffers sb{};
  uint64_t B_cap = 0, mid_cap = 0, out_cap = 0;
  uint32_t B = 0, S = 1, ordered = 1, cap = 0, ctr_shift = 3;
  uint64_t ctr_cap = 0;
  ShuffleBuffers rv{};  // view the result accessors use (single GPU: == sb)
  // multi-GPU (one ctx per rank; partition p is owned by rank p % world)
  int world = 1, rank = 0;
  uint32_t Pl = 0, pbase[9] = {0}, bin_base = 0;
  uint64_t N_recv = 0;
  uint32_t *d_hd = nullptr, *d_hall = nullptr, *d_tot = nullptr, *d_outoff = nullptr, *d_segoff = nullptr;
  uint64_t hd_cap = 0, bl_cap = 0;
  void *recvbuf = nullptr, *bigbuf = nullptr;
  uint64_t recv_cap = 0, big_cap = 0;
  // coarse regions of the two-level split (optimistic layout).  On several GPUs every rank's buffer is exported
  // with cudaIpcGetMemHandle and mapped by its peers: level 2 of a region's owner pulls it from all of them.
  void* regions = nullptr;
  uint64_t regions_cap = 0;                   // records (the same on all ranks)
  void* peer_regions[8] = {nullptr};          // peer_regions[rank] == regions
  bool ipc_failed = false;                    // peer mapping unavailable: NCCL exchange (exact layout) instead
  uint32_t* d_l1all = nullptr;                // all ranks' level-1 fill levels
  uint64_t l1all_cap = 0;
  uint32_t *d_ipc = nullptr, *h_ipc = nullptr;  // 16 words per rank: the IPC handles
  uint32_t *d_sample = nullptr, *h_sample = nullptr;  // 256 buckets of the key sample
  uint32_t tune = 0;                          // MRHBM_TUNE, read once at init, measurement hooks only: bit 0 = no fast path,
                                              // bit 6 = print when the CTAs of level 1 and of the combiner start / end (with bit 24:
                                              // of the u64 sort), bits 8-15 = the combiner's L2 prefetch distance in trips
  std::recursive_mutex mu;                    // entry points serialise per ctx
  uint32_t *d_small = nullptr, *h_small = nullptr;  // 64 words each
  bool no_optimistic = false;  // sticky: a fixed-capacity bin overflowed once (skewed keys)
  bool no_ordered = false;     // sticky: key-ordered sub-bins overflowed once (clustered keys)
  void* l1buf = nullptr;  // coarse regions of the two-level split
  uint64_t l1_cap = 0;
  // device tokeniser scratch (mrhbm_map_wordcount)
  unsigned char* d_tok_text[2] = {nullptr, nullptr};
  uint32_t *d_tok_cnt = nullptr, *d_tok_off = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t tok_ev[2] = {nullptr, nullptr};
  // map-side combiner: the global (L2 resident) hash table and the records compacted out of it
  void *comb = nullptr, *gtab = nullptr;
  uint64_t comb_cap = 0, gtab_cap = 0;  // records
  uint32_t gtab_extra = 0;              // log2 growth of the table beyond its L2-resident size after it filled up once
  bool no_combine = false;              // sticky: more distinct keys than the largest table takes
  bool combine_checked = false;         // sticky: values large enough that u32 sums must be checked add by add
  uint64_t N = 0, groups = 0;
  bool shuffled = false;
  std::vector<uint32_t> h_bin_off, h_uoff;
  std::vector<uint32_t> h_big;  // group-only mode: oversized bins, each holding runs of `cap` rows
  void* ckeys = nullptr;
  uint64_t* csums = nullptr;
  uint64_t c_cap = 0;
  bool compacted = false;
  uint64_t* d_acc = nullptr;       // 8 x u64 scratch
  uint32_t* h_counters = nullptr;  // pinned, 8 x u32
  uint64_t* h_acc = nullptr;       // pinned, 8 x u64
  // zipf table cache
  uint64_t* d_table = nullptr;
  uint64_t d_table_V = 0, d_table_first = 0, d_table_last = 0;
  mrhbm_stats stats{};
  cudaEvent_t ev[EV_N]{};
  Comm* comm = nullptr;
  std::string err;
};

struct mrhbm_map {
  mrhbm_ctx* ctx;
  std::string job;
  std::vector<LongPair> long_pairs;  // keys longer than a record slot (host side, see mrhbm_ctx::long_jobs)
  uint64_t serial;
  unsigned char* stage[2] = {nullptr, nullptr};
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  bool stage_busy[2] = {false, false};
  int cur = 0;
  size_t fill = 0;
  bool async_reads = false;  // pinned user memory still being read by the copy engine
};

struct RunCursor {
  uint64_t pos, end;
};
struct mrhbm_iter {
  mrhbm_ctx* ctx;
  std::vector<uint64_t> valbuf;  // group-only mode: the values of the current key
  std::vector<unsigned char> unesc;  // the current key with its 0x00 / 0x01 bytes restored
  std::vector<unsigned char> keys;
  std::vector<uint64_t> sums;
  std::vector<RunCursor> runs;
  std::vector<uint32_t> heap;  // the runs that still have rows, a binary min-heap on their current key (heap.lua:29-93)
  uint64_t base = 0;
  const std::vector<LongGroup>* longs = nullptr;  // the partition's long-key groups and the next one to hand out
  size_t long_pos = 0;
  unsigned char keybuf[8];
  bool sorted = true;
  size_t cur_run = 0;
};

namespace {

int fail(mrhbm_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}
#define CU(c, expr)                                                                               \
  do {                                                                                            \
    cudaError_t e_ = (expr);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      cudaGetLastError();                                                                         \
      return fail((c), e_ == cudaErrorMemoryAllocation ? MRHBM_E_NOMEM : MRHBM_E_CUDA,            \
                  "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__,         \
                  cudaGetErrorString(e_));                                                        \
    }                                                                                             \
  } while (0)

// Every entry point that touches CUDA runs with the ctx's device current (a ctx may be driven from any
// thread, and several ctxs on different GPUs may share one thread) and holds the ctx lock: the handle is
// single-caller by contract (include/mrhbm.h), the lock turns a violation into serialisation instead of a race.
struct Entry {
  std::lock_guard<std::recursive_mutex> lk;
  int prev = -1;
  explicit Entry(mrhbm_ctx* c) : lk(c->mu) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != c->dev) cudaSetDevice(c->dev);
    else prev = -1;
  }
  ~Entry() {
    if (prev >= 0) cudaSetDevice(prev);
    cudaGetLastError();
  }
};

int pool_reserve(mrhbm_ctx* c, uint64_t n, uint64_t* off) {
  if (c->pool_used + n > c->pool_cap) {
    uint64_t want = std::max<uint64_t>(c->pool_used + n, c->pool_cap * 2);
    want = std::max<uint64_t>(want, 1 << 16);
    void* np = nullptr;
    CU(c, cudaMalloc(&np, want * c->rb));
    if (c->pool_used) CU(c, cudaMemcpyAsync(np, c->pool, c->pool_used * c->rb, cudaMemcpyDeviceToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    if (c->pool) CU(c, cudaFree(c->pool));
    c->pool = np;
    c->pool_cap = want;
  }
  *off = c->pool_used;
  c->pool_used += n;
  return 0;
}

void add_range(mrhbm_map* m, uint64_t off, uint64_t cnt) {
  mrhbm_ctx* c = m->ctx;
  if (!c->ranges.empty()) {
    Range& r = c->ranges.back();
    if (r.owner == m->serial && r.state == R_OPEN && r.off + r.cnt == off) {
      r.cnt += cnt;
      return;
    }
  }
  c->ranges.push_back(Range{off, cnt, m->job, R_OPEN, m->serial});
}

int stage_flush(mrhbm_map* m) {
  mrhbm_ctx* c = m->ctx;
  if (!m->fill) return 0;
  uint64_t off;
  int rc = pool_reserve(c, m->fill, &off);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync((char*)c->pool + off * c->rb, m->stage[m->cur], m->fill * c->rb, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaEventRecord(m->stage_ev[m->cur], c->stream));
  m->stage_busy[m->cur] = true;
  add_range(m, off, m->fill);
  m->fill = 0;
  m->cur ^= 1;
  if (m->stage_busy[m->cur]) {
    CU(c, cudaEventSynchronize(m->stage_ev[m->cur]));
    m->stage_busy[m->cur] = false;
  }
  return 0;
}

int stage_slot(mrhbm_map* m, unsigned char** slot) {
  mrhbm_ctx* c = m->ctx;
  if (!m->stage[0]) {
    for (int i = 0; i < 2; i++) {
      CU(c, cudaHostAlloc((void**)&m->stage[i], kStageRecs * c->rb, cudaHostAllocDefault));
      CU(c, cudaEventCreateWithFlags(&m->stage_ev[i], cudaEventDisableTiming));
    }
  }
  if (m->fill == kStageRecs) {
    int rc = stage_flush(m);
    if (rc) return rc;
  }
  *slot = m->stage[m->cur] + m->fill * c->rb;
  m->fill++;
  return 0;
}

void map_release(mrhbm_map* m) {
  for (int i = 0; i < 2; i++) {
    if (m->stage_busy[i]) cudaEventSynchronize(m->stage_ev[i]);
    if (m->stage_ev[i]) cudaEventDestroy(m->stage_ev[i]);
    if (m->stage[i]) cudaFreeHost(m->stage[i]);
  }
  delete m;
}

void invalidate(mrhbm_ctx* c) {
  c->shuffled = false;
  c->compacted = false;
  c->long_groups.clear();
  c->long_group_count = 0;
}

// per-bin arrays (offsets, group counts, spread-out atomic counters) for B bins
int ensure_small_arrays(mrhbm_ctx* c, uint64_t B) {
  if (B + 1 > c->B_cap) {
    uint64_t nb = B + 1 + (B >> 2);
    uint32_t** ptrs[] = {&c->sb.bin_off, &c->sb.ucount, &c->sb.uoff, &c->sb.big_list};
    for (auto p : ptrs) {
      if (*p) CU(c, cudaFree(*p));
      *p = nullptr;
      CU(c, cudaMalloc((void**)p, nb * sizeof(uint32_t)));
    }
    c->B_cap = nb;
  }
  if ((B << c->ctr_shift) > c->ctr_cap) {  // spread-out atomic counters
    uint64_t nc = (B + (B >> 2) + 1) << c->ctr_shift;
    uint32_t** ptrs[] = {&c->sb.hist, &c->sb.cursor};
    for (auto p : ptrs) {
      if (*p) CU(c, cudaFree(*p));
      *p = nullptr;
      CU(c, cudaMalloc((void**)p, nc * sizeof(uint32_t)));
    }
    c->ctr_cap = nc;
  }
  if (!c->sb.counters) CU(c, cudaMalloc((void**)&c->sb.counters, 8 * sizeof(uint32_t)));
  return 0;
}

int ensure_buffers(mrhbm_ctx* c, uint64_t B, uint64_t N) {
  int rc = ensure_small_arrays(c, B);
  if (rc) return rc;
  uint64_t need = std::max<uint64_t>(N, 1);
  if (need > c->mid_cap) {
    if (c->sb.mid) CU(c, cudaFree(c->sb.mid));
    c->sb.mid = nullptr;
    c->mid_cap = 0;
    CU(c, cudaMalloc(&c->sb.mid, need * c->rb));
    c->mid_cap = need;
  }
  if (need > c->out_cap) {
    if (c->sb.out_keys) CU(c, cudaFree(c->sb.out_keys));
    if (c->sb.out_sums) CU(c, cudaFree(c->sb.out_sums));
    c->sb.out_keys = nullptr;
    c->sb.out_sums = nullptr;
    c->out_cap = 0;
    CU(c, cudaMalloc(&c->sb.out_keys, need * c->kb));
    CU(c, cudaMalloc((void**)&c->sb.out_sums, need * sizeof(uint64_t)));
    c->out_cap = need;
  }
  return 0;
}

// committed ranges, adjacent ones merged
std::vector<std::pair<uint64_t, uint64_t>> live_ranges(const mrhbm_ctx* c) {
  std::vector<std::pair<uint64_t, uint64_t>> v;
  for (const Range& r : c->ranges) {
    if (r.state != R_COMMITTED || !r.cnt) continue;
    if (!v.empty() && v.back().first + v.back().second == r.off)
      v.back().second += r.cnt;
    else
      v.emplace_back(r.off, r.cnt);
  }
  return v;
}

int ensure_compact(mrhbm_ctx* c) {
  if (!c->shuffled) return fail(c, MRHBM_E_INVAL, "no shuffle result (call mrhbm_shuffle first)");
  if (c->compacted) return 0;
  uint64_t need = std::max<uint64_t>(c->groups, 1);
  if (need > c->c_cap) {
    if (c->ckeys) CU(c, cudaFree(c->ckeys));
    if (c->csums) CU(c, cudaFree(c->csums));
    c->ckeys = nullptr;
    c->csums = nullptr;
    c->c_cap = 0;
    CU(c, cudaMalloc(&c->ckeys, need * c->kb));
    CU(c, cudaMalloc((void**)&c->csums, need * sizeof(uint64_t)));
    c->c_cap = need;
  }
  launch_compact(c->rb, c->rv, c->B, c->ckeys, c->csums, c->stream);
  CU(c, cudaGetLastError());
  CU(c, cudaStreamSynchronize(c->stream));
  c->compacted = true;
  return 0;
}

float ev_ms(mrhbm_ctx* c, int a, int b) {
  float ms = 0;
  if (cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return ms;
}

}  // namespace

extern "C" {

int mrhbm_abi_version(void) { return MRHBM_ABI_VERSION; }

int mrhbm_init(const mrhbm_config* cfg, mrhbm_ctx** out) {
  if (!cfg || !out || cfg->struct_size != sizeof(mrhbm_config)) return MRHBM_E_INVAL;
  *out = nullptr;
  mrhbm_ctx* c = new mrhbm_ctx();
  *out = c;  // returned even on failure so that the caller can read last_error, then destroy
  c->cfg = *cfg;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(c, MRHBM_E_NODEVICE, "no CUDA device (%s): mrhbm has no CPU path",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  int dev = cfg->device;
  if (dev < 0) CU(c, cudaGetDevice(&dev));
  if (dev >= ndev) return fail(c, MRHBM_E_INVAL, "device %d out of range (%d devices)", dev, ndev);
  CU(c, cudaSetDevice(dev));
  cudaDeviceProp prop;
  CU(c, cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(c, MRHBM_E_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev,
                prop.major, prop.minor);
  c->dev = dev;
  c->sm_count = prop.multiProcessorCount;
  if (cfg->num_partitions == 0) return fail(c, MRHBM_E_INVAL, "num_partitions must be >= 1");
  if (cfg->key_kind == MRHBM_KEY_U64) {
    c->rb = 16;
    c->kb = 8;
    if (cfg->partitioner > MRHBM_PART_WORDHASH) return fail(c, MRHBM_E_INVAL, "unknown partitioner");
  } else if (cfg->key_kind == MRHBM_KEY_STR) {
    uint32_t m = cfg->max_key_bytes ? cfg->max_key_bytes : 27;
    if (m <= 27) c->rb = 32;
    else if (m <= 59) c->rb = 64;
    else if (m <= 123) c->rb = 128;
    else return fail(c, MRHBM_E_KEY, "max_key_bytes %u exceeds the largest record class (123)", m);
    c->kb = c->rb - 4;
    if (cfg->partitioner != MRHBM_PART_FNV_LUA && cfg->partitioner != MRHBM_PART_WORDHASH)
      return fail(c, MRHBM_E_INVAL, "string keys need partitioner FNV_LUA or WORDHASH");
  } else
    return fail(c, MRHBM_E_INVAL, "unknown key_kind %u", cfg->key_kind);
  if (cfg->reducer != MRHBM_RED_SUM && cfg->reducer != MRHBM_RED_NONE) return fail(c, MRHBM_E_INVAL, "unknown reducer %u", cfg->reducer);
  if (cfg->reducer == MRHBM_RED_NONE && cfg->combiner) return fail(c, MRHBM_E_INVAL, "a general reducer cannot be combined on the device (combiner must be 0)");
  c->sb.no_reduce = cfg->reducer == MRHBM_RED_NONE;
  CU(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (int i = 0; i < EV_N; i++) CU(c, cudaEventCreate(&c->ev[i]));
  CU(c, kernels_configure());
  CU(c, cudaMalloc((void**)&c->d_acc, 8 * sizeof(uint64_t)));
  CU(c, cudaHostAlloc((void**)&c->h_counters, 8 * sizeof(uint32_t), cudaHostAllocDefault));
  CU(c, cudaHostAlloc((void**)&c->h_acc, 8 * sizeof(uint64_t), cudaHostAllocDefault));
  CU(c, cudaMalloc((void**)&c->d_small, (64 + kScanScratchWords) * sizeof(uint32_t)));  // [64..): scan scratch
  CU(c, cudaHostAlloc((void**)&c->h_small, 64 * sizeof(uint32_t), cudaHostAllocDefault));
  c->Pl = cfg->num_partitions;
  for (int r = 1; r <= 8; r++) c->pbase[r] = cfg->num_partitions;
  c->cap = (cfg->flags & MRHBM_F_SMALL_BINS) ? 96 : cap_records(c->rb);
  CU(c, cudaMalloc((void**)&c->d_sample, 256 * sizeof(uint32_t)));
  CU(c, cudaHostAlloc((void**)&c->h_sample, 256 * sizeof(uint32_t), cudaHostAllocDefault));
  if (const char* e = getenv("MRHBM_TUNE")) c->tune = (uint32_t)strtoul(e, nullptr, 0);  // measurement hooks, read once
  kernels_set_tune(c->tune);
  if (cfg->reserve_pairs) {
    uint64_t off;
    int rc = pool_reserve(c, cfg->reserve_pairs, &off);
    if (rc) return rc;
    c->pool_used = 0;
  }
  return MRHBM_OK;
}

void mrhbm_destroy(mrhbm_ctx* c) {
  if (!c) return;
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  if (c->stream) cudaSetDevice(c->dev);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int r = 0; r < 8; r++)
    if (c->world > 1 && r != c->rank && c->peer_regions[r]) cudaIpcCloseMemHandle(c->peer_regions[r]);
  if (c->d_ipc) cudaFree(c->d_ipc);
  if (c->h_ipc) cudaFreeHost(c->h_ipc);
  if (c->h_sample) cudaFreeHost(c->h_sample);
  if (c->comm) comm_destroy(c->comm);
  void* frees[] = {c->pool,        c->sb.hist,     c->sb.bin_off, c->sb.cursor,   c->sb.ucount, c->sb.uoff,
                   c->sb.big_list, c->sb.counters, c->sb.mid,     c->sb.out_keys, c->sb.out_sums, c->ckeys,
                   c->csums,       c->d_acc,       c->d_table,    c->d_hd,        c->d_hall,    c->d_tot,
                   c->d_outoff,    c->d_segoff,    c->recvbuf,    c->bigbuf,      c->d_small,   c->comb,
                   c->gtab,        c->l1buf,       c->regions,    c->d_l1all,     c->d_sample,   c->d_tok_text[0],
                   c->d_tok_text[1], c->d_tok_cnt, c->d_tok_off};
  for (void* p : frees)
    if (p) cudaFree(p);
  if (c->h_counters) cudaFreeHost(c->h_counters);
  if (c->h_acc) cudaFreeHost(c->h_acc);
  if (c->h_small) cudaFreeHost(c->h_small);
  for (int i = 0; i < EV_N; i++)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (int i = 0; i < 2; i++)
    if (c->tok_ev[i]) cudaEventDestroy(c->tok_ev[i]);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (prev >= 0) cudaSetDevice(prev);
  cudaGetLastError();
  delete c;
}

const char* mrhbm_last_error(const mrhbm_ctx* c) { return c ? c->err.c_str() : "null ctx"; }
uint32_t mrhbm_record_bytes(const mrhbm_ctx* c) { return c ? (uint32_t)c->rb : 0; }

void* mrhbm_host_alloc(mrhbm_ctx* c, size_t bytes) {
  void* p = nullptr;
  Entry g(c);
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    if (c) c->err = "cudaHostAlloc failed";
    return nullptr;
  }
  return p;
}
void mrhbm_host_free(mrhbm_ctx*, void* p) {
  if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------------------
// map side
// ---------------------------------------------------------------------------
int mrhbm_map_begin(mrhbm_ctx* c, const char* job, mrhbm_map** out) {
  if (!c || !job || !out) return MRHBM_E_INVAL;
  Entry g(c);
  if (!c->stream) return fail(c, MRHBM_E_INVAL, "ctx failed to initialise");
  mrhbm_map* m = new mrhbm_map();
  m->ctx = c;
  m->job = job;
  m->serial = ++c->map_serial;
  *out = m;
  return MRHBM_OK;
}

int mrhbm_emit_str(mrhbm_map* m, const void* key, size_t klen, uint32_t value) {
  if (!m) return MRHBM_E_INVAL;
  mrhbm_ctx* c = m->ctx;
  Entry g(c);
  if (c->cfg.key_kind != MRHBM_KEY_STR) return fail(c, MRHBM_E_INVAL, "ctx holds u64 keys");
  // Key slots are zero padded, so a key byte 0x00 cannot be stored as such.  Bytes 0x00 and 0x01 travel escaped --
  // 0x00 -> 01 01, 0x01 -> 01 02 -- which keeps the bytewise order of the keys ("a" < "a\0" < "a\1" < "ab")
  // and is the identity for text.  The device partitioner hashes the unescaped bytes, the iterator unescapes.
  const unsigned char* k = (const unsigned char*)key;
  size_t extra = 0;
  for (size_t i = 0; i < klen; i++) extra += k[i] <= 1;
  if (klen + extra >= (size_t)c->kb) {  // longer than a record slot: kept on the host (escaped like a slot would hold it)
    if (klen > kMaxLongKey) return fail(c, MRHBM_E_KEY, "key of %zu bytes (limit %zu)", klen, (size_t)kMaxLongKey);
    LongPair lp;
    lp.esc.reserve(klen + extra);
    for (size_t i = 0; i < klen; i++) {
      if (k[i] <= 1) {
        lp.esc.push_back((char)1);
        lp.esc.push_back((char)(k[i] + 1));
      } else {
        lp.esc.push_back((char)k[i]);
      }
    }
    lp.value = value;
    m->long_pairs.push_back(std::move(lp));
    return MRHBM_OK;
  }
  unsigned char* slot;
  int rc = stage_slot(m, &slot);
  if (rc) return rc;
  if (!extra) {
    memcpy(slot, key, klen);
  } else {
    size_t o = 0;
    for (size_t i = 0; i < klen; i++) {
      if (k[i] <= 1) {
        slot[o++] = 1;
        slot[o++] = (unsigned char)(k[i] + 1);
      } else {
        slot[o++] = k[i];
      }
    }
  }
  memset(slot + klen + extra, 0, c->kb - (klen + extra));
  memcpy(slot + c->kb, &value, 4);
  return MRHBM_OK;
}

int mrhbm_emit_u64(mrhbm_map* m, uint64_t key, uint64_t value) {
  if (!m) return MRHBM_E_INVAL;
  mrhbm_ctx* c = m->ctx;
  Entry g(c);
  if (c->cfg.key_kind != MRHBM_KEY_U64) return fail(c, MRHBM_E_INVAL, "ctx holds string keys");
  unsigned char* slot;
  int rc = stage_slot(m, &slot);
  if (rc) return rc;
  memcpy(slot, &key, 8);
  memcpy(slot + 8, &value, 8);
  return MRHBM_OK;
}

int mrhbm_emit_batch(mrhbm_map* m, const void* recs, size_t n) {
  if (!m || (!recs && n)) return MRHBM_E_INVAL;
  mrhbm_ctx* c = m->ctx;
  Entry g(c);
  if (!n) return MRHBM_OK;
  int rc = stage_flush(m);  // keep emission order
  if (rc) return rc;
  uint64_t off;
  rc = pool_reserve(c, n, &off);
  if (rc) return rc;
  cudaPointerAttributes at{};
  bool pinned = cudaPointerGetAttributes(&at, recs) == cudaSuccess && at.type == cudaMemoryTypeHost;
  cudaGetLastError();
  CU(c, cudaMemcpyAsync((char*)c->pool + off * c->rb, recs, n * c->rb, cudaMemcpyHostToDevice, c->stream));
  if (pinned) m->async_reads = true;  // pageable sources are fully staged before the call returns
  add_range(m, off, n);
  return MRHBM_OK;
}

int mrhbm_emit_device(mrhbm_map* m, const void* drecs, size_t n) {
  if (!m || (!drecs && n)) return MRHBM_E_INVAL;
  mrhbm_ctx* c = m->ctx;
  Entry g(c);
  if (!n) return MRHBM_OK;
  int rc = stage_flush(m);
  if (rc) return rc;
  uint64_t off;
  rc = pool_reserve(c, n, &off);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync((char*)c->pool + off * c->rb, drecs, n * c->rb, cudaMemcpyDeviceToDevice, c->stream));
  add_range(m, off, n);
  return MRHBM_OK;
}

int mrhbm_map_gen_u64(mrhbm_map* m, uint64_t seed, uint64_t start, uint64_t n) {
  if (!m) return MRHBM_E_INVAL;
  mrhbm_ctx* c = m->ctx;
  Entry g(c);
  if (c->rb != 16) return fail(c, MRHBM_E_INVAL, "gen_u64 needs a u64-key ctx");
  int rc = stage_flush(m);
  if (rc) return rc;
  uint64_t off;
  rc = pool_reserve(c, n, &off);
  if (rc) return rc;
  launch_gen_u64((char*)c->pool + off * c->rb, seed, start, n, c->stream);
  CU(c, cudaGetLastError());
  add_range(m, off, n);
  return MRHBM_OK;
}

int mrhbm_map_gen_zipf(mrhbm_map* m, uint64_t seed, uint64_t start, uint64_t n, const uint64_t* table,
                       uint64_t V) {
  if (!m || !table || !V) return MRHBM_E_INVAL;
  mrhbm_ctx* c = m->ctx;
  Entry g(c);
  if (c->rb != 32) return fail(c, MRHBM_E_INVAL, "gen_zipf needs 32-byte string records (max_key_bytes <= 27)");
  if (!c->d_table || c->d_table_V != V || c->d_table_first != table[0] || c->d_table_last != table[V - 1]) {
    if (c->d_table) CU(c, cudaFree(c->d_table));
    c->d_table = nullptr;
    CU(c, cudaMalloc((void**)&c->d_table, V * sizeof(uint64_t)));
    CU(c, cudaMemcpyAsync(c->d_table, table, V * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    c->d_table_V = V;
    c->d_table_first = table[0];
    c->d_table_last = table[V - 1];
  }
  int rc = stage_flush(m);
  if (rc) return rc;
  uint64_t off;
  rc = pool_reserve(c, n, &off);
  if (rc) return rc;
  launch_gen_zipf32((char*)c->pool + off * c->rb, seed, start, n, c->d_table, V, c->stream);
  CU(c, cudaGetLastError());
  add_range(m, off, n);
  return MRHBM_OK;
}

int mrhbm_map_wordcount(mrhbm_map* m, const void* text, size_t len, uint64_t* words) {
  if (!m || (!text && len)) return MRHBM_E_INVAL;
  mrhbm_ctx* c = m->ctx;
  Entry g(c);
  if (words) *words = 0;
  if (c->cfg.key_kind != MRHBM_KEY_STR) return fail(c, MRHBM_E_INVAL, "wordcount needs a string-key ctx");
  if (!len) return MRHBM_OK;
  int rc = stage_flush(m);
  if (rc) return rc;
  // The text is tokenised in pieces of <= kTokPiece bytes that end at white space, through two device buffers: the
  // upload of piece k+1 (copy stream) runs while piece k is counted, scanned and emitted (ctx stream).  The scratch
  // (text x 2, per-256-byte-block word counts and offsets) lives with the ctx.
  constexpr size_t kTokPiece = 64u << 20;
  const unsigned char* t = (const unsigned char*)text;
  auto is_space = [](unsigned char ch) { return ch == ' ' || (ch >= 9 && ch <= 13); };
  if (!c->d_tok_text[0]) {
    const uint64_t tcap = kTokPiece + 4096, bcap = tok_blocks(tcap) + 1;
    cudaError_t ea = cudaMalloc((void**)&c->d_tok_text[0], tcap);
    if (ea == cudaSuccess) ea = cudaMalloc((void**)&c->d_tok_text[1], tcap);
    if (ea == cudaSuccess) ea = cudaMalloc((void**)&c->d_tok_cnt, bcap * 4);
    if (ea == cudaSuccess) ea = cudaMalloc((void**)&c->d_tok_off, (bcap + 1) * 4);
    if (ea == cudaSuccess) ea = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && ea == cudaSuccess; i++) ea = cudaEventCreateWithFlags(&c->tok_ev[i], cudaEventDisableTiming);
    if (ea != cudaSuccess) {
      cudaGetLastError();
      return fail(c, MRHBM_E_NOMEM, "wordcount: %s", cudaGetErrorString(ea));
    }
  }
  // undo information: a word that does not fit emits NOTHING of this call
  const size_t nr0 = c->ranges.size();
  const uint64_t last_cnt0 = nr0 ? c->ranges.back().cnt : 0, used0 = c->pool_used;
  const size_t long0 = m->long_pairs.size();
  auto undo = [&]() {
    c->ranges.resize(nr0);
    if (nr0) c->ranges.back().cnt = last_cnt0;
    c->pool_used = used0;
    m->long_pairs.resize(long0);  // (words of a piece that was tokenised on the host)
    m->fill = 0;                  // the staging buffer was empty when the call began (stage_flush above)
  };
  auto piece_end = [&](size_t p) -> size_t {  // end of the piece that starts at p: <= kTokPiece bytes, cut after white space
    size_t e = std::min(len, p + kTokPiece);
    if (e < len) {
      size_t q = e;
      while (q > p && !is_space(t[q - 1])) q--;
      if (q > p) e = q;  // (no white space in 64 MB: one giant "word", the length check below refuses it)
    }
    return e;
  };
  cudaError_t e = cudaSuccess;
  uint64_t total_words = 0;
  uint32_t* flagw = c->d_small + 1;
  size_t p = 0, pe = piece_end(0);
  int k = 0;
  e = cudaMemcpyAsync(c->d_tok_text[0], t, pe, cudaMemcpyHostToDevice, c->copy_stream);
  if (e == cudaSuccess) e = cudaEventRecord(c->tok_ev[0], c->copy_stream);
  while (e == cudaSuccess && rc == 0 && p < len) {
    const size_t n = pe - p, pn = pe, pne = pe < len ? piece_end(pe) : pe;
    unsigned char* d_text = c->d_tok_text[k & 1];
    if (pn < len) {  // next piece's upload: its buffer was last read by piece k-1, whose kernels are behind a sync
      e = cudaMemcpyAsync(c->d_tok_text[(k + 1) & 1], t + pn, pne - pn, cudaMemcpyHostToDevice, c->copy_stream);
      if (e == cudaSuccess) e = cudaEventRecord(c->tok_ev[(k + 1) & 1], c->copy_stream);
      if (e != cudaSuccess) break;
    }
    const uint64_t nb = tok_blocks(n);
    if ((e = cudaStreamWaitEvent(c->stream, c->tok_ev[k & 1], 0)) != cudaSuccess) break;
    launch_tok_count(d_text, n, c->d_tok_cnt, c->stream);
    launch_exscan(c->d_tok_cnt, (uint32_t)nb, c->d_tok_off, nullptr, nullptr, 0xffffffffu, nullptr, nullptr, c->d_small, 0, c->stream,
                  c->d_small + 64);
    if ((e = cudaMemcpyAsync(c->h_small, c->d_small, 4, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) break;
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) break;
    const uint64_t total = c->h_small[0];
    if (total) {
      uint64_t off = 0;
      rc = pool_reserve(c, total, &off);
      if (rc) break;
      if ((e = cudaMemsetAsync(flagw, 0, 4, c->stream)) != cudaSuccess) break;
      launch_tok_emit(c->rb, d_text, n, c->d_tok_off, (char*)c->pool + off * c->rb, flagw, c->stream);
      if ((e = cudaMemcpyAsync(c->h_small + 1, flagw, 4, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) break;
      if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) break;
      if (c->h_small[1] & ERRF_KEYLEN) {
        // a word of this piece does not fit the key slot (a URL in a corpus): the piece's records are dropped again and
        // the piece is tokenised on the host -- words that fit are staged like emit_str pairs, the long ones go to the
        // host-side store of long keys.  Rare by assumption: 64 MB take a fraction of a second here.
        c->pool_used = off;
        uint64_t words_here = 0;
        for (size_t a = p; a < pn && rc == 0;) {
          while (a < pn && is_space(t[a])) a++;
          size_t b = a;
          while (b < pn && !is_space(t[b])) b++;
          if (b > a) {
            rc = mrhbm_emit_str(m, t + a, b - a, 1u);
            words_here++;
          }
          a = b;
        }
        if (rc) break;
        total_words += words_here;
      } else {
        add_range(m, off, total);
        total_words += total;
      }
    }
    p = pn;
    pe = pne;
    k++;
  }
  if (e != cudaSuccess || rc) {
    cudaStreamSynchronize(c->copy_stream);  // an upload may still be reading the caller's text
    cudaStreamSynchronize(c->stream);
    undo();
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(c, MRHBM_E_CUDA, "wordcount: %s", cudaGetErrorString(e));
    }
    return rc;
  }
  if (words) *words = total_words;
  return MRHBM_OK;
}

// ---- synthetic word-count text, host side: the input of the end-to-end word-count measurements
namespace {
// (splitmix64: mrhbm_dev.cuh)
inline int synth_rank_to_key(uint64_t rank, char* out) {  // bijective base 26 prefix + hashed upper-case suffix
  char tmp[16];
  int n = 0, o = 0;
  for (uint64_t r = rank; r > 0; r /= 26) {
    r -= 1;
    tmp[n++] = (char)('a' + (r % 26));
  }
  while (n > 0) out[o++] = tmp[--n];
  const uint64_t h = splitmix64(rank ^ 0xA5A5A5A5A5A5A5A5ull);
  unsigned l = (unsigned)(h % 8);
  if (((h >> 8) % 64) == 0) l = 8 + (unsigned)((h >> 16) % 15);
  for (unsigned j = 0; j < l; j++) out[o++] = (char)('A' + (splitmix64(h + j) % 26));
  return o;
}
inline uint64_t synth_zipf_rank(const uint64_t* T, uint64_t V, uint64_t u) {
  uint64_t lo = 0, hi = V;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (T[mid] < u) lo = mid + 1;
    else hi = mid;
  }
  return lo + 1 > V ? V : lo + 1;
}
}  // namespace

int mrhbm_synth_zipf_text(uint64_t seed, uint64_t first, uint64_t n, uint32_t words_per_line, const uint64_t* table,
                          uint64_t V, void* out, size_t cap, size_t* len, int threads) {
  if (!table || !V || !len || !words_per_line || (!out && cap)) return MRHBM_E_INVAL;
  threads = std::max(1, std::min(threads, 256));
  const uint64_t lines = (n + words_per_line - 1) / words_per_line;
  std::vector<uint64_t> bytes((size_t)threads + 1, 0);
  auto span = [&](int t, uint64_t* l0, uint64_t* l1) {
    *l0 = lines / threads * t;
    *l1 = t + 1 == threads ? lines : lines / threads * (t + 1);
  };
  auto run = [&](int t, char* dst) -> uint64_t {  // dst == nullptr: only measure
    uint64_t l0, l1, total = 0;
    span(t, &l0, &l1);
    char key[32];
    for (uint64_t l = l0; l < l1; l++) {
      const uint64_t w0 = l * words_per_line, w1 = std::min<uint64_t>(n, w0 + words_per_line);
      for (uint64_t w = w0; w < w1; w++) {
        const uint64_t u = splitmix64(seed + (1ull << 41) + first + w);
        const int kl = synth_rank_to_key(synth_zipf_rank(table, V, u), key);
        if (dst) {
          memcpy(dst + total, key, (size_t)kl);
          dst[total + kl] = w + 1 == w1 ? '\n' : ' ';
        }
        total += (uint64_t)kl + 1;
      }
    }
    return total;
  };
  {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back([&, t] { bytes[(size_t)t + 1] = run(t, nullptr); });
    for (auto& x : th) x.join();
  }
  for (int t = 0; t < threads; t++) bytes[(size_t)t + 1] += bytes[(size_t)t];
  *len = (size_t)bytes[(size_t)threads];
  if (*len > cap) return MRHBM_E_INVAL;
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++) th.emplace_back([&, t] { run(t, (char*)out + bytes[(size_t)t]); });
  for (auto& x : th) x.join();
  return MRHBM_OK;
}

int mrhbm_pool_read(mrhbm_ctx* c, uint64_t first, uint64_t n, void* host_out) {
  if (!c || (!host_out && n)) return MRHBM_E_INVAL;
  Entry g(c);
  uint64_t skip = first, done = 0;
  for (auto& r : live_ranges(c)) {
    if (done == n) break;
    if (skip >= r.second) {
      skip -= r.second;
      continue;
    }
    uint64_t take = std::min<uint64_t>(n - done, r.second - skip);
    CU(c, cudaMemcpyAsync((char*)host_out + done * c->rb, (char*)c->pool + (r.first + skip) * c->rb, take * c->rb,
                          cudaMemcpyDeviceToHost, c->stream));
    done += take;
    skip = 0;
  }
  CU(c, cudaStreamSynchronize(c->stream));
  if (done != n) return fail(c, MRHBM_E_INVAL, "pool_read: only %llu of %llu pairs committed", (unsigned long long)done, (unsigned long long)n);
  return MRHBM_OK;
}

int mrhbm_map_commit(mrhbm_map* m) {
  if (!m) return MRHBM_E_INVAL;
  mrhbm_ctx* c = m->ctx;
  Entry g(c);
  int rc = stage_flush(m);
  if (rc == 0 && m->async_reads) {
    cudaError_t e = cudaStreamSynchronize(c->stream);  // caller may reuse its pinned buffers now
    if (e != cudaSuccess) rc = fail(c, MRHBM_E_CUDA, "commit: %s", cudaGetErrorString(e));
  }
  if (rc) {
    mrhbm_map_abort(m);
    return rc;
  }
  for (Range& r : c->ranges) {
    if (r.state == R_COMMITTED && r.job == m->job) r.state = R_DEAD;  // replace-by-job-id
  }
  for (Range& r : c->ranges)
    if (r.owner == m->serial && r.state == R_OPEN) r.state = R_COMMITTED;
  c->long_jobs.erase(m->job);  // replace-by-job-id on the host side too
  if (!m->long_pairs.empty()) c->long_jobs[m->job] = std::move(m->long_pairs);
  invalidate(c);
  map_release(m);
  return MRHBM_OK;
}

void mrhbm_map_abort(mrhbm_map* m) {
  if (!m) return;
  mrhbm_ctx* c = m->ctx;
  Entry g(c);
  cudaStreamSynchronize(c->stream);
  for (Range& r : c->ranges)
    if (r.owner == m->serial && r.state == R_OPEN) r.state = R_DEAD;
  while (!c->ranges.empty() && c->ranges.back().state == R_DEAD &&
         c->ranges.back().off + c->ranges.back().cnt == c->pool_used) {
    c->pool_used = c->ranges.back().off;
    c->ranges.pop_back();
  }
  map_release(m);
}

int mrhbm_reset(mrhbm_ctx* c) {
  if (!c) return MRHBM_E_INVAL;
  Entry g(c);
  c->ranges.clear();
  c->pool_used = 0;
  c->long_jobs.clear();
  invalidate(c);
  return MRHBM_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// shuffle
// ---------------------------------------------------------------------------
namespace {

void set_range_hint(ShuffleBuffers& b, uint32_t S, uint32_t ordered) {
  b.hint_S = (ordered && S > 1) ? S : 0;
  b.hint_q = S > 1 ? (uint64_t)(((unsigned __int128)1 << 64) / S) : 0;
}

BinParams make_bp(const mrhbm_ctx* c, uint32_t S, uint32_t ordered) {
  BinParams bp{};
  bp.rep_shift = 0;
  bp.P = c->cfg.num_partitions;
  bp.S = S;
  bp.partitioner = c->cfg.partitioner;
  bp.ordered = ordered;
  bp.ctr_shift = c->ctr_shift;
  bp.world = (uint32_t)c->world;
  bp.wshift = 0xffffffffu;
  for (uint32_t k = 0; k < 4; k++)
    if ((1u << k) == bp.world) bp.wshift = k;
  for (int r = 0; r <= 8; r++) bp.pbase[r] = c->pbase[r];
  return bp;
}

// all ranks: one u32 from every rank (host value in, host values out); synchronises the stream
int gather_u32(mrhbm_ctx* c, uint32_t mine, uint32_t* all) {
  if (c->world == 1) {
    all[0] = mine;
    return 0;
  }
  c->h_small[0] = mine;
  CU(c, cudaMemcpyAsync(c->d_small, c->h_small, 4, cudaMemcpyHostToDevice, c->stream));
  int rc = comm_allgather_u32(c->comm, c->d_small, c->d_small + 16, 1, c->stream, &c->err);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync(c->h_small + 16, c->d_small + 16, 4 * c->world, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (int r = 0; r < c->world; r++) all[r] = c->h_small[16 + r];
  return 0;
}

// all ranks: k <= 2 words from every rank (host values in, all[rank * k + i] out); synchronises the stream
int gather_words(mrhbm_ctx* c, const uint32_t* mine, int k, uint32_t* all) {
  if (c->world == 1) {
    for (int i = 0; i < k; i++) all[i] = mine[i];
    return 0;
  }
  for (int i = 0; i < k; i++) c->h_small[i] = mine[i];
  CU(c, cudaMemcpyAsync(c->d_small, c->h_small, 4 * k, cudaMemcpyHostToDevice, c->stream));
  int rc = comm_allgather_u32(c->comm, c->d_small, c->d_small + 16, (size_t)k, c->stream, &c->err);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync(c->h_small + 16, c->d_small + 16, 4 * (size_t)k * c->world, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < k * c->world; i++) all[i] = c->h_small[16 + i];
  return 0;
}

int ensure_records(mrhbm_ctx* c, void** buf, uint64_t* cap, uint64_t need);

// The region buffer of the two-level split.  One GPU: a plain allocation.  Several GPUs: COLLECTIVE -- every
// rank calls it with the same `need`; all buffers grow together and their IPC handles are re-exchanged, so that
// level 2 of every rank can pull its regions from every peer.  *shared_ok = false (on all ranks alike) when peer
// mapping is unavailable on any rank.
int ensure_regions(mrhbm_ctx* c, uint64_t need, bool* shared_ok) {
  *shared_ok = true;
  if (c->world == 1) return ensure_records(c, &c->regions, &c->regions_cap, need);
  if (c->ipc_failed) {
    *shared_ok = false;
    return 0;
  }
  if (need <= c->regions_cap) return 0;
  const int G = c->world, me = c->rank;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CU(c, cudaStreamSynchronize(c->stream));
  for (int r = 0; r < G; r++) {
    if (r != me && c->peer_regions[r]) cudaIpcCloseMemHandle(c->peer_regions[r]);
    c->peer_regions[r] = nullptr;
  }
  cudaGetLastError();
  uint32_t flags[8];
  int rc = gather_u32(c, 0, flags);  // every rank has unmapped its peers before anybody frees
  if (rc) return rc;
  need += need / 8 + 1024;
  if (c->regions) CU(c, cudaFree(c->regions));
  c->regions = nullptr;
  c->regions_cap = 0;
  CU(c, cudaMalloc(&c->regions, need * c->rb));
  bool ok = true;
  cudaIpcMemHandle_t h;
  memset(&h, 0, sizeof h);
  if (cudaIpcGetMemHandle(&h, c->regions) != cudaSuccess) {
    cudaGetLastError();
    ok = false;
  }
  memcpy(c->h_ipc, &h, 64);
  CU(c, cudaMemcpyAsync(c->d_ipc, c->h_ipc, 64, cudaMemcpyHostToDevice, c->stream));
  rc = comm_allgather_u32(c->comm, c->d_ipc, c->d_ipc + 16, 16, c->stream, &c->err);
  if (rc) return rc;
  CU(c, cudaMemcpyAsync(c->h_ipc + 16, c->d_ipc + 16, 64 * (size_t)G, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  rc = gather_u32(c, ok ? 0u : 1u, flags);  // a rank that could not export: nobody opens anything
  if (rc) return rc;
  for (int r = 0; r < G; r++) ok = ok && flags[r] == 0;
  if (ok) {
    for (int r = 0; r < G; r++) {
      if (r == me) {
        c->peer_regions[r] = c->regions;
        continue;
      }
      cudaIpcMemHandle_t hr;
      memcpy(&hr, c->h_ipc + 16 + 16 * r, 64);
      if (cudaIpcOpenMemHandle(&c->peer_regions[r], hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        c->peer_regions[r] = nullptr;
        ok = false;
      }
    }
    rc = gather_u32(c, ok ? 0u : 1u, flags);
    if (rc) return rc;
    for (int r = 0; r < G; r++) ok = ok && flags[r] == 0;
  }
  if (!ok) {
    for (int r = 0; r < G; r++) {
      if (r != me && c->peer_regions[r]) cudaIpcCloseMemHandle(c->peer_regions[r]);
      c->peer_regions[r] = nullptr;
    }
    cudaGetLastError();
    c->ipc_failed = true;
    *shared_ok = false;
    return 0;
  }
  c->regions_cap = need;
  return 0;
}

int ensure_multi_buffers(mrhbm_ctx* c, uint64_t B, uint64_t Bl) {
  const int G = c->world;
  if (B > c->hd_cap) {
    uint64_t nb = B + (B >> 2) + 1;
    for (uint32_t** p : {&c->d_hd, &c->d_hall}) {
      if (*p) CU(c, cudaFree(*p));
      *p = nullptr;
    }
    CU(c, cudaMalloc((void**)&c->d_hd, nb * 4));
    CU(c, cudaMalloc((void**)&c->d_hall, nb * 4 * G));
    c->hd_cap = nb;
  }
  if (Bl + 1 > c->bl_cap) {
    uint64_t nb = Bl + (Bl >> 2) + 2;
    for (uint32_t** p : {&c->d_tot, &c->d_outoff, &c->d_segoff}) {
      if (*p) CU(c, cudaFree(*p));
      *p = nullptr;
    }
    CU(c, cudaMalloc((void**)&c->d_tot, nb * 4));
    CU(c, cudaMalloc((void**)&c->d_outoff, nb * 4));
    CU(c, cudaMalloc((void**)&c->d_segoff, nb * 4 * G));
    c->bl_cap = nb;
  }
  return 0;
}

int ensure_records(mrhbm_ctx* c, void** buf, uint64_t* cap, uint64_t need);
int gather_u32(mrhbm_ctx* c, uint32_t mine, uint32_t* all);

int ensure_records(mrhbm_ctx* c, void** buf, uint64_t* cap, uint64_t need) {
  need = std::max<uint64_t>(need, 1);
  if (need <= *cap) return 0;
  if (*buf) CU(c, cudaFree(*buf));
  *buf = nullptr;
  *cap = 0;
  CU(c, cudaMalloc(buf, need * c->rb));
  *cap = need;
  return 0;
}

int ensure_out(mrhbm_ctx* c, uint64_t need) {
  need = std::max<uint64_t>(need, 1);
  if (need <= c->out_cap) return 0;
  if (c->sb.out_keys) CU(c, cudaFree(c->sb.out_keys));
  if (c->sb.out_sums) CU(c, cudaFree(c->sb.out_sums));
  c->sb.out_keys = nullptr;
  c->sb.out_sums = nullptr;
  c->out_cap = 0;
  CU(c, cudaMalloc(&c->sb.out_keys, need * c->kb));
  CU(c, cudaMalloc((void**)&c->sb.out_sums, need * sizeof(uint64_t)));
  c->out_cap = need;
  return 0;
}

struct Src {
  const char* p;
  uint64_t n;
};
constexpr uint32_t kGtabLog = 21;       // global combiner table: 2^21 16-byte entries = 32 MB, resident in the 126 MB L2
constexpr uint32_t kGtabMaxExtra = 5;   // ... grown up to 32x (1 GB) when the keys do not fit

// The pairs the shuffle partitions: the committed pool ranges, or -- with a combiner declared
// (job.lua:92-96,198-202) and enough pairs to pay for it -- one record per distinct key of this rank's pairs
// with their sum (shared-memory table per SM + one L2-resident global table, see k_combine).  A rank-local
// decision: the stages that follow only see "the pairs this rank contributes".  EV_START .. EV_CSTART.
int collect_sources(mrhbm_ctx* c, std::vector<Src>& srcs, uint64_t* N, mrhbm_stats& st) {
  srcs.clear();
  uint64_t n = 0;
  for (auto& r : live_ranges(c)) {
    srcs.push_back(Src{(const char*)c->pool + r.first * c->rb, r.second});
    n += r.second;
  }
  *N = n;
  if (!c->cfg.combiner || c->no_combine || n < (1u << 20)) return 0;
  cudaStream_t s = c->stream;
  bool checked = c->combine_checked;
  for (;;) {
    const uint32_t glog = kGtabLog + c->gtab_extra;
    const uint64_t tab_bytes = gtab_bytes_host(c->rb, glog);
    const uint64_t slots = (1ull << glog) + (c->rb == 16 ? 0 : 1ull << (glog - 4));  // records the compaction can emit
    int rc = ensure_records(c, &c->gtab, &c->gtab_cap, (tab_bytes + c->rb - 1) / c->rb);
    if (rc) return rc;
    rc = ensure_records(c, &c->comb, &c->comb_cap, slots);
    if (rc) return rc;
    CU(c, cudaMemsetAsync(c->gtab, 0, tab_bytes, s));
    CU(c, cudaMemsetAsync(c->d_small, 0, 3 * sizeof(uint32_t), s));  // [0] flags, [1] records out, [2] largest value
    if (c->tune & 64u) {  // measurement hook: k_combine's first CTA start, last CTA end, first CTA end at words 8..13
      CU(c, cudaMemsetAsync(c->d_small + 8, 0xff, 8, s));
      CU(c, cudaMemsetAsync(c->d_small + 10, 0, 8, s));
      CU(c, cudaMemsetAsync(c->d_small + 12, 0xff, 8, s));
    }
    for (auto& sr : srcs)
      st.launches += launch_combine(c->rb, sr.p, sr.n, (uint32_t*)c->gtab, glog, c->d_small, checked, c->sm_count, s);
    st.launches += launch_gtab_compact(c->rb, (const uint32_t*)c->gtab, glog, c->comb, c->d_small + 1, s);
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(c->h_small, c->d_small, 14 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));
    if (c->tune & 64u) {
      const uint64_t* t = (const uint64_t*)(c->h_small + 8);
      fprintf(stderr, "[mrhbm] k_combine: first CTA ends after %.3f ms, last after %.3f ms\n", (t[2] - t[0]) * 1e-6, (t[1] - t[0]) * 1e-6);
    }
    const uint32_t ef = c->h_small[0];
    if (ef & ERRF_OVERFLOW) return fail(c, MRHBM_E_OVERFLOW, "u32 partial sum overflow while combining a hot key");
    if (ef & ERRF_SKEW) {  // the table filled up: more distinct keys than it takes
      if (c->gtab_extra < kGtabMaxExtra) {
        c->gtab_extra++;
        continue;
      }
      c->no_combine = true;  // the partition / sort / reduce stages take the raw pairs
      return 0;
    }
    if (!checked && c->rb != 16 && (uint64_t)c->h_small[2] * n >= 0xfffffff0ull) {
      // the fire-and-forget adds could have wrapped a u32 sum unnoticed: again, with checked adds (sticky)
      checked = c->combine_checked = true;
      continue;
    }
    srcs.clear();
    srcs.push_back(Src{(const char*)c->comb, c->h_small[1]});
    *N = c->h_small[1];
    return 0;
  }
}

uint32_t pick_sub_bins(const mrhbm_ctx* c, uint64_t n_total, uint32_t copies = 1) {
  // mean bin = capacity of one CTA's shared-memory sort minus 6 sigma of a Poisson fill.  `copies`: records arrive in
  // clumps of up to that many with one key (every rank's combiner emits each key once: a bin then holds Poisson-many
  // KEYS times `copies` records, its sigma is sqrt(copies) times wider -- measured on 8 GPUs: 1 % of the bins of the
  // combined Zipf stream overflowed a 6-sigma-of-records capacity and sent every shuffle to the exact layout)
  uint64_t target = (uint64_t)std::max(1.0, (double)c->cap - 6.0 * std::sqrt((double)c->cap * (double)copies));
  uint64_t P = c->cfg.num_partitions;
  return (uint32_t)std::max<uint64_t>(1, (n_total + P * target - 1) / (P * target));
}

void finish_stats(mrhbm_ctx* c, mrhbm_stats& st) {
  st.ms_total = ev_ms(c, EV_START, EV_END);
  st.ms_combine = ev_ms(c, EV_START, EV_CSTART);
  st.ms_hist = ev_ms(c, EV_COMBINE, EV_HIST);
  st.ms_plan = ev_ms(c, EV_HIST, EV_PLAN);
  st.ms_scatter = ev_ms(c, EV_PLAN, EV_SCATTER);
  st.ms_sort_reduce = ev_ms(c, EV_EXCH, EV_SORT);
  st.ms_bigbins = ev_ms(c, EV_SORT, EV_BIG);
  st.ms_setup = ev_ms(c, EV_CSTART, EV_COMBINE);
  st.ms_finish = ev_ms(c, EV_BIG, EV_END);
  c->stats = st;
}

// ---- the default path: optimistic fixed-capacity layout, no histogram pass, any number of GPUs --------
//
//   level 1  k_split_tma   committed pairs -> coarse regions (this rank's region buffer, all regions of the job)
//   [G > 1]  all-gather of the level-1 fill levels (a few KB): counts for the peers + the barrier that orders
//            every rank's level 1 before anybody's level 2
//   level 2  k_split_tma   the owner of a region pulls it from every rank's region buffer (NVLink bulk copies
//            into shared memory, overlapped with the split of the previous tile) -> fine bins
//   sort     k_sort_reduce_u64 / k_sort_reduce on local bins
// Every bin and region has a fixed capacity (mean + slack); the claim cursor doubles as the count.  A full one
// sets ERRF_CAPACITY, nothing is lost (the input stays in the pool) and the exact layout below takes over, for
// good on this ctx.  Returns 0 with *done = true, 0 with *done = false (fall back), or an error.
int shuffle_fast(mrhbm_ctx* c, std::vector<Src>& live, uint64_t N, uint64_t N_in, mrhbm_stats& st, bool* done) {
  *done = false;
  const int G = c->world, me = c->rank;
  const uint32_t P = c->cfg.num_partitions;
  cudaStream_t s = c->stream;
  int rc = 0;
  // ---- would key-ordered sub-bins be balanced?  sample the keys before anything moves
  const bool want_ordered = c->cfg.key_kind == MRHBM_KEY_U64 && !(c->cfg.flags & MRHBM_F_FORCE_RUNS) && !c->no_ordered;
  uint32_t nonuniform = 0;
  if (want_ordered && N >= 4096) {
    CU(c, cudaMemsetAsync(c->d_sample, 0, 256 * sizeof(uint32_t), s));
    uint64_t total = 0;
    for (auto& r : live) {
      uint32_t ns = (uint32_t)std::min<uint64_t>(r.n, (uint64_t)((double)kSampleKeys * (double)r.n / (double)N) + 1);
      st.launches += launch_sample_u64(r.p, r.n, ns, c->d_sample, s);
      total += ns;
    }
    CU(c, cudaMemcpyAsync(c->h_sample, c->d_sample, 256 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));
    uint32_t mx = 0;
    for (int i = 0; i < 256; i++) mx = std::max(mx, c->h_sample[i]);
    const double mean = (double)total / 256.0;
    nonuniform = (double)mx > mean * 1.08 + 5.0 * std::sqrt(mean) + 2.0;
  }
  // ---- job-wide facts.  On several GPUs this all-gather is also the barrier that keeps a rank from refilling
  // its region buffer while a peer's level 2 of the previous shuffle may still read it.
  uint32_t mine[2] = {(uint32_t)N, nonuniform}, all[16];
  rc = gather_words(c, mine, 2, all);
  if (rc) return rc;
  uint64_t Nglobal = 0, Nmax = 0;
  for (int r = 0; r < G; r++) {
    Nglobal += all[2 * r];
    Nmax = std::max<uint64_t>(Nmax, all[2 * r]);
    nonuniform |= all[2 * r + 1];
  }
  if (Nglobal == 0) return 0;  // nothing to do here: the exact path handles the empty shuffle
  const uint32_t S = pick_sub_bins(c, Nglobal, c->cfg.combiner && G > 1 ? (uint32_t)G : 1u);
  {
    const uint64_t B = (uint64_t)P * S;
    if (B >= (1ull << 31)) return 0;
    const uint32_t ordered = (want_ordered && !nonuniform) || S == 1;
    uint64_t Bl_of[8], Bl_max = 0;
    for (int r = 0; r < G; r++) {
      Bl_of[r] = (uint64_t)(c->pbase[r + 1] - c->pbase[r]) * S;
      Bl_max = std::max(Bl_max, Bl_of[r]);
    }
    const uint64_t Bl = Bl_of[me];
    const uint32_t bin_stride = c->cap;  // slots per fine bin: what one CTA sorts in shared memory
    // ---- levels.  One GPU, few bins: level 1 goes straight into the fine bins.
    SplitPlan pl{};
    const bool single_level = G == 1 && B <= kSplitMaxBinsHost;
    uint32_t F = 1;
    if (!single_level) {
      while ((uint64_t)F * F < B) F <<= 1;  // power of two >= sqrt(B): both levels see about the same fan-out
      for (;; F <<= 1) {
        uint64_t c1 = 0;
        for (int r = 0; r < G; r++) c1 += (Bl_of[r] + F - 1) / F;
        if (c1 <= kSplitMaxBinsHost) break;
      }
      if (F > kSplitMaxBinsHost) return 0;  // more bins than two levels reach: exact path
    }
    pl.rbase[0] = 0;
    for (int r = 0; r < 8; r++) {
      const uint64_t bl = r < G ? Bl_of[r] : 0;
      pl.rbase[r + 1] = pl.rbase[r] + (uint32_t)((bl + F - 1) / F);
      pl.fbase[r] = c->pbase[r] * S;
    }
    pl.fbase[8] = c->pbase[8] * S;
    pl.B = (uint32_t)B;
    pl.F = F;
    pl.C1 = pl.rbase[G];
    pl.C1_local = pl.rbase[me + 1] - pl.rbase[me];
    pl.cap = bin_stride;
    pl.ndest = (uint32_t)G;
    pl.me = (uint32_t)me;
    {
      // one rank's contribution to one region: F of the B bins, hash balanced
      const double mean = (double)Nmax * (double)F / (double)B;
      const double ss = single_level ? (double)bin_stride : mean + 8.0 * std::sqrt(mean) + 64.0;
      if (ss >= 4.0e9) return 0;
      pl.sub_stride = ((uint64_t)ss + 7u) & ~7ull;
    }
    // ---- buffers
    rc = ensure_small_arrays(c, std::max<uint64_t>(Bl, pl.C1));
    if (rc) return rc;
    // k_split_tma addresses destination slots with 32 bits, and lets the records of a run that does not fit (the
    // attempt is abandoned then) land at the start of the run's region: one tile of slack behind the last region
    const uint64_t tile_slack = kSplitTileBytes / c->rb;
    if (Bl * bin_stride + tile_slack >= (1ull << 32) || (uint64_t)pl.C1 * pl.sub_stride + tile_slack >= (1ull << 32)) return 0;
    rc = ensure_records(c, &c->sb.mid, &c->mid_cap, Bl * bin_stride + tile_slack);
    if (rc) return rc;
    rc = ensure_out(c, Bl * (uint64_t)c->cap);
    if (rc) return rc;
    if (!single_level) {
      bool shared_ok = true;
      rc = ensure_regions(c, (uint64_t)pl.C1 * pl.sub_stride + tile_slack, &shared_ok);
      if (rc) return rc;
      if (!shared_ok) {  // peer mapping unavailable (decided by all ranks together): NCCL exchange instead
        c->no_optimistic = true;
        return 0;
      }
      if (G > 1) {
        const uint64_t words = ((uint64_t)pl.C1 << c->ctr_shift) * G;
        if (words > c->l1all_cap) {
          if (c->d_l1all) CU(c, cudaFree(c->d_l1all));
          c->d_l1all = nullptr;
          c->l1all_cap = 0;
          CU(c, cudaMalloc((void**)&c->d_l1all, (words + words / 4) * sizeof(uint32_t)));
          c->l1all_cap = words + words / 4;
        }
      }
    }
    pl.cursor1 = c->sb.hist;
    pl.cursor = c->sb.cursor;
    pl.l1 = single_level ? c->sb.mid : c->regions;
    pl.mid = c->sb.mid;
    pl.err_flags = c->sb.counters + CNT_ERR;
    pl.ticket = c->sb.counters + 5;  // (words 5 and 6 of the 8 counters cleared below)
    pl.base_off = nullptr;
    for (int r = 0; r < 8; r++) pl.peer[r] = (unsigned long long)(uintptr_t)(G == 1 ? (r == 0 ? c->regions : nullptr) : c->peer_regions[r]);
    pl.l1_counts = G == 1 ? c->sb.hist : c->d_l1all;
    pl.l1_zstride = G == 1 ? 0 : ((uint64_t)pl.C1 << c->ctr_shift);
    const BinParams bp = make_bp(c, S, ordered);
    st.attempts++;
    CU(c, cudaMemsetAsync(c->sb.hist, 0, (((uint64_t)pl.C1) << c->ctr_shift) * sizeof(uint32_t), s));
    if (!single_level) CU(c, cudaMemsetAsync(c->sb.cursor, 0, (std::max<uint64_t>(Bl, 1) << c->ctr_shift) * sizeof(uint32_t), s));
    CU(c, cudaMemsetAsync(c->sb.counters, 0, 8 * sizeof(uint32_t), s));
    if (c->tune & 64u) {  // measurement hook: when do the CTAs of level 1 (bit 24 set: of the u64 sort) start and end (%globaltimer)?
      if (!(c->tune & (1u << 24))) pl.span = (unsigned long long*)(c->d_acc + 4);
      CU(c, cudaMemsetAsync(c->d_acc + 4, 0xff, 8, s));
      CU(c, cudaMemsetAsync(c->d_acc + 5, 0, 16, s));
      CU(c, cudaMemsetAsync(c->d_acc + 7, 0xff, 8, s));
    }
    CU(c, cudaEventRecord(c->ev[EV_COMBINE], s));
    CU(c, cudaEventRecord(c->ev[EV_HIST], s));
    for (auto& r : live) st.launches += launch_split_l1(c->rb, r.p, r.n, bp, pl, s);
    CU(c, cudaEventRecord(c->ev[EV_PLAN], s));
    if (G > 1) {
      rc = comm_allgather_u32(c->comm, c->sb.hist, c->d_l1all, (size_t)pl.C1 << c->ctr_shift, s, &c->err);
      if (rc) return rc;
    }
    CU(c, cudaEventRecord(c->ev[EV_GATH], s));
    if (!single_level) st.launches += launch_split_l2(c->rb, bp, pl, s);
    CU(c, cudaEventRecord(c->ev[EV_SCATTER], s));
    CU(c, cudaEventRecord(c->ev[EV_EXCH], s));
    c->sb.src = c->sb.mid;
    c->sb.nseg = 1;
    c->sb.stride = bin_stride;
    c->sb.out_stride = c->cap;
    c->sb.rep_shift = 0;
    c->sb.ctr_shift = c->ctr_shift;
    // single level: the level-1 cursors ARE the fine fill levels
    uint32_t* fine_cursor = single_level ? c->sb.hist : c->sb.cursor;
    ShuffleBuffers v = c->sb;
    v.cursor = fine_cursor;
    set_range_hint(v, S, ordered);
    if ((c->tune & 64u) && (c->tune & (1u << 24))) v.span = (unsigned long long*)(c->d_acc + 4);
    if (Bl) st.launches += launch_sort_reduce(c->rb, v, (uint32_t)Bl, c->cap, c->sm_count, s);
    v.span = nullptr;
    CU(c, cudaEventRecord(c->ev[EV_SORT], s));
    CU(c, cudaEventRecord(c->ev[EV_BIG], s));
    c->h_uoff.assign(Bl + 1, 0);
    if (Bl) {
      st.launches += launch_exscan(c->sb.ucount, (uint32_t)Bl, c->sb.uoff, nullptr, nullptr, 0xffffffffu, nullptr, nullptr,
                                   c->sb.counters + CNT_TOTAL, 0, s, c->d_small + 64);
      CU(c, cudaMemcpyAsync(c->h_uoff.data(), c->sb.uoff, (Bl + 1) * 4, cudaMemcpyDeviceToHost, s));
    }
    // pairs this rank reduced (and how many of them came from peers)
    st.launches += launch_region_totals(pl.l1_counts, pl.l1_zstride, (uint32_t)G, (uint32_t)me, pl.rbase[me], pl.C1_local,
                                        c->ctr_shift, (uint32_t)std::min<uint64_t>(single_level ? bin_stride : pl.sub_stride, 0xffffffffull),
                                        c->d_acc, s);
    CU(c, cudaMemcpyAsync(c->h_acc, c->d_acc, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CU(c, cudaMemcpyAsync(c->h_counters, c->sb.counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU(c, cudaEventRecord(c->ev[EV_END], s));
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(s));
    if (c->tune & 64u)
      fprintf(stderr, (c->tune & (1u << 24)) ? "[mrhbm] u64 sort: CTAs start over %.3f ms, first ends after %.3f ms, last after %.3f ms (level 1 events: %.3f ms)\n" : "[mrhbm] level 1: CTAs start over %.3f ms, first ends after %.3f ms, last after %.3f ms (events: %.3f ms)\n",
              (c->h_acc[6] - c->h_acc[4]) * 1e-6, (c->h_acc[7] - c->h_acc[4]) * 1e-6, (c->h_acc[5] - c->h_acc[4]) * 1e-6,
              ev_ms(c, EV_HIST, EV_PLAN));
    uint32_t ef = c->h_counters[CNT_ERR];
    if (G > 1) {  // every rank must take the same branch; also: nobody leaves while a peer still reads its regions
      uint32_t efs[8];
      rc = gather_u32(c, ef | (c->long_jobs.empty() ? 0u : kErrfLongKeys), efs);
      if (rc) return rc;
      for (int r = 0; r < G; r++) ef |= efs[r];
      c->any_long = (ef & kErrfLongKeys) != 0;
      ef &= ~kErrfLongKeys;
    }
    if (ef & ERRF_OVERFLOW) return fail(c, MRHBM_E_OVERFLOW, "u32 partial sum overflow while combining a hot key");
    if (ef & ERRF_CAPACITY) {  // skewed keys: the exact layout from now on
      c->sb.stride = 0;
      c->no_optimistic = true;
      if ((ordered && S > 1) || nonuniform) c->no_ordered = true;  // key-ordered sub-bins overflowed, or would
      return 0;
    }
    c->h_big.clear();
    c->rv = v;
    c->sb.stride = 0;
    c->B = (uint32_t)Bl;
    c->S = S;
    c->ordered = ordered;
    c->N = N_in;
    c->N_recv = c->h_acc[0];
    c->bin_base = c->pbase[me] * S;
    c->groups = c->h_uoff[Bl];
    c->shuffled = true;
    c->compacted = false;
    st.bins = (uint32_t)Bl;
    st.sub_bins = S;
    st.big_bins = 0;
    st.groups = c->groups;
    st.bytes_exchanged = (c->h_acc[0] - c->h_acc[1]) * (uint64_t)c->rb;
    finish_stats(c, st);
    // plan = level 1, exchange = the fill-level all-gather (waits for the slowest rank's level 1), scatter = level 2
    // (on several GPUs: including the NVLink pulls)
    c->stats.ms_exchange = ev_ms(c, EV_PLAN, EV_GATH);
    c->stats.ms_scatter = ev_ms(c, EV_GATH, EV_SCATTER);
    *done = true;
    return MRHBM_OK;
  }
}

// ---- one GPU, exact layout: hist -> scan -> scatter -> sort+reduce (skewed keys, oversized bins) ---------
int shuffle_single_exact(mrhbm_ctx* c, std::vector<Src>& live, uint64_t N, uint64_t N_in, mrhbm_stats& st) {
  const uint32_t P = c->cfg.num_partitions;
  int rc = 0;
  uint32_t nbig = 0, ordered = 1;
  cudaStream_t s = c->stream;
  uint64_t B = 0;
  uint32_t S = pick_sub_bins(c, N);
  bool skip_ordered = c->no_ordered;
  // A bin that holds more distinct keys than one CTA sorts (ERRF_SKEW) is retried with
  // twice the sub-bins: distinct keys spread, hot keys keep collapsing in k_big_bins.
  for (int widen = 0;; widen++) {
    B = (uint64_t)P * S;
    if (B >= (1ull << 31)) return fail(c, MRHBM_E_INVAL, "too many bins");
    const uint64_t Bv = B;
    rc = ensure_buffers(c, Bv, N);
    if (rc) return rc;
    ordered = ((c->cfg.key_kind == MRHBM_KEY_U64 && !(c->cfg.flags & MRHBM_F_FORCE_RUNS) && widen == 0 && !skip_ordered) || S == 1);
    c->sb.stride = 0;
    for (int attempt = 0;; attempt++) {
      st.attempts++;
      BinParams bp = make_bp(c, S, ordered);
      CU(c, cudaMemsetAsync(c->sb.hist, 0, (Bv << c->ctr_shift) * sizeof(uint32_t), s));
      CU(c, cudaMemsetAsync(c->sb.counters, 0, 8 * sizeof(uint32_t), s));
      CU(c, cudaEventRecord(c->ev[EV_COMBINE], s));
      for (auto& r : live) st.launches += launch_hist(c->rb, r.p, r.n, bp, c->sb.hist, s);
      CU(c, cudaEventRecord(c->ev[EV_HIST], s));
      st.launches += launch_exscan(c->sb.hist, (uint32_t)Bv, c->sb.bin_off, c->sb.cursor, nullptr, c->cap, c->sb.big_list,
                                   c->sb.counters + CNT_NBIG, c->sb.counters + CNT_TOTAL, c->ctr_shift, s, c->d_small + 64);
      CU(c, cudaMemcpyAsync(c->h_counters, c->sb.counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
      CU(c, cudaEventRecord(c->ev[EV_PROBE], s));
      CU(c, cudaEventRecord(c->ev[EV_PLAN], s));
      {
        // exact layout through the two-level coalesced split when there are enough bins (claims relative
        // to bin_off); else the direct scatter
        uint32_t F = 1;
        while ((uint64_t)F * F < B) F <<= 1;
        uint32_t C1 = (uint32_t)((B + F - 1) / F);
        if (B >= 2048 && F <= 1024 && C1 <= 1024) {
          rc = ensure_records(c, &c->l1buf, &c->l1_cap, N);
          if (rc) return rc;
          // relative fill levels: coarse levels live in hist (dead after the scan), fine levels in cursor
          CU(c, cudaMemsetAsync(c->sb.hist, 0, ((uint64_t)C1 << c->ctr_shift) * sizeof(uint32_t), s));
          CU(c, cudaMemsetAsync(c->sb.cursor, 0, (B << c->ctr_shift) * sizeof(uint32_t), s));
          SplitPlan pl{};
          pl.B = (uint32_t)B;
          pl.F = F;
          pl.C1 = pl.C1_local = C1;
          pl.cap = c->cap;
          pl.cursor1 = c->sb.hist;
          pl.cursor = c->sb.cursor;
          pl.l1 = c->l1buf;
          pl.mid = c->sb.mid;
          pl.err_flags = c->sb.counters + CNT_ERR;
          pl.base_off = c->sb.bin_off;
          pl.ndest = 1;
          for (auto& r : live) st.launches += launch_split_l1(c->rb, r.p, r.n, bp, pl, s);
          st.launches += launch_split_l2(c->rb, bp, pl, s);
        } else {
          for (auto& r : live) st.launches += launch_scatter(c->rb, r.p, r.n, bp, c->sb.cursor, c->sb.mid, s);
        }
      }
      CU(c, cudaEventRecord(c->ev[EV_SCATTER], s));
      CU(c, cudaEventRecord(c->ev[EV_EXCH], s));
      c->sb.src = c->sb.mid;
      c->sb.nseg = 1;
      c->sb.seg_off[0] = c->sb.bin_off;
      c->sb.seg_base[0] = 0;
      c->sb.rep_shift = 0;
      set_range_hint(c->sb, S, ordered);
      st.launches += launch_sort_reduce(c->rb, c->sb, (uint32_t)B, c->cap, c->sm_count, s);
      CU(c, cudaEventRecord(c->ev[EV_SORT], s));
      CU(c, cudaGetLastError());
      CU(c, cudaEventSynchronize(c->ev[EV_PROBE]));  // overlaps with scatter / sort on the device
      nbig = c->h_counters[CNT_NBIG];
      if (nbig && ordered && S > 1) {
        // key-ordered sub-bins are unbalanced for this key distribution: redo with hash sub-bins
        ordered = 0;
        c->no_ordered = true;
        continue;
      }
      break;
    }
    st.launches += launch_big_bins(c->rb, c->sb, nbig, c->cap, s);
    CU(c, cudaEventRecord(c->ev[EV_BIG], s));
    c->h_big.clear();
    if (c->sb.no_reduce && nbig) {
      c->h_big.resize(nbig);
      CU(c, cudaMemcpyAsync(c->h_big.data(), c->sb.big_list, nbig * 4, cudaMemcpyDeviceToHost, s));
    }
    st.launches += launch_exscan(c->sb.ucount, (uint32_t)B, c->sb.uoff, nullptr, nullptr, 0xffffffffu, nullptr, nullptr,
                                 c->sb.counters + CNT_TOTAL, 0, s, c->d_small + 64);
    c->h_bin_off.resize(B + 1);
    c->h_uoff.resize(B + 1);
    CU(c, cudaMemcpyAsync(c->h_bin_off.data(), c->sb.bin_off, (B + 1) * 4, cudaMemcpyDeviceToHost, s));
    CU(c, cudaMemcpyAsync(c->h_uoff.data(), c->sb.uoff, (B + 1) * 4, cudaMemcpyDeviceToHost, s));
    CU(c, cudaMemcpyAsync(c->h_counters, c->sb.counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU(c, cudaEventRecord(c->ev[EV_END], s));
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(s));
    uint32_t ef = c->h_counters[CNT_ERR];
    if (ef & ERRF_OVERFLOW) return fail(c, MRHBM_E_OVERFLOW, "u32 partial sum overflow while combining a hot key");
    if (ef & ERRF_SKEW) {
      if (widen < 4 && (uint64_t)P * S * 2 < (1ull << 31)) {
        S *= 2;
        continue;
      }
      return fail(c, MRHBM_E_SKEW, "a bin holds more distinct keys than one SM can sort (%u oversized bins)", nbig);
    }
    break;
  }
  c->rv = c->sb;
  c->B = (uint32_t)B;
  c->S = S;
  c->ordered = ordered;
  c->N = N_in;
  c->N_recv = N;
  c->bin_base = 0;
  c->groups = c->h_uoff[B];
  c->shuffled = true;
  c->compacted = false;
  st.bins = (uint32_t)B;
  st.sub_bins = S;
  st.big_bins = nbig;
  st.groups = c->groups;
  st.ms_exchange = 0;
  finish_stats(c, st);
  return MRHBM_OK;
}

// ---- several GPUs, exact layout: hist -> all-gather counts -> scatter (destination-major) -> NCCL all-to-all
//      -> sort+reduce of the owned partitions, each bin gathered from one segment per source
int shuffle_multi_exact(mrhbm_ctx* c, std::vector<Src>& live, uint64_t N, uint64_t N_in, mrhbm_stats& st) {
  const int G = c->world, me = c->rank;
  const uint32_t P = c->cfg.num_partitions;
  uint32_t all[8];
  int rc = gather_u32(c, (uint32_t)N, all);
  if (rc) return rc;
  uint64_t Nglobal = 0;
  for (int r = 0; r < G; r++) Nglobal += all[r];
  uint32_t S = pick_sub_bins(c, Nglobal);
  uint32_t nbig = 0, ordered = 1;
  cudaStream_t s = c->stream;
  uint64_t B = 0, Bl = 0, total_recv = 0;
  uint64_t send_off[9], send_cnt[8], recv_off[9], recv_cnt[8];
  ShuffleBuffers v{};
  for (int widen = 0;; widen++) {
    B = (uint64_t)P * S;
    Bl = (uint64_t)c->Pl * S;
    const uint32_t bin_base = c->pbase[me] * S;
    if (B >= (1ull << 31)) return fail(c, MRHBM_E_INVAL, "too many bins");
    rc = ensure_buffers(c, B, N);
    if (rc) return rc;
    rc = ensure_multi_buffers(c, B, Bl);
    if (rc) return rc;
    ordered = ((c->cfg.key_kind == MRHBM_KEY_U64 && !(c->cfg.flags & MRHBM_F_FORCE_RUNS) && widen == 0 && !c->no_ordered) || S == 1);
    c->sb.stride = 0;
    for (int attempt = 0;; attempt++) {
      st.attempts++;
      BinParams bp = make_bp(c, S, ordered);
      CU(c, cudaMemsetAsync(c->sb.hist, 0, (B << c->ctr_shift) * sizeof(uint32_t), s));
      CU(c, cudaMemsetAsync(c->sb.counters, 0, 8 * sizeof(uint32_t), s));
      CU(c, cudaEventRecord(c->ev[EV_COMBINE], s));
      for (auto& r : live) st.launches += launch_hist(c->rb, r.p, r.n, bp, c->sb.hist, s);
      CU(c, cudaEventRecord(c->ev[EV_HIST], s));
      // send layout (destination-major bins) + dense counts for the all-gather
      st.launches += launch_exscan(c->sb.hist, (uint32_t)B, c->sb.bin_off, c->sb.cursor, c->d_hd, 0xffffffffu, nullptr,
                                   nullptr, nullptr, c->ctr_shift, s, c->d_small + 64);
      rc = comm_allgather_u32(c->comm, c->d_hd, c->d_hall, B, s, &c->err);
      if (rc) return rc;
      // receive layout: per owned bin the total and the per-source offsets
      st.launches += launch_sum_src(c->d_hall, G, (uint32_t)B, bin_base, (uint32_t)Bl, c->d_tot, c->cap,
                                    c->sb.counters + CNT_GBIG, s);
      st.launches += launch_exscan(c->d_tot, (uint32_t)Bl, c->d_outoff, nullptr, nullptr, c->cap, c->sb.big_list,
                                   c->sb.counters + CNT_NBIG, c->sb.counters + CNT_TOTAL, 0, s, c->d_small + 64);
      st.launches += launch_exscan_rows(c->d_hall, G, (uint32_t)B, bin_base, (uint32_t)Bl, c->d_segoff, c->d_small + 32, s);
      for (int d = 0; d <= G; d++)
        CU(c, cudaMemcpyAsync(c->h_small + 48 + d, c->sb.bin_off + (uint64_t)c->pbase[d] * S, 4, cudaMemcpyDeviceToHost, s));
      CU(c, cudaMemcpyAsync(c->h_small + 32, c->d_small + 32, 4 * G, cudaMemcpyDeviceToHost, s));
      CU(c, cudaMemcpyAsync(c->h_counters, c->sb.counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
      CU(c, cudaEventRecord(c->ev[EV_PLAN], s));
      {
        // send buffer (destination-major exact layout) through the two-level coalesced split
        uint32_t F = 1;
        while ((uint64_t)F * F < B) F <<= 1;
        uint32_t C1 = (uint32_t)((B + F - 1) / F);
        if (B >= 2048 && F <= 1024 && C1 <= 1024) {
          rc = ensure_records(c, &c->l1buf, &c->l1_cap, N);
          if (rc) return rc;
          CU(c, cudaMemsetAsync(c->sb.hist, 0, ((uint64_t)C1 << c->ctr_shift) * sizeof(uint32_t), s));
          CU(c, cudaMemsetAsync(c->sb.cursor, 0, (B << c->ctr_shift) * sizeof(uint32_t), s));
          SplitPlan pl{};
          pl.B = (uint32_t)B;
          pl.F = F;
          pl.C1 = pl.C1_local = C1;
          pl.cap = c->cap;
          pl.cursor1 = c->sb.hist;
          pl.cursor = c->sb.cursor;
          pl.l1 = c->l1buf;
          pl.mid = c->sb.mid;
          pl.err_flags = c->sb.counters + CNT_ERR;
          pl.base_off = c->sb.bin_off;
          pl.ndest = 1;
          for (auto& r : live) st.launches += launch_split_l1(c->rb, r.p, r.n, bp, pl, s);
          st.launches += launch_split_l2(c->rb, bp, pl, s);
        } else {
          for (auto& r : live) st.launches += launch_scatter(c->rb, r.p, r.n, bp, c->sb.cursor, c->sb.mid, s);
        }
      }
      CU(c, cudaEventRecord(c->ev[EV_SCATTER], s));
      CU(c, cudaGetLastError());
      CU(c, cudaStreamSynchronize(s));
      nbig = c->h_counters[CNT_NBIG];
      total_recv = c->h_counters[CNT_TOTAL];
      // every rank counted the oversized bins of ALL ranks from the same all-gathered counts:
      // the decision is identical everywhere without another collective
      bool redo = c->h_counters[CNT_GBIG] && ordered && S > 1;
      if (redo) {
        ordered = 0;
        continue;
      }
      break;
    }
    send_off[0] = 0;
    recv_off[0] = 0;
    st.bytes_exchanged = 0;
    for (int d = 0; d < G; d++) {
      send_off[d] = (uint64_t)c->h_small[48 + d] * c->rb;
      send_cnt[d] = (uint64_t)(c->h_small[48 + d + 1] - c->h_small[48 + d]) * c->rb;
      recv_cnt[d] = (uint64_t)c->h_small[32 + d] * c->rb;
      recv_off[d + 1] = recv_off[d] + recv_cnt[d];
      if (d != me) st.bytes_exchanged += send_cnt[d];
    }
    rc = ensure_records(c, &c->recvbuf, &c->recv_cap, total_recv);
    if (rc) return rc;
    rc = ensure_out(c, total_recv);
    if (rc) return rc;
    if (nbig) {
      rc = ensure_records(c, &c->bigbuf, &c->big_cap, total_recv);
      if (rc) return rc;
    }
    // one grouped send/recv over NVLink replaces the GridFS / scp store-and-forward.  The rank's own
    // share never moves: the sort kernels read that segment in place from the send buffer.
    uint64_t net_send[8], net_recv[8];
    for (int d = 0; d < G; d++) {
      net_send[d] = d == me ? 0 : send_cnt[d];
      net_recv[d] = d == me ? 0 : recv_cnt[d];
    }
    rc = comm_alltoallv(c->comm, c->sb.mid, send_off, net_send, c->recvbuf, recv_off, net_recv, s, &c->err);
    if (rc) return rc;
    CU(c, cudaEventRecord(c->ev[EV_EXCH], s));
    v = c->sb;
    v.stride = 0;
    v.rep_shift = 0;
    v.bin_off = c->d_outoff;
    v.src = c->recvbuf;
    v.mid = nbig ? c->bigbuf : c->recvbuf;
    v.nseg = (uint32_t)G;
    set_range_hint(v, S, ordered);
    for (int r = 0; r < G; r++) {
      v.seg_off[r] = c->d_segoff + (uint64_t)r * (Bl + 1);
      v.seg_base[r] = recv_off[r] / c->rb;
    }
    {
      // own segment: record offset of (send buffer + send_off[me]) relative to v.src, possibly "negative"
      // (two's complement; the kernels add it to the source pointer with 64-bit wrap-around).  Both
      // buffers come from cudaMalloc (>= 256-byte aligned), so the difference is a multiple of rb.
      const int64_t delta = (int64_t)((const char*)c->sb.mid - (const char*)c->recvbuf) + (int64_t)send_off[me];
      v.seg_base[me] = (uint64_t)(delta / (int64_t)c->rb);
    }
    st.launches += launch_sort_reduce(c->rb, v, (uint32_t)Bl, c->cap, c->sm_count, s);
    CU(c, cudaEventRecord(c->ev[EV_SORT], s));
    st.launches += launch_big_bins(c->rb, v, nbig, c->cap, s);
    CU(c, cudaEventRecord(c->ev[EV_BIG], s));
    c->h_big.clear();
    if (c->sb.no_reduce && nbig) {  // group-only mode: the iterator merges the runs of an oversized bin
      c->h_big.resize(nbig);
      CU(c, cudaMemcpyAsync(c->h_big.data(), c->sb.big_list, nbig * 4, cudaMemcpyDeviceToHost, s));
    }
    st.launches += launch_exscan(c->sb.ucount, (uint32_t)Bl, c->sb.uoff, nullptr, nullptr, 0xffffffffu, nullptr, nullptr,
                                 c->sb.counters + CNT_TOTAL, 0, s, c->d_small + 64);
    c->h_bin_off.resize(Bl + 1);
    c->h_uoff.resize(Bl + 1);
    CU(c, cudaMemcpyAsync(c->h_bin_off.data(), c->d_outoff, (Bl + 1) * 4, cudaMemcpyDeviceToHost, s));
    CU(c, cudaMemcpyAsync(c->h_uoff.data(), c->sb.uoff, (Bl + 1) * 4, cudaMemcpyDeviceToHost, s));
    CU(c, cudaMemcpyAsync(c->h_counters, c->sb.counters, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU(c, cudaEventRecord(c->ev[EV_END], s));
    CU(c, cudaGetLastError());
    CU(c, cudaStreamSynchronize(s));
    uint32_t ef = c->h_counters[CNT_ERR];
    rc = gather_u32(c, ef | (c->long_jobs.empty() ? 0u : kErrfLongKeys), all);  // every rank must take the same branch
    if (rc) return rc;
    uint32_t ef_any = 0;
    for (int r = 0; r < G; r++) ef_any |= all[r];
    c->any_long = (ef_any & kErrfLongKeys) != 0;
    ef_any &= ~kErrfLongKeys;
    if (ef_any & ERRF_OVERFLOW) return fail(c, MRHBM_E_OVERFLOW, "u32 partial sum overflow while combining a hot key");
    if (ef_any & ERRF_SKEW) {
      if (widen < 4 && (uint64_t)P * S * 2 < (1ull << 31)) {
        S *= 2;
        continue;
      }
      return fail(c, MRHBM_E_SKEW, "a bin holds more distinct keys than one SM can sort");
    }
    break;
  }
  c->rv = v;
  c->B = (uint32_t)Bl;
  c->S = S;
  c->ordered = ordered;
  c->N = N_in;
  c->N_recv = total_recv;
  c->bin_base = c->pbase[me] * S;
  c->groups = c->h_uoff[Bl];
  c->shuffled = true;
  c->compacted = false;
  st.bins = (uint32_t)Bl;
  st.sub_bins = S;
  st.big_bins = nbig;
  st.groups = c->groups;
  st.ms_exchange = ev_ms(c, EV_SCATTER, EV_EXCH);
  finish_stats(c, st);
  c->stats.ms_exchange = st.ms_exchange;
  return MRHBM_OK;
}

// partition p -> local slot on this rank, or -1 when another rank owns it
inline int64_t slot_of(const mrhbm_ctx* c, uint32_t p) {
  return (p % (uint32_t)c->world) == (uint32_t)c->rank ? (int64_t)(p / (uint32_t)c->world) : -1;
}

}  // namespace

namespace {
// ---- keys longer than a record slot: partitioned, exchanged and grouped on the host --------------------------
std::string unescape_key(const std::string& esc) {
  std::string raw;
  raw.reserve(esc.size());
  for (size_t i = 0; i < esc.size(); i++) {
    if (esc[i] == 1 && i + 1 < esc.size())
      raw.push_back((char)(esc[++i] - 1));
    else
      raw.push_back(esc[i]);
  }
  return raw;
}
// the partition bin_of (mrhbm_dev.cuh) would give the key if it fitted a slot: FNV-in-doubles over the ORIGINAL bytes
// (examples/WordCount/partitionfn.lua:8-16), or the word hash over the little-endian words of the STORED bytes
uint32_t host_partition(const mrhbm_ctx* c, const std::string& esc) {
  const uint32_t P = c->cfg.num_partitions;
  if (c->cfg.partitioner == MRHBM_PART_FNV_LUA) {
    uint32_t h = 2166136261u;
    for (unsigned char b : unescape_key(esc)) h = fnv_lua_step(h, b);
    return h % P;
  }
  uint64_t h = 0x9E3779B97F4A7C15ull;
  for (size_t i = 0; i < esc.size(); i += 4) {
    uint32_t w = 0;
    memcpy(&w, esc.data() + i, std::min<size_t>(4, esc.size() - i));
    h = (h ^ w) * 0xBF58476D1CE4E5B9ull;
    h ^= h >> 29;
  }
  h ^= h >> 32;
  h *= 0x94D049BB133111EBull;
  return mulhi_u64_u32(h, P);
}

int build_long_groups(mrhbm_ctx* c) {
  c->long_groups.clear();
  c->long_group_count = 0;
  const int G = c->world;
  // this rank's committed long pairs as one blob: {u32 key bytes, u64 value, key bytes padded to 4}*
  std::vector<uint32_t> blob;
  for (auto& job : c->long_jobs)
    for (const LongPair& lp : job.second) {
      const size_t w0 = blob.size(), kw = (lp.esc.size() + 3) / 4;
      blob.resize(w0 + 3 + kw, 0u);
      blob[w0] = (uint32_t)lp.esc.size();
      memcpy(&blob[w0 + 1], &lp.value, 8);
      memcpy(&blob[w0 + 3], lp.esc.data(), lp.esc.size());
    }
  std::vector<uint32_t> all_blobs;
  std::vector<uint32_t> sizes(G, 0u);
  if (G == 1) {
    if (blob.empty()) return MRHBM_OK;
    sizes[0] = (uint32_t)blob.size();
    all_blobs.swap(blob);
  } else {  // COLLECTIVE, entered by all ranks or none (any_long travelled with the shuffle's error flags): sizes first,
            // then the blobs padded to the largest
    if (!c->any_long) return MRHBM_OK;
    if (blob.size() >= (1u << 28)) return fail(c, MRHBM_E_NOMEM, "more than 1 GB of long keys on one rank");
    int rc = gather_u32(c, (uint32_t)blob.size(), sizes.data());
    if (rc) return rc;
    uint32_t mx = 0;
    for (int r = 0; r < G; r++) mx = std::max(mx, sizes[r]);
    if (!mx) return MRHBM_OK;
    uint32_t *d_send = nullptr, *d_recv = nullptr;
    CU(c, cudaMalloc((void**)&d_send, (size_t)mx * 4));
    cudaError_t e = cudaMalloc((void**)&d_recv, (size_t)mx * 4 * G);
    if (e != cudaSuccess) {
      cudaFree(d_send);
      cudaGetLastError();
      return fail(c, MRHBM_E_NOMEM, "long keys: %s", cudaGetErrorString(e));
    }
    blob.resize(mx, 0u);
    all_blobs.resize((size_t)mx * G);
    e = cudaMemcpyAsync(d_send, blob.data(), (size_t)mx * 4, cudaMemcpyHostToDevice, c->stream);
    int rc2 = e == cudaSuccess ? comm_allgather_u32(c->comm, d_send, d_recv, mx, c->stream, &c->err) : 0;
    if (e == cudaSuccess && !rc2) e = cudaMemcpyAsync(all_blobs.data(), d_recv, (size_t)mx * 4 * G, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && !rc2) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_send);
    cudaFree(d_recv);
    if (rc2) return rc2;
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(c, MRHBM_E_CUDA, "long keys: %s", cudaGetErrorString(e));
    }
    // compact: rank r's blob starts at r * mx and holds sizes[r] words
    std::vector<uint32_t> packed;
    for (int r = 0; r < G; r++) packed.insert(packed.end(), all_blobs.begin() + (size_t)r * mx, all_blobs.begin() + (size_t)r * mx + sizes[r]);
    all_blobs.swap(packed);
  }
  // group the pairs of the partitions this rank owns (partition p lives on rank p % G, slot p / G)
  std::vector<std::map<std::string, std::vector<uint64_t>>> acc(c->Pl);
  const bool sum = c->cfg.reducer != MRHBM_RED_NONE;
  for (size_t w = 0; w + 3 <= all_blobs.size();) {
    const uint32_t klen = all_blobs[w];
    uint64_t v;
    memcpy(&v, &all_blobs[w + 1], 8);
    std::string esc((const char*)&all_blobs[w + 3], klen);
    w += 3 + (klen + 3) / 4;
    const uint32_t pid = host_partition(c, esc);
    if ((int)(pid % (uint32_t)G) != c->rank) continue;
    const uint32_t slot = pid / (uint32_t)G;
    if (slot >= c->Pl) continue;
    std::vector<uint64_t>& vals = acc[slot][esc];
    if (sum && !vals.empty())
      vals[0] += v;
    else
      vals.push_back(v);
  }
  c->long_groups.resize(c->Pl);
  for (uint32_t sl = 0; sl < c->Pl; sl++)
    for (auto& kv : acc[sl]) {  // std::map iterates in ascending bytewise key order (unsigned char_traits compare)
      c->long_groups[sl].push_back(LongGroup{kv.first, std::move(kv.second)});
      c->long_group_count++;
    }
  return MRHBM_OK;
}
}  // namespace

extern "C" {

int mrhbm_shuffle(mrhbm_ctx* c) {
  if (!c || !c->stream) return MRHBM_E_INVAL;
  Entry g(c);
  for (const Range& r : c->ranges)
    if (r.state == R_OPEN) return fail(c, MRHBM_E_INVAL, "map job '%s' is still open", r.job.c_str());
  invalidate(c);
  uint64_t N_in = 0, N = 0;
  for (auto& r : live_ranges(c)) N_in += r.second;
  if (N_in >= 0xfffffff0ull) return fail(c, MRHBM_E_INVAL, "more than 2^32 pairs on one GPU (%llu)", (unsigned long long)N_in);
  mrhbm_stats st{};
  st.pairs = N_in;
  std::vector<Src> live;
  CU(c, cudaEventRecord(c->ev[EV_START], c->stream));
  int rc = collect_sources(c, live, &N, st);
  if (rc) return rc;
  CU(c, cudaEventRecord(c->ev[EV_CSTART], c->stream));
  if (!c->no_optimistic && !(c->cfg.flags & MRHBM_F_NO_OPTIMISTIC) && !(c->tune & 1u)) {
    bool done = false;
    rc = shuffle_fast(c, live, N, N_in, st, &done);
    if (rc) return rc;
    if (done) return build_long_groups(c);
  }
  rc = c->world > 1 ? shuffle_multi_exact(c, live, N, N_in, st) : shuffle_single_exact(c, live, N, N_in, st);
  return rc ? rc : build_long_groups(c);
}

int mrhbm_stats_get(mrhbm_ctx* c, mrhbm_stats* out) {
  if (!c || !out) return MRHBM_E_INVAL;
  Entry g(c);
  *out = c->stats;
  return MRHBM_OK;
}

int mrhbm_partitions(mrhbm_ctx* c, uint32_t* ids, size_t cap, size_t* n) {
  if (!c || !n) return MRHBM_E_INVAL;
  Entry g(c);
  if (!c->shuffled) return fail(c, MRHBM_E_INVAL, "no shuffle result (call mrhbm_shuffle first)");
  size_t k = 0;
  for (uint32_t i = 0; i < c->Pl; i++) {
    const bool longs = i < c->long_groups.size() && !c->long_groups[i].empty();
    if (c->h_uoff[(uint64_t)(i + 1) * c->S] > c->h_uoff[(uint64_t)i * c->S] || longs) {
      if (ids && k < cap) ids[k] = (uint32_t)c->rank + i * (uint32_t)c->world;
      k++;
    }
  }
  *n = k;
  return MRHBM_OK;
}

int mrhbm_result_info_get(mrhbm_ctx* c, mrhbm_result_info* info) {
  if (!c || !info) return MRHBM_E_INVAL;
  Entry g(c);
  if (!c->shuffled) return fail(c, MRHBM_E_INVAL, "no shuffle result (call mrhbm_shuffle first)");
  size_t np = 0;
  mrhbm_partitions(c, nullptr, 0, &np);
  info->pairs_in = c->N;
  info->pairs_recv = c->N_recv;
  info->groups = c->groups;
  info->key_bytes = (uint32_t)c->kb;
  info->sorted = (c->ordered || c->S == 1) ? 1 : 0;
  info->runs_per_partition = c->S;
  info->partitions_nonempty = (uint32_t)np;
  return MRHBM_OK;
}

int mrhbm_result_copy(mrhbm_ctx* c, void* keys, uint64_t* sums, uint64_t* part_off) {
  if (!c) return MRHBM_E_INVAL;
  Entry g(c);
  int rc = ensure_compact(c);
  if (rc) return rc;
  if (c->long_group_count)
    return fail(c, MRHBM_E_KEY, "%llu groups have keys longer than the %d-byte slots of mrhbm_result_copy: iterate with mrhbm_groups_*",
                (unsigned long long)c->long_group_count, c->kb);
  if (keys && c->groups) CU(c, cudaMemcpyAsync(keys, c->ckeys, c->groups * c->kb, cudaMemcpyDeviceToHost, c->stream));
  if (sums && c->groups) CU(c, cudaMemcpyAsync(sums, c->csums, c->groups * 8, cudaMemcpyDeviceToHost, c->stream));
  if (part_off)
    for (uint32_t p = 0; p <= c->cfg.num_partitions; p++) {
      // groups of the owned partitions with id < p (other ranks' partitions are empty ranges here)
      uint64_t slots = p <= (uint32_t)c->rank ? 0 : ((uint64_t)p - c->rank + c->world - 1) / c->world;
      part_off[p] = c->h_uoff[std::min<uint64_t>(slots, c->Pl) * c->S];
    }
  CU(c, cudaStreamSynchronize(c->stream));
  return MRHBM_OK;
}

int mrhbm_checksum_input(mrhbm_ctx* c, uint64_t in[4]) {
  if (!c || !in) return MRHBM_E_INVAL;
  Entry g(c);
  CU(c, cudaMemsetAsync(c->d_acc, 0, 8 * sizeof(uint64_t), c->stream));
  for (auto& r : live_ranges(c)) launch_checksum_in(c->rb, (char*)c->pool + r.first * c->rb, r.second, c->d_acc, c->stream);
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(c->h_acc, c->d_acc, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  memcpy(in, c->h_acc, 4 * sizeof(uint64_t));
  return MRHBM_OK;
}

int mrhbm_checksum_result(mrhbm_ctx* c, uint64_t out[6]) {
  if (!c || !out) return MRHBM_E_INVAL;
  Entry g(c);
  if (!c->shuffled) return fail(c, MRHBM_E_INVAL, "no shuffle result (call mrhbm_shuffle first)");
  CU(c, cudaMemsetAsync(c->d_acc, 0, 8 * sizeof(uint64_t), c->stream));
  launch_checksum_out(c->rb, c->rv, c->B, make_bp(c, c->S, c->ordered), c->bin_base, c->d_acc, c->stream);
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(c->h_acc, c->d_acc, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  memcpy(out, c->h_acc, 6 * sizeof(uint64_t));
  return MRHBM_OK;
}

// ---------------------------------------------------------------------------
// reduce side iteration
// ---------------------------------------------------------------------------
static inline int slot_cmp(const mrhbm_ctx* c, const unsigned char* a, const unsigned char* b) {
  if (c->rb == 16) {
    uint64_t x, y;
    memcpy(&x, a, 8);
    memcpy(&y, b, 8);
    return (x > y) - (x < y);
  }
  return memcmp(a, b, c->kb);
}

// k-way merge of the partition's ascending runs (utils.lua:206-271 merges spill files the same way): a binary heap of
// run indices ordered by the key each run stands on (ties: lower run first)
static inline bool run_less(const mrhbm_iter* it, uint32_t a, uint32_t b) {
  const mrhbm_ctx* c = it->ctx;
  const int x = slot_cmp(c, &it->keys[it->runs[a].pos * c->kb], &it->keys[it->runs[b].pos * c->kb]);
  return x < 0 || (x == 0 && a < b);
}
static inline void heap_sift_down(mrhbm_iter* it, size_t i) {
  std::vector<uint32_t>& h = it->heap;
  const size_t n = h.size();
  for (;;) {
    const size_t l = 2 * i + 1, r = l + 1;
    size_t m = i;
    if (l < n && run_less(it, h[l], h[m])) m = l;
    if (r < n && run_less(it, h[r], h[m])) m = r;
    if (m == i) return;
    std::swap(h[i], h[m]);
    i = m;
  }
}
// the run at the top moved on (or ran dry): restore the heap
static inline void heap_fix_top(mrhbm_iter* it) {
  std::vector<uint32_t>& h = it->heap;
  if (it->runs[h[0]].pos >= it->runs[h[0]].end) {
    h[0] = h.back();
    h.pop_back();
  }
  if (!h.empty()) heap_sift_down(it, 0);
}

int mrhbm_groups_open(mrhbm_ctx* c, uint32_t part, mrhbm_iter** out) {
  if (!c || !out) return MRHBM_E_INVAL;
  Entry g(c);
  if (part >= c->cfg.num_partitions) return fail(c, MRHBM_E_INVAL, "partition %u out of range", part);
  int rc = ensure_compact(c);
  if (rc) return rc;
  int64_t slot = slot_of(c, part);
  if (slot < 0) return fail(c, MRHBM_E_INVAL, "partition %u is owned by rank %u", part, part % (uint32_t)c->world);
  mrhbm_iter* it = new mrhbm_iter();
  it->ctx = c;
  uint64_t b0 = (uint64_t)slot * c->S;
  uint64_t lo = c->h_uoff[b0], hi = c->h_uoff[b0 + c->S];
  it->base = lo;
  it->keys.resize((hi - lo) * c->kb);
  it->sums.resize(hi - lo);
  if (hi > lo) {
    cudaError_t e = cudaMemcpyAsync(it->keys.data(), (char*)c->ckeys + lo * c->kb, (hi - lo) * c->kb, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(it->sums.data(), c->csums + lo, (hi - lo) * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
      delete it;
      cudaGetLastError();
      return fail(c, MRHBM_E_CUDA, "groups_open: %s", cudaGetErrorString(e));
    }
  }
  if ((uint64_t)slot < c->long_groups.size() && !c->long_groups[(size_t)slot].empty()) it->longs = &c->long_groups[(size_t)slot];
  it->sorted = (c->ordered || c->S == 1) && c->h_big.empty();
  if (it->sorted) {
    it->runs.push_back(RunCursor{0, hi - lo});
  } else {
    for (uint32_t r = 0; r < c->S; r++) {
      uint64_t a = c->h_uoff[b0 + r] - lo, b = c->h_uoff[b0 + r + 1] - lo;
      if (b <= a) continue;
      bool big = std::find(c->h_big.begin(), c->h_big.end(), (uint32_t)(b0 + r)) != c->h_big.end();
      if (!big) {
        it->runs.push_back(RunCursor{a, b});
      } else {  // group-only mode: an oversized bin is a sequence of ascending runs of cap rows
        for (uint64_t q = a; q < b; q += c->cap) it->runs.push_back(RunCursor{q, std::min<uint64_t>(q + c->cap, b)});
      }
    }
  }
  for (uint32_t r = 0; r < it->runs.size(); r++)
    if (it->runs[r].pos < it->runs[r].end) it->heap.push_back(r);
  for (size_t i = it->heap.size() / 2; i-- > 0;) heap_sift_down(it, i);
  *out = it;
  return MRHBM_OK;
}

int mrhbm_groups_next(mrhbm_iter* it, const void** key, size_t* klen, const uint64_t** values, size_t* nvalues) {
  if (!it || !key || !klen || !values || !nvalues) return MRHBM_E_INVAL;
  mrhbm_ctx* c = it->ctx;
  // k-way merge over the ascending runs of the partition (a single run when sorted): the smallest key is on top
  const int best = it->heap.empty() ? -1 : (int)it->heap[0];
  if (it->longs && it->long_pos < it->longs->size()) {
    // a key longer than a slot comes next when it sorts before the smallest device key: both are compared as the
    // escaped byte strings a slot holds (the device key ends at its zero padding; a proper prefix sorts first)
    const LongGroup& lg = (*it->longs)[it->long_pos];
    bool take = best < 0;
    if (!take) {
      const unsigned char* dk = &it->keys[it->runs[best].pos * c->kb];
      const size_t dl = strnlen((const char*)dk, c->kb);
      const int cmp = memcmp(lg.esc.data(), dk, std::min(dl, lg.esc.size()));
      take = cmp < 0 || (cmp == 0 && lg.esc.size() < dl);
    }
    if (take) {
      it->long_pos++;
      const std::string raw = unescape_key(lg.esc);
      it->unesc.assign(raw.begin(), raw.end());
      *key = it->unesc.data();
      *klen = it->unesc.size();
      *values = lg.values.data();
      *nvalues = lg.values.size();
      return 1;
    }
  }
  if (best < 0) return 0;
  uint64_t i = it->runs[best].pos++;
  const unsigned char* k = &it->keys[i * c->kb];
  size_t nv = 1;
  bool gathered = false;
  if (c->cfg.reducer == MRHBM_RED_NONE) {
    // every row of this key: adjacent in its run, and -- for an oversized bin -- at the head of
    // sibling runs too
    it->valbuf.clear();
    it->valbuf.push_back(it->sums[i]);
    for (;;) {  // the run on top, then whichever run comes to the top with the same key
      RunCursor& rc = it->runs[it->heap[0]];
      while (rc.pos < rc.end && slot_cmp(c, &it->keys[rc.pos * c->kb], k) == 0) it->valbuf.push_back(it->sums[rc.pos++]);
      heap_fix_top(it);
      if (it->heap.empty()) break;
      const RunCursor& top = it->runs[it->heap[0]];
      if (slot_cmp(c, &it->keys[top.pos * c->kb], k) != 0) break;
    }
    nv = it->valbuf.size();
    gathered = true;
  } else {
    heap_fix_top(it);
  }
  if (c->rb == 16) {
    for (int b = 0; b < 8; b++) it->keybuf[b] = k[7 - b];  // 8-byte big-endian string
    *key = it->keybuf;
    *klen = 8;
  } else {
    size_t len = strnlen((const char*)k, c->kb);
    if (memchr(k, 1, len)) {  // escaped 0x00 / 0x01 bytes (see mrhbm_emit_str)
      it->unesc.clear();
      for (size_t i = 0; i < len; i++) {
        if (k[i] == 1 && i + 1 < len)
          it->unesc.push_back((unsigned char)(k[++i] - 1));
        else
          it->unesc.push_back(k[i]);
      }
      k = it->unesc.data();
      len = it->unesc.size();
    }
    *key = k;
    *klen = len;
  }
  *values = gathered ? it->valbuf.data() : &it->sums[i];
  *nvalues = nv;
  return 1;
}

void mrhbm_groups_close(mrhbm_iter* it) { delete it; }

// ---------------------------------------------------------------------------
// multi-GPU
// ---------------------------------------------------------------------------
int mrhbm_comm_unique_id(mrhbm_ctx* c, void* id) {
  if (!c || !id) return MRHBM_E_INVAL;
  return comm_unique_id(id, &c->err);
}
int mrhbm_comm_init(mrhbm_ctx* c, const void* id, int rank, int world) {
  if (!c || !id || world < 1 || rank < 0 || rank >= world) return MRHBM_E_INVAL;
  Entry g(c);
  if (world > 8) return fail(c, MRHBM_E_INVAL, "at most 8 ranks (one NVSwitch box)");
  if (c->world != 1) return fail(c, MRHBM_E_INVAL, "communicator already initialised");
  if (world == 1) return MRHBM_OK;
  int rc = comm_create(&c->comm, id, rank, world, c->dev, &c->err);
  if (rc) return rc;
  c->world = world;
  c->rank = rank;
  const uint32_t P = c->cfg.num_partitions;
  c->pbase[0] = 0;
  for (int r = 0; r < 8; r++) {
    uint32_t owned = r < world && (uint32_t)r < P ? (P - (uint32_t)r + (uint32_t)world - 1) / (uint32_t)world : 0;
    c->pbase[r + 1] = c->pbase[r] + owned;
  }
  c->Pl = c->pbase[rank + 1] - c->pbase[rank];
  CU(c, cudaMalloc((void**)&c->d_ipc, 16 * 9 * sizeof(uint32_t)));
  CU(c, cudaHostAlloc((void**)&c->h_ipc, 16 * 9 * sizeof(uint32_t), cudaHostAllocDefault));
  invalidate(c);
  return MRHBM_OK;
}

}  // extern "C"
