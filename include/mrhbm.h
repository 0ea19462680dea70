/*
 * mrhbm.h -- C ABI of the B200-native shuffle / sort / reduce core ("storage = hbm").
 *
 * This is the drop-in boundary for lua-mapreduce's map-side emit buffer and
 * reduce-side group-by.  Each entry point names the reference interface it
 * replaces (file:line in pakozm/lua-mapreduce @ 767321e).  The library is plain
 * C ABI: pointers and sizes only, no C++/torch types, no exception or CUDA
 * sticky error crosses it.  Every function that returns int returns >= 0 on
 * success and a negative MRHBM_E_* on failure; the message is available from
 * mrhbm_last_error() (luamongo's `nil, "<msg>"` convention,
 * external/luamongo/mongo_gridfilebuilder.cpp:64-69).
 *
 * Threading: a ctx and everything created from it is single-caller (the
 * reference's concurrency model is one single-threaded Lua worker process per
 * ctx, mapreduce/worker.lua:42-105).  One ctx drives one GPU; a multi-GPU job
 * is one ctx per rank joined by mrhbm_comm_init().
 *
 * There is NO CPU fallback: mrhbm_init() fails if no sm_100-class device is
 * usable.
 */
#ifndef MRHBM_H
#define MRHBM_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MRHBM_ABI_VERSION 2

enum { /* error codes */
  MRHBM_OK = 0,
  MRHBM_E_INVAL = -1,     /* bad argument / bad state */
  MRHBM_E_CUDA = -2,      /* CUDA runtime error (message has the CUDA string) */
  MRHBM_E_NOMEM = -3,     /* host or device allocation failed */
  MRHBM_E_KEY = -4,       /* key does not fit: a fixed-width record or result slot, a word of the device tokeniser, > 1 MB */
  MRHBM_E_SKEW = -5,      /* a bin holds more distinct keys than one SM can sort */
  MRHBM_E_OVERFLOW = -6,  /* u32 partial sum overflow while combining string-keyed records */
  MRHBM_E_NCCL = -7,      /* NCCL missing or failed */
  MRHBM_E_NODEVICE = -8   /* no usable sm_100 GPU: there is no CPU path */
};

enum { MRHBM_KEY_U64 = 0, MRHBM_KEY_STR = 1 };

/* built-in partitionfn (the user's Lua partitionfn, mapreduce/job.lua:203-207, is a
 * pure key -> integer function; these are the ones the device evaluates) */
enum {
  MRHBM_PART_FNV_LUA = 0,  /* examples/WordCount/partitionfn.lua:8-16, bit-exact incl. the
                              double-precision rounding; pid = h % num_partitions */
  MRHBM_PART_MULHASH = 1,  /* u64 keys: mulhi(key * 0x9E3779B97F4A7C15, P) (SURVEY 8d cfg 4) */
  MRHBM_PART_WORDHASH = 2  /* string keys: 64-bit word hash, mulhi(h, P) */
};
/* built-in reducefn/combinerfn (examples/WordCount/reducefn.lua:1-15: integer sum,
 * associative + commutative + idempotent flags set) */
enum {
  MRHBM_RED_SUM = 0,
  /* general (non-built-in) Lua reducefn, job.lua:264-284: the device only partitions, sorts and
   * groups; mrhbm_groups_next hands out every value of a key (order unspecified, like the
   * reference's heap-pop order among equal keys) and the host calls reducefn per group.
   * Needs combiner = 0.  mrhbm_result_copy then returns one row per pair (keys repeat).  A key may
   * carry any number of values, on one GPU or several: a bin larger than one shared-memory sort is
   * sorted run by run on its owner and the iterator merges the runs. */
  MRHBM_RED_NONE = 1
};

typedef struct mrhbm_ctx mrhbm_ctx;
typedef struct mrhbm_map mrhbm_map;
typedef struct mrhbm_iter mrhbm_iter;

typedef struct mrhbm_config {
  uint32_t struct_size;    /* = sizeof(mrhbm_config) */
  int32_t device;          /* CUDA ordinal; -1 = current device */
  uint32_t key_kind;       /* MRHBM_KEY_* */
  uint32_t max_key_bytes;  /* STR: <=27 -> 32 B records, <=59 -> 64 B, <=123 -> 128 B */
  uint32_t num_partitions; /* P >= 1 (reduce jobs; server.lua:316-323) */
  uint32_t partitioner;    /* MRHBM_PART_* */
  uint32_t reducer;        /* MRHBM_RED_* */
  uint32_t combiner;       /* 0 = none, 1 = combine with the reducer on the map side
                              (job.lua:92-96,198-202) */
  uint64_t reserve_pairs;  /* pre-size the HBM emit pool (0 = grow on demand) */
  uint32_t flags;          /* MRHBM_F_* */
  uint32_t reserved;
} mrhbm_config;
#define MRHBM_F_FORCE_RUNS 1u   /* always use hash sub-bins (skip the key-ordered attempt) */
#define MRHBM_F_SMALL_BINS 2u   /* test hook: tiny smem bins to exercise the overflow paths */
#define MRHBM_F_NO_OPTIMISTIC 4u /* always run the exact two-pass (histogram) partition layout */

/* record layouts (little endian) moved by emit_batch / gen / result_copy:
 *   U64 : { uint64_t key; uint64_t value; }                             16 B
 *   STR : { uint8_t key[RB-4] zero padded, no NUL inside (see mrhbm_emit_str); uint32_t value; }  RB = 32/64/128 */
uint32_t mrhbm_record_bytes(const mrhbm_ctx *);

/* ---- lifecycle (replaces cnn(...) + fs.router(...), mapreduce/fs.lua:185-208) ---- */
int mrhbm_init(const mrhbm_config *, mrhbm_ctx **out);
void mrhbm_destroy(mrhbm_ctx *);
const char *mrhbm_last_error(const mrhbm_ctx *); /* owned by ctx; valid until next call */
int mrhbm_abi_version(void);

/* pinned host memory for zero-staging emit_batch / result_copy */
void *mrhbm_host_alloc(mrhbm_ctx *, size_t bytes);
void mrhbm_host_free(mrhbm_ctx *, void *);

/* ---- map side: job.lua:83-97 (emit) + job.lua:186-227 (sort/partition/spill) ---- */
int mrhbm_map_begin(mrhbm_ctx *, const char *map_job_id, mrhbm_map **out);
/* key bytes are copied before return (Lua strings may be collected).  Any byte string is a key: bytes 0x00 and
 * 0x01 are stored escaped (01 01 / 01 02, order preserving; each costs one more byte of the key slot), hashed by
 * the built-in partitioners as the original bytes and handed back unescaped by mrhbm_groups_next.  Key slots in
 * emit_batch records and in mrhbm_result_copy are in the stored (escaped) form.
 * A key LONGER than the ctx record class (max_key_bytes) is accepted too (the reference takes any key,
 * utils.lua:104-110): such pairs are assumed rare, stay on the host, are partitioned (same partitioner), exchanged
 * and grouped at the barrier, and mrhbm_groups_next hands them out at their place in the key order.  They follow
 * commit / abort / replace-by-job-id like every other pair; the combiner does not see them; the fixed-width bulk
 * calls (emit_batch, map_wordcount, result_copy, checksums) do not carry them. */
int mrhbm_emit_str(mrhbm_map *, const void *key, size_t klen, uint32_t value);
int mrhbm_emit_u64(mrhbm_map *, uint64_t key, uint64_t value); /* value < 2^53 keeps Lua-number sums exact */
/* n records in the ctx layout.  Pageable memory is consumed before return; memory from
 * mrhbm_host_alloc() is read asynchronously and must stay untouched until commit/abort. */
int mrhbm_emit_batch(mrhbm_map *, const void *records, size_t n);
/* records already in HBM (device pointer on the ctx device), copied device-to-device */
int mrhbm_emit_device(mrhbm_map *, const void *dev_records, size_t n);
/* device-side mapfn for the synthetic streams of SURVEY App. B (bench + parity tests):
 * u64: key=splitmix64(seed+i), val=splitmix64(seed+2^40+i)>>32, i in [start,start+n)
 * zipf: rank by lower-bound search of splitmix64(seed+2^41+i) in table[V] (host pointer),
 *       key = rank->string, val = 1 (needs 32 B records) */
int mrhbm_map_gen_u64(mrhbm_map *, uint64_t seed, uint64_t start, uint64_t n);
int mrhbm_map_gen_zipf(mrhbm_map *, uint64_t seed, uint64_t start, uint64_t n,
                       const uint64_t *table, uint64_t V);
/* device-side WordCount mapfn (examples/WordCount/mapfn.lua:3-9, misc/naive.lua:2-5): emits
 * (word, 1) for every maximal run of non-space bytes of `text`; the space class is C-locale
 * isspace (' ' \t \n \v \f \r), i.e. Lua's "[^%s]+".  A 64 MB piece of the text that holds a word longer than
 * the ctx record class (a URL in a corpus) is tokenised on the host instead, the long words going the way of
 * mrhbm_emit_str's long keys.  On failure nothing of the call is emitted.  *words (optional) receives the token count. */
int mrhbm_map_wordcount(mrhbm_map *, const void *text, size_t len, uint64_t *words);
/* host-side generator of the synthetic word-count text of SURVEY App. B (bench + tests; needs no ctx):
 * words first .. first+n-1 of the Zipf(table) word stream, one space between words, '\n' after every
 * words_per_line-th word.  *len receives the bytes the text needs; MRHBM_E_INVAL (nothing usable written) if
 * that exceeds cap.  Call with cap = 0 to size the buffer. */
int mrhbm_synth_zipf_text(uint64_t seed, uint64_t first, uint64_t n, uint32_t words_per_line,
                          const uint64_t *table, uint64_t V, void *out, size_t cap, size_t *len, int threads);
/* copies n committed pairs starting at pair index `first` (commit order) back to host
 * memory in the record layout (bench + tests: checks the device generators) */
int mrhbm_pool_read(mrhbm_ctx *, uint64_t first, uint64_t n, void *host_out);
/* atomic publish; a second commit under the same job id REPLACES the first
 * (job.lua:217-221 remove_file + build).  Consumes the handle. */
int mrhbm_map_commit(mrhbm_map *);
/* BROKEN job: nothing becomes visible (worker.lua:120-127).  Consumes the handle. */
void mrhbm_map_abort(mrhbm_map *);

/* ---- barrier between MAP and REDUCE (server.lua:279-329) ---- */
/* hash-partition [+ NCCL all-to-all] + sort + segmented reduce of everything committed */
int mrhbm_shuffle(mrhbm_ctx *);
/* NON-EMPTY partitions owned by this rank, ascending -> red_jobs (server.lua:300-324) */
int mrhbm_partitions(mrhbm_ctx *, uint32_t *ids, size_t cap, size_t *n);

/* ---- reduce side: utils.merge_iterator consumer (utils.lua:206-271, job.lua:264-284)
 *      and finalfn's pair iterator (server.lua:360-385) ---- */
int mrhbm_groups_open(mrhbm_ctx *, uint32_t partition, mrhbm_iter **out);
/* 1 = one group, 0 = end, <0 = error.  Ascending key order (C-locale bytewise; u64
 * numeric).  *key points at klen key bytes (u64 keys: 8 bytes big endian);
 * values/nvalues is the reduced list (built-in sum: one value; MRHBM_RED_NONE: all values of the
 * key).  Pointers stay valid
 * until the next call on this iterator. */
int mrhbm_groups_next(mrhbm_iter *, const void **key, size_t *klen, const uint64_t **values,
                      size_t *nvalues);
void mrhbm_groups_close(mrhbm_iter *);

/* ---- bulk result access (finalfn feed without per-group calls) ---- */
typedef struct mrhbm_result_info {
  uint64_t pairs_in;   /* pairs committed on this rank */
  uint64_t pairs_recv; /* pairs this rank reduced (after the exchange) */
  uint64_t groups;     /* distinct keys owned by this rank */
  uint32_t key_bytes;  /* bytes per key slot in result_copy (8 or RB-4) */
  uint32_t sorted;     /* 1: each partition is one ascending run on the device;
                          0: a partition is several ascending runs, merged by the iterator */
  uint32_t runs_per_partition;
  uint32_t partitions_nonempty;
} mrhbm_result_info;
/* `groups` counts the rows mrhbm_result_copy returns; groups whose key is longer than a slot (mrhbm_emit_str) are not
 * among them -- mrhbm_partitions and mrhbm_groups_* see both kinds. */
int mrhbm_result_info_get(mrhbm_ctx *, mrhbm_result_info *);
/* copies all groups of this rank, partition-major, to host: keys (groups*key_bytes; u64
 * keys native little endian) and sums (groups*8); part_off[P+1] receives group offsets.
 * Inside a partition the order is ascending iff info.sorted, else run-major.
 * MRHBM_E_KEY when the result holds keys longer than a slot (see mrhbm_emit_str): iterate instead. */
int mrhbm_result_copy(mrhbm_ctx *, void *keys, uint64_t *sums, uint64_t *part_off);
/* size-independent parity properties computed on the device:
 *  in[0..3]  = { sum f1(key)*v, sum f2(key)*v, sum v, pairs }       over committed pairs
 *  out[0..3] = { sum f1(key)*s, sum f2(key)*s, sum s, groups }      over reduced groups
 *  out[4] = adjacent keys not strictly ascending inside a run, out[5] = groups whose
 *  partitioner(key) differs from the partition they sit in.  All mod 2^64.
 * Linearity of the sum makes in[0..2] == out[0..2] for a correct group-by. */
int mrhbm_checksum_input(mrhbm_ctx *, uint64_t in[4]);
int mrhbm_checksum_result(mrhbm_ctx *, uint64_t out[6]);

/* ---- measurement ---- */
typedef struct mrhbm_stats {
  float ms_total;     /* last shuffle, CUDA events on the ctx stream */
  float ms_combine, ms_hist, ms_plan, ms_scatter, ms_exchange, ms_sort_reduce, ms_bigbins;
  uint32_t launches;  /* kernels launched by the last shuffle */
  uint32_t bins, big_bins, sub_bins, attempts;
  uint64_t pairs, groups, bytes_exchanged;
  float ms_setup;     /* key sampling, job-wide agreement, buffer set-up, counter clears (between combine and hist/level 1) */
  float ms_finish;    /* offsets of the groups, totals, error flags back to the host (after the sort) */
} mrhbm_stats;
int mrhbm_stats_get(mrhbm_ctx *, mrhbm_stats *);
/* drops committed pairs and results, keeps buffers (next iteration of a "loop" task,
 * server.lua:386-404) */
int mrhbm_reset(mrhbm_ctx *);

/* ---- multi-GPU: one ctx per rank, partition p is owned by rank p % world ---- */
#define MRHBM_UNIQUE_ID_BYTES 128
int mrhbm_comm_unique_id(mrhbm_ctx *, void *id /* MRHBM_UNIQUE_ID_BYTES, rank 0 */);
int mrhbm_comm_init(mrhbm_ctx *, const void *id, int rank, int world);

#ifdef __cplusplus
}
#endif
#endif
