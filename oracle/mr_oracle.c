/*
 * mr_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See mr_oracle.h for scope and pinning.  Every function cites the reference
 * file:line it restates (pakozm/lua-mapreduce @ 767321e).
 *
 * The engine reproduces the reference's *data path*, not only its results:
 * map-side emit table -> keys_sorted -> combiner -> partitionfn -> one text
 * line per key ("return <k>,{v,...}\n") appended to a per-(partition,mapper)
 * spill file; reduce-side listing, line parsing, binary-heap k-way merge,
 * reducer algebra, result file; finalfn iteration order.  That makes it both
 * the parity checker and a reference-shaped CPU baseline.
 */
#define _GNU_SOURCE
#include "mr_oracle.h"
#include <ctype.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_MAP_RESULT 5000 /* mapreduce/utils.lua:53 */

/* ======================================================================== */
/* small utilities                                                          */
/* ======================================================================== */
typedef struct {
  char *p;
  size_t n, cap;
} sbuf_t;
static void sb_reserve(sbuf_t *b, size_t extra) {
  if (b->n + extra <= b->cap) return;
  size_t c = b->cap ? b->cap * 2 : 256;
  while (c < b->n + extra) c *= 2;
  b->p = (char *)realloc(b->p, c);
  b->cap = c;
}
static void sb_put(sbuf_t *b, const void *s, size_t n) {
  sb_reserve(b, n);
  memcpy(b->p + b->n, s, n);
  b->n += n;
}
static void sb_puts(sbuf_t *b, const char *s) { sb_put(b, s, strlen(s)); }

/* Lua string comparison in the C locale (lvm.c l_strcmp with strcoll == strcmp
 * per NUL-delimited chunk): unsigned bytewise, proper prefix first. SURVEY A.3 */
static int bytes_cmp(const void *a, size_t la, const void *b, size_t lb) {
  size_t m = la < lb ? la : lb;
  int c = m ? memcmp(a, b, m) : 0;
  if (c) return c;
  return (la > lb) - (la < lb);
}

/* ======================================================================== */
/* wire format  (mapreduce/utils.lua:100-120)                               */
/* ======================================================================== */
/* utils.lua:101-102: numbers go through tostring == "%.14g" (Lua 5.2 LUA_NUMBER_FMT) */
size_t mro_escape_num(double v, char *out) { return (size_t)sprintf(out, "%.14g", v); }

/* utils.lua:104-110: string.format("%q") (Lua 5.2 lstrlib.c addquoted) followed
 * by gsub("\\\n","\\n") */
size_t mro_escape_str(const void *sv, size_t len, char *out) {
  const unsigned char *s = (const unsigned char *)sv;
  char *o = out;
  *o++ = '"';
  for (size_t i = 0; i < len; i++) {
    unsigned char c = s[i];
    if (c == '"' || c == '\\') {
      *o++ = '\\';
      *o++ = (char)c;
    } else if (c == '\n') {
      /* %q emits backslash + newline; escape() rewrites that pair to \n */
      *o++ = '\\';
      *o++ = 'n';
    } else if (c == '\0' || iscntrl(c)) {
      int next_digit = (i + 1 < len) && isdigit(s[i + 1]);
      o += sprintf(o, next_digit ? "\\%03d" : "\\%d", (int)c);
    } else {
      *o++ = (char)c;
    }
  }
  *o++ = '"';
  return (size_t)(o - out);
}

/* ======================================================================== */
/* partitioners                                                             */
/* ======================================================================== */
/* examples/WordCount/partitionfn.lua:8-16 -- every operation in IEEE doubles,
 * Lua's a % b == a - floor(a/b)*b (exact here: b is a power of two). */
double mro_fnv_lua(const void *key, size_t len) {
  const unsigned char *k = (const unsigned char *)key;
  const double FNV_prime = 16777619.0, MAXV = 4294967296.0;
  double h = 2166136261.0;
  for (size_t i = 0; i < len; i++) {
    volatile double prod = h * FNV_prime; /* rounds to 53 bits above 2^53 */
    h = prod - floor(prod / MAXV) * MAXV;
    h = (double)(((uint32_t)h) ^ (uint32_t)k[i]); /* bit32.bxor */
  }
  return h;
}
uint32_t mro_part_fnv_lua(const void *key, size_t len, uint32_t nparts) {
  double h = mro_fnv_lua(key, len);
  return (uint32_t)(h - floor(h / (double)nparts) * (double)nparts);
}
static inline uint64_t mulhi64(uint64_t a, uint64_t b) {
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
}
/* SURVEY 8d config 4 */
uint32_t mro_part_mulhash(uint64_t key, uint32_t nparts) {
  uint64_t h = key * 0x9E3779B97F4A7C15ull;
  return (uint32_t)mulhi64(h, (uint64_t)nparts);
}
/* product built-in for fixed-slot string keys: little-endian u32 words of the
 * zero padded key while non-zero (keys hold no NUL, so this is length-exact) */
uint32_t mro_part_fnv64(const void *key, size_t len, uint32_t nparts) {
  const unsigned char *k = (const unsigned char *)key;
  uint64_t h = 0x9E3779B97F4A7C15ull;
  for (size_t i = 0; i < len; i += 4) {
    uint32_t w = 0;
    for (size_t j = 0; j < 4 && i + j < len; j++) w |= (uint32_t)k[i + j] << (8 * j);
    if (!w) break;
    h = (h ^ w) * 0xBF58476D1CE4E5B9ull;
    h ^= h >> 29;
  }
  h ^= h >> 32;
  h *= 0x94D049BB133111EBull;
  return (uint32_t)mulhi64(h, (uint64_t)nparts);
}
/* server.lua:134-145 */
int mro_count_digits(long n) {
  if (n == 0) return 1;
  int c = 0;
  while (n > 0) {
    n /= 10;
    c++;
  }
  return c;
}

/* ======================================================================== */
/* keys and values                                                          */
/* ======================================================================== */
typedef struct {
  int is_num;
  double num;
  char *s; /* owned */
  size_t len;
} okey_t;
static int key_cmp(const okey_t *a, const okey_t *b, int *err) {
  if (a->is_num != b->is_num) { /* Lua: attempt to compare number with string */
    if (err) *err = 1;
    return a->is_num ? -1 : 1;
  }
  if (a->is_num) return (a->num > b->num) - (a->num < b->num);
  return bytes_cmp(a->s, a->len, b->s, b->len);
}
typedef struct {
  double *v;
  size_t n, cap;
} vals_t;
static void vals_push(vals_t *a, double x) {
  if (a->n == a->cap) {
    a->cap = a->cap ? a->cap * 2 : 4;
    a->v = (double *)realloc(a->v, a->cap * sizeof(double));
  }
  a->v[a->n++] = x;
}

/* reducefn (examples/WordCount/reducefn.lua:1-5): count=0; count=count+v; emit(count).
 * The emitted list replaces values (job.lua:94-95,268-269). */
static void apply_reducer(int kind, vals_t *vals) {
  if (kind == MRO_RED_SUM) {
    double c = 0;
    for (size_t i = 0; i < vals->n; i++) c = c + vals->v[i];
    vals->n = 0;
    vals_push(vals, c);
  } /* IDENTITY: emits each value unchanged */
}

/* ======================================================================== */
/* heap.lua:29-93 -- array binary heap with user cmp                        */
/* ======================================================================== */
typedef struct {
  okey_t k;
  vals_t v;
  size_t which;
} hent_t;
typedef struct {
  hent_t *d; /* 1-based */
  size_t n, cap;
  int err;
} heap_t;
static int hent_less(heap_t *h, const hent_t *a, const hent_t *b) { /* utils.lua:214 */
  return key_cmp(&a->k, &b->k, &h->err) < 0;
}
static void heap_push(heap_t *h, hent_t v) { /* heap.lua:55-70 */
  if (h->n + 2 > h->cap) {
    h->cap = h->cap ? h->cap * 2 : 16;
    h->d = (hent_t *)realloc(h->d, h->cap * sizeof(hent_t));
  }
  size_t pos = ++h->n;
  while (pos > 1) {
    size_t p = pos / 2;
    if (hent_less(h, &v, &h->d[p])) {
      h->d[pos] = h->d[p];
      pos = p;
    } else
      break;
  }
  h->d[pos] = v;
}
static void heap_pop(heap_t *h) { /* heap.lua:33-53 */
  hent_t v = h->d[h->n];
  h->n--;
  if (h->n == 0) return;
  size_t pos = 1;
  for (;;) {
    size_t l = 2 * pos, r = 2 * pos + 1;
    if (l > h->n) break;
    size_t child = (r <= h->n && hent_less(h, &h->d[r], &h->d[l])) ? r : l;
    if (hent_less(h, &h->d[child], &v)) {
      h->d[pos] = h->d[child];
      pos = child;
    } else
      break;
  }
  h->d[pos] = v;
}
int mro_heap_sort(const double *in, size_t n, double *out) {
  heap_t h = {0};
  for (size_t i = 0; i < n; i++) {
    hent_t e;
    memset(&e, 0, sizeof e);
    e.k.is_num = 1;
    e.k.num = in[i];
    heap_push(&h, e);
  }
  for (size_t i = 0; i < n; i++) {
    out[i] = h.d[1].k.num;
    heap_pop(&h);
  }
  free(h.d);
  return 0;
}

/* ======================================================================== */
/* engine                                                                   */
/* ======================================================================== */
typedef struct {
  char *name;
  sbuf_t data;
  long part;
} ofile_t;
struct mro {
  int partitioner, combiner, reducer, aci;
  uint32_t nparts;
  ofile_t *files;
  size_t nfiles, capfiles;
  uint32_t *findex; /* open-addressing name -> file index + 1 (listing stays a table scan
                       only in spirit: GridFS answers name queries from an index too) */
  size_t findex_cap;
  size_t *part_first, *part_next; /* per-partition file chains built by mro_reduce_all */
  ofile_t *results;
  size_t nresults;
  pthread_mutex_t mu;
  char err[256];
};
/* open-addressing emit table (job.lua:83-97: result[key] = {values in order}) */
typedef struct {
  okey_t k;
  vals_t v;
  uint64_t h;
  int used;
} slot_t;
struct mro_map {
  mro_t *o;
  char *map_key;
  slot_t *tab;
  size_t cap, n;
};

mro_t *mro_new(int partitioner, uint32_t nparts, int combiner, int reducer, int aci) {
  mro_t *o = (mro_t *)calloc(1, sizeof *o);
  o->partitioner = partitioner;
  o->nparts = nparts;
  o->combiner = combiner;
  o->reducer = reducer;
  o->aci = aci;
  pthread_mutex_init(&o->mu, NULL);
  return o;
}
static void free_files(ofile_t *f, size_t n) {
  for (size_t i = 0; i < n; i++) {
    free(f[i].name);
    free(f[i].data.p);
  }
  free(f);
}
static uint64_t hash_bytes(const void *p, size_t n, uint64_t seed);
static void findex_rebuild(mro_t *o, size_t cap) {
  free(o->findex);
  o->findex_cap = cap;
  o->findex = (uint32_t *)calloc(cap, sizeof(uint32_t));
  for (size_t i = 0; i < o->nfiles; i++) {
    size_t h = hash_bytes(o->files[i].name, strlen(o->files[i].name), 7) & (cap - 1);
    while (o->findex[h]) h = (h + 1) & (cap - 1);
    o->findex[h] = (uint32_t)i + 1;
  }
}
/* find-or-create in the global spill table (caller holds o->mu) */
static ofile_t *global_file_get(mro_t *o, const char *name) {
  if (!o->findex_cap || (o->nfiles + 1) * 2 > o->findex_cap) findex_rebuild(o, o->findex_cap ? o->findex_cap * 2 : 1024);
  size_t cap = o->findex_cap, h = hash_bytes(name, strlen(name), 7) & (cap - 1);
  while (o->findex[h]) {
    ofile_t *f = &o->files[o->findex[h] - 1];
    if (strcmp(f->name, name) == 0) return f;
    h = (h + 1) & (cap - 1);
  }
  if (o->nfiles == o->capfiles) {
    o->capfiles = o->capfiles ? o->capfiles * 2 : 64;
    o->files = (ofile_t *)realloc(o->files, o->capfiles * sizeof(ofile_t));
  }
  ofile_t *f = &o->files[o->nfiles++];
  memset(f, 0, sizeof *f);
  f->name = strdup(name);
  o->findex[h] = (uint32_t)o->nfiles;
  return f;
}
void mro_free(mro_t *o) {
  if (!o) return;
  free(o->findex);
  free(o->part_first);
  free(o->part_next);
  free_files(o->files, o->nfiles);
  free_files(o->results, o->nresults);
  pthread_mutex_destroy(&o->mu);
  free(o);
}
const char *mro_error(const mro_t *o) { return o->err; }

static uint64_t hash_bytes(const void *p, size_t n, uint64_t seed) {
  const unsigned char *s = (const unsigned char *)p;
  uint64_t h = 1469598103934665603ull ^ seed;
  for (size_t i = 0; i < n; i++) h = (h ^ s[i]) * 1099511628211ull;
  return h ^ (h >> 29);
}
static uint64_t key_hash(const okey_t *k) {
  return k->is_num ? hash_bytes(&k->num, sizeof(double), 1) : hash_bytes(k->s, k->len, 0);
}
mro_map_t *mro_map_begin(mro_t *o, const char *map_key) {
  mro_map_t *m = (mro_map_t *)calloc(1, sizeof *m);
  m->o = o;
  m->map_key = strdup(map_key);
  m->cap = 1024;
  m->tab = (slot_t *)calloc(m->cap, sizeof(slot_t));
  return m;
}
static void map_free(mro_map_t *m) {
  for (size_t i = 0; i < m->cap; i++)
    if (m->tab[i].used) {
      free(m->tab[i].k.s);
      free(m->tab[i].v.v);
    }
  free(m->tab);
  free(m->map_key);
  free(m);
}
static slot_t *map_find(mro_map_t *m, const okey_t *k, uint64_t h) {
  size_t mask = m->cap - 1, i = h & mask;
  for (;;) {
    slot_t *s = &m->tab[i];
    if (!s->used) return s;
    if (s->h == h && s->k.is_num == k->is_num &&
        (k->is_num ? s->k.num == k->num
                   : (s->k.len == k->len && memcmp(s->k.s, k->s, k->len) == 0)))
      return s;
    i = (i + 1) & mask;
  }
}
static void map_grow(mro_map_t *m) {
  slot_t *old = m->tab;
  size_t oc = m->cap;
  m->cap *= 2;
  m->tab = (slot_t *)calloc(m->cap, sizeof(slot_t));
  for (size_t i = 0; i < oc; i++)
    if (old[i].used) *map_find(m, &old[i].k, old[i].h) = old[i];
  free(old);
}
/* job.lua:83-97 */
static int map_emit(mro_map_t *m, okey_t *k, double v) {
  uint64_t h = key_hash(k);
  slot_t *s = map_find(m, k, h);
  if (!s->used) {
    if ((m->n + 1) * 10 > m->cap * 7) {
      map_grow(m);
      s = map_find(m, k, h);
    }
    s->used = 1;
    s->h = h;
    s->k = *k;
    if (!k->is_num) {
      s->k.s = (char *)malloc(k->len ? k->len : 1);
      memcpy(s->k.s, k->s, k->len);
    }
    m->n++;
  }
  size_t N = s->v.n;        /* job.lua:89  local N = #result[key] */
  vals_push(&s->v, v);      /* job.lua:91 */
  if (m->o->combiner >= 0 && N > MAX_MAP_RESULT) /* job.lua:92-96 */
    apply_reducer(m->o->combiner, &s->v);
  return 0;
}
int mro_emit_str(mro_map_t *m, const void *key, size_t klen, double v) {
  okey_t k = {0, 0, (char *)key, klen};
  return map_emit(m, &k, v);
}
int mro_emit_num(mro_map_t *m, double key, double v) {
  okey_t k = {1, key, NULL, 0};
  return map_emit(m, &k, v);
}
void mro_map_abort(mro_map_t *m) { map_free(m); }

static int slot_cmp_err;
static int slot_ptr_cmp(const void *a, const void *b) {
  return key_cmp(&(*(slot_t *const *)a)->k, &(*(slot_t *const *)b)->k, &slot_cmp_err);
}
static void put_key(sbuf_t *b, const okey_t *k) {
  if (k->is_num) {
    sb_reserve(b, 40);
    b->n += mro_escape_num(k->num, b->p + b->n);
  } else {
    sb_reserve(b, 4 * k->len + 2);
    b->n += mro_escape_str(k->s, k->len, b->p + b->n);
  }
}
/* utils.lua:114-120 serialize_table_ipairs + job.lua:212-214 line assembly */
static void put_line(sbuf_t *b, const okey_t *k, const vals_t *v) {
  sb_puts(b, "return ");
  put_key(b, k);
  sb_puts(b, ",{");
  for (size_t i = 0; i < v->n; i++) {
    if (i) sb_puts(b, ",");
    sb_reserve(b, 40);
    b->n += mro_escape_num(v->v[i], b->p + b->n);
  }
  sb_puts(b, "}\n");
}
static long partition_of(mro_t *o, const okey_t *k) {
  if (k->is_num) { /* numeric keys only appear in fixtures: identity mod nparts */
    double p = k->num - floor(k->num / o->nparts) * o->nparts;
    return (long)p;
  }
  switch (o->partitioner) {
    case MRO_PART_FNV_LUA: return mro_part_fnv_lua(k->s, k->len, o->nparts);
    case MRO_PART_MULHASH: {
      uint64_t x = 0; /* key is an 8-byte big-endian string */
      for (size_t i = 0; i < k->len && i < 8; i++) x = (x << 8) | (unsigned char)k->s[i];
      return mro_part_mulhash(x, o->nparts);
    }
    default: return mro_part_fnv64(k->s, k->len, o->nparts);
  }
}
static ofile_t *file_get(ofile_t **files, size_t *n, size_t *cap, const char *name) {
  for (size_t i = 0; i < *n; i++)
    if (strcmp((*files)[i].name, name) == 0) return &(*files)[i];
  if (*n == *cap) {
    *cap = *cap ? *cap * 2 : 16;
    *files = (ofile_t *)realloc(*files, *cap * sizeof(ofile_t));
  }
  ofile_t *f = &(*files)[(*n)++];
  memset(f, 0, sizeof *f);
  f->name = strdup(name);
  return f;
}
/* job.lua:186-227 */
int mro_map_commit(mro_map_t *m) {
  mro_t *o = m->o;
  /* keys_sorted (utils.lua:123-128) */
  slot_t **keys = (slot_t **)malloc((m->n ? m->n : 1) * sizeof(slot_t *));
  size_t nk = 0;
  for (size_t i = 0; i < m->cap; i++)
    if (m->tab[i].used) keys[nk++] = &m->tab[i];
  slot_cmp_err = 0;
  qsort(keys, nk, sizeof(slot_t *), slot_ptr_cmp);
  if (slot_cmp_err) {
    snprintf(o->err, sizeof o->err, "attempt to compare number with string");
    free(keys);
    map_free(m);
    return -1;
  }
  /* builders keyed by "<results_ns>.P<part>.M<mapkey>" (job.lua:208-211); local first */
  ofile_t *local = NULL;
  size_t nl = 0, capl = 0;
  char name[512];
  long last_part = -1;
  ofile_t *last = NULL;
  for (size_t i = 0; i < nk; i++) {
    slot_t *s = keys[i];
    if (s->v.n > 1 && o->combiner >= 0) apply_reducer(o->combiner, &s->v); /* job.lua:198-202 */
    long part = partition_of(o, &s->k);                                   /* job.lua:203-207 */
    ofile_t *f;
    if (part == last_part && last)
      f = last;
    else {
      snprintf(name, sizeof name, "map_results.P%ld.M%s", part, m->map_key);
      f = file_get(&local, &nl, &capl, name);
      f->part = part;
      /* file_get may realloc: refresh cache */
      last = f;
      last_part = part;
    }
    put_line(&f->data, &s->k, &s->v);
  }
  free(keys);
  /* job.lua:217-221: remove_file + build == replace by name */
  pthread_mutex_lock(&o->mu);
  for (size_t i = 0; i < nl; i++) {
    ofile_t *g = global_file_get(o, local[i].name);
    free(g->data.p);
    g->data = local[i].data;
    g->part = local[i].part;
    free(local[i].name);
  }
  pthread_mutex_unlock(&o->mu);
  free(local);
  map_free(m);
  return 0;
}

size_t mro_nfiles(const mro_t *o) { return o->nfiles; }
const char *mro_file_name(const mro_t *o, size_t i) { return o->files[i].name; }
const char *mro_file_data(const mro_t *o, size_t i, size_t *len) {
  *len = o->files[i].data.n;
  return o->files[i].data.p;
}
int mro_add_file(mro_t *o, const char *name, const void *data, size_t len) {
  /* server.lua:305 parses P<part>.M<mapper> from the name */
  const char *p = strstr(name, ".P");
  if (!p) return -1;
  ofile_t *f = global_file_get(o, name);
  f->part = strtol(p + 2, NULL, 10);
  f->data.n = 0;
  sb_put(&f->data, data, len);
  return 0;
}

/* ---- line parser: the subset of Lua that escape()/serialize produce ------ */
static int parse_scalar(const char **pp, const char *end, okey_t *k, sbuf_t *tmp) {
  const char *p = *pp;
  if (p < end && *p == '"') {
    p++;
    tmp->n = 0;
    while (p < end && *p != '"') {
      if (*p == '\\') {
        p++;
        if (p >= end) return -1;
        if (*p == 'n') {
          sb_put(tmp, "\n", 1);
          p++;
        } else if (isdigit((unsigned char)*p)) {
          int v = 0, d = 0;
          while (d < 3 && p < end && isdigit((unsigned char)*p)) {
            v = v * 10 + (*p - '0');
            p++;
            d++;
          }
          char c = (char)v;
          sb_put(tmp, &c, 1);
        } else {
          sb_put(tmp, p, 1);
          p++;
        }
      } else {
        sb_put(tmp, p, 1);
        p++;
      }
    }
    if (p >= end) return -1;
    p++;
    k->is_num = 0;
    k->len = tmp->n;
    k->s = (char *)malloc(tmp->n ? tmp->n : 1);
    memcpy(k->s, tmp->p, tmp->n);
  } else {
    char *e;
    k->is_num = 1;
    k->num = strtod(p, &e);
    if (e == p) return -1;
    k->s = NULL;
    k->len = 0;
    p = e;
  }
  *pp = p;
  return 0;
}
/* utils.lua:222-224: load(line)() -> k, v */
static int parse_line(const char *line, size_t len, okey_t *k, vals_t *v, sbuf_t *tmp) {
  const char *p = line, *end = line + len;
  if (len < 7 || memcmp(p, "return ", 7)) return -1;
  p += 7;
  if (parse_scalar(&p, end, k, tmp)) return -1;
  if (end - p < 2 || p[0] != ',' || p[1] != '{') return -1;
  p += 2;
  v->n = 0;
  while (p < end && *p != '}') {
    char *e;
    double x = strtod(p, &e);
    if (e == p) return -1;
    vals_push(v, x);
    p = e;
    if (p < end && *p == ',') p++;
  }
  return (p < end && *p == '}') ? 0 : -1;
}

typedef struct {
  const char *p, *end;
} liter_t;
/* io.lines / gridfs_lines_iterator (utils.lua:133-200): next '\n'-terminated line */
static int next_line(liter_t *it, const char **line, size_t *len) {
  while (it->p < it->end) {
    const char *nl = (const char *)memchr(it->p, '\n', (size_t)(it->end - it->p));
    const char *e = nl ? nl : it->end;
    *line = it->p;
    *len = (size_t)(e - it->p);
    it->p = nl ? nl + 1 : it->end;
    if (*len) return 1; /* utils.lua:177-185 skips empty lines */
  }
  return 0;
}

static int name_cmp(const void *a, const void *b) {
  return strcmp((*(ofile_t *const *)a)->name, (*(ofile_t *const *)b)->name);
}
/* job.lua:230-296 for one partition; appends result lines to out */
static int reduce_partition(mro_t *o, long part, sbuf_t *out, char *err, size_t errn) {
  /* job.lua:255-260: files matching ^<path>/map_results.P<part>\..*  (listed sorted by name) */
  size_t nf = 0, cnt = 0;
  for (size_t i = o->part_first[part]; i != (size_t)-1; i = o->part_next[i]) cnt++;
  ofile_t **fl = (ofile_t **)malloc((cnt ? cnt : 1) * sizeof *fl);
  for (size_t i = o->part_first[part]; i != (size_t)-1; i = o->part_next[i]) fl[nf++] = &o->files[i];
  qsort(fl, nf, sizeof *fl, name_cmp);
  liter_t *its = (liter_t *)calloc(nf ? nf : 1, sizeof *its);
  heap_t h = {0};
  sbuf_t tmp = {0};
  int rc = 0;
#define TAKE_NEXT(w)                                                       \
  do { /* utils.lua:218-230 */                                             \
    const char *ln;                                                        \
    size_t ll;                                                             \
    if (its[w].p && next_line(&its[w], &ln, &ll)) {                        \
      hent_t e;                                                            \
      memset(&e, 0, sizeof e);                                             \
      if (parse_line(ln, ll, &e.k, &e.v, &tmp)) {                          \
        snprintf(err, errn, "Impossible to load line '%.*s' from '%s'",    \
                 (int)(ll > 80 ? 80 : ll), ln, fl[w]->name);               \
        rc = -1;                                                           \
      } else {                                                             \
        e.which = (w);                                                     \
        heap_push(&h, e);                                                  \
      }                                                                    \
    } else                                                                 \
      its[w].p = NULL;                                                     \
  } while (0)
  for (size_t i = 0; i < nf; i++) {
    its[i].p = fl[i]->data.p;
    its[i].end = fl[i]->data.p + fl[i]->data.n;
    if (!its[i].p) its[i].p = its[i].end = "";
  }
  for (size_t i = 0; i < nf && !rc; i++) TAKE_NEXT(i); /* utils.lua:249 */
  while (h.n && !rc) {
    /* merge_min_keys (utils.lua:232-247) */
    hent_t top = h.d[1];
    heap_pop(&h);
    TAKE_NEXT(top.which);
    while (h.n && !rc && key_cmp(&top.k, &h.d[1].k, &h.err) == 0) {
      hent_t aux = h.d[1];
      heap_pop(&h);
      TAKE_NEXT(aux.which);
      for (size_t j = 0; j < aux.v.n; j++) vals_push(&top.v, aux.v.v[j]);
      free(aux.k.s);
      free(aux.v.v);
    }
    if (h.err) {
      snprintf(err, errn, "attempt to compare number with string");
      rc = -1;
    }
    /* job.lua:264-284 */
    if (o->aci) {
      if (top.v.n > 1) apply_reducer(o->reducer, &top.v);
    } else
      apply_reducer(o->reducer, &top.v);
    put_line(out, &top.k, &top.v); /* job.lua:272-273 */
    free(top.k.s);
    free(top.v.v);
  }
  for (size_t i = 1; i <= h.n; i++) {
    free(h.d[i].k.s);
    free(h.d[i].v.v);
  }
  free(h.d);
  free(tmp.p);
  free(its);
  free(fl);
  return rc;
}
typedef struct {
  mro_t *o;
  size_t next;
  int rc;
  pthread_mutex_t mu;
} rctx_t;
static void *reduce_worker(void *arg) {
  rctx_t *c = (rctx_t *)arg;
  for (;;) {
    pthread_mutex_lock(&c->mu);
    size_t i = c->next++;
    pthread_mutex_unlock(&c->mu);
    if (i >= c->o->nresults) break;
    char err[256];
    if (reduce_partition(c->o, c->o->results[i].part, &c->o->results[i].data, err, sizeof err)) {
      pthread_mutex_lock(&c->mu);
      c->rc = -1;
      snprintf(c->o->err, sizeof c->o->err, "%s", err);
      pthread_mutex_unlock(&c->mu);
    }
  }
  return NULL;
}
static int long_cmp(const void *a, const void *b) {
  long x = *(const long *)a, y = *(const long *)b;
  return (x > y) - (x < y);
}
int mro_reduce_all(mro_t *o, int nthreads) {
  /* server.lua:300-324: one reduce job per distinct P<part> seen in file names;
   * result name "<result_ns>.P%0<digits>d" with digits from the max part key */
  free_files(o->results, o->nresults);
  o->results = NULL;
  o->nresults = 0;
  long *parts = (long *)malloc((o->nfiles ? o->nfiles : 1) * sizeof(long));
  size_t np = 0;
  long maxp = 0;
  for (size_t i = 0; i < o->nfiles; i++) parts[np++] = o->files[i].part;
  qsort(parts, np, sizeof(long), long_cmp);
  size_t u = 0;
  for (size_t i = 0; i < np; i++)
    if (i == 0 || parts[i] != parts[i - 1]) parts[u++] = parts[i];
  for (size_t i = 0; i < u; i++)
    if (parts[i] > maxp) maxp = parts[i];
  int digits = mro_count_digits(maxp);
  free(o->part_first);
  free(o->part_next);
  o->part_first = (size_t *)malloc((size_t)(maxp + 1) * sizeof(size_t));
  o->part_next = (size_t *)malloc((o->nfiles ? o->nfiles : 1) * sizeof(size_t));
  for (long q = 0; q <= maxp; q++) o->part_first[q] = (size_t)-1;
  for (size_t i = o->nfiles; i-- > 0;) {
    o->part_next[i] = o->part_first[o->files[i].part];
    o->part_first[o->files[i].part] = i;
  }
  o->results = (ofile_t *)calloc(u ? u : 1, sizeof(ofile_t));
  o->nresults = u;
  for (size_t i = 0; i < u; i++) {
    char name[64];
    snprintf(name, sizeof name, "result.P%0*ld", digits, parts[i]);
    o->results[i].name = strdup(name);
    o->results[i].part = parts[i];
  }
  free(parts);
  rctx_t c = {o, 0, 0, PTHREAD_MUTEX_INITIALIZER};
  if (nthreads <= 1)
    reduce_worker(&c);
  else {
    pthread_t th[64];
    if (nthreads > 64) nthreads = 64;
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, reduce_worker, &c);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  }
  /* job.lua:293: remove all map result files */
  if (!c.rc) {
    free_files(o->files, o->nfiles);
    o->files = NULL;
    o->nfiles = o->capfiles = 0;
    free(o->findex);
    o->findex = NULL;
    o->findex_cap = 0;
  }
  return c.rc ? -1 : (int)u;
}
size_t mro_nresults(const mro_t *o) { return o->nresults; }
const char *mro_result_name(const mro_t *o, size_t i) { return o->results[i].name; }
const char *mro_result_data(const mro_t *o, size_t i, size_t *len) {
  *len = o->results[i].data.n;
  return o->results[i].data.p;
}
long mro_result_part(const mro_t *o, size_t i) { return o->results[i].part; }

/* server.lua:360-385 pair_iterator */
struct mro_iter {
  mro_t *o;
  size_t k;
  liter_t it;
  okey_t key;
  vals_t vals;
  sbuf_t tmp;
};
mro_iter_t *mro_final_open(mro_t *o) {
  mro_iter_t *it = (mro_iter_t *)calloc(1, sizeof *it);
  it->o = o; /* results are already in sorted-name order (zero padded, server.lua:367) */
  return it;
}
int mro_final_next(mro_iter_t *it, int *key_is_num, double *key_num, const char **key,
                   size_t *klen, const double **vals, size_t *nvals, long *part) {
  const char *ln;
  size_t ll;
  for (;;) {
    if (it->it.p && next_line(&it->it, &ln, &ll)) break;
    if (it->k >= it->o->nresults) return 0;
    ofile_t *f = &it->o->results[it->k++];
    it->it.p = f->data.p ? f->data.p : "";
    it->it.end = it->it.p + f->data.n;
  }
  free(it->key.s);
  it->key.s = NULL;
  if (parse_line(ln, ll, &it->key, &it->vals, &it->tmp)) return -1;
  *key_is_num = it->key.is_num;
  *key_num = it->key.num;
  *key = it->key.s;
  *klen = it->key.len;
  *vals = it->vals.v;
  *nvals = it->vals.n;
  *part = it->o->results[it->k - 1].part;
  return 1;
}
void mro_final_close(mro_iter_t *it) {
  free(it->key.s);
  free(it->vals.v);
  free(it->tmp.p);
  free(it);
}

/* ======================================================================== */
/* naive word count (misc/naive.lua) and the WordCount mapfn tokeniser      */
/* ======================================================================== */
/* Lua %s == isspace() in the C locale: ' ' \t \n \v \f \r */
static inline int lua_isspace(unsigned char c) {
  return c == ' ' || (c >= '\t' && c <= '\r');
}
struct mro_naive {
  mro_map_t *m; /* reuse the emit table with no combiner */
  mro_t *o;
  size_t tokens;
  mro_wc_t *out;
  size_t nout;
};
mro_naive_t *mro_naive_new(void) {
  mro_naive_t *n = (mro_naive_t *)calloc(1, sizeof *n);
  n->o = mro_new(MRO_PART_FNV_LUA, 1, -1, MRO_RED_SUM, 0);
  n->m = mro_map_begin(n->o, "naive");
  return n;
}
void mro_naive_feed(mro_naive_t *n, const void *text, size_t len) {
  const unsigned char *s = (const unsigned char *)text;
  size_t i = 0;
  while (i < len) {
    while (i < len && lua_isspace(s[i])) i++;
    size_t b = i;
    while (i < len && !lua_isspace(s[i])) i++;
    if (i > b) { /* vocab[w] = (vocab[w] or 0) + 1 */
      okey_t k = {0, 0, (char *)s + b, i - b};
      uint64_t h = key_hash(&k);
      slot_t *sl = map_find(n->m, &k, h);
      if (sl->used)
        sl->v.v[0] += 1;
      else
        map_emit(n->m, &k, 1);
      n->tokens++;
    }
  }
}
static int wc_cmp(const void *a, const void *b) {
  const mro_wc_t *x = (const mro_wc_t *)a, *y = (const mro_wc_t *)b;
  return bytes_cmp(x->key, x->klen, y->key, y->klen);
}
size_t mro_naive_finish(mro_naive_t *n, const mro_wc_t **out) {
  free(n->out);
  n->out = (mro_wc_t *)malloc((n->m->n ? n->m->n : 1) * sizeof(mro_wc_t));
  n->nout = 0;
  for (size_t i = 0; i < n->m->cap; i++)
    if (n->m->tab[i].used) {
      mro_wc_t *w = &n->out[n->nout++];
      w->key = n->m->tab[i].k.s;
      w->klen = n->m->tab[i].k.len;
      w->count = n->m->tab[i].v.v[0];
    }
  qsort(n->out, n->nout, sizeof(mro_wc_t), wc_cmp);
  *out = n->out;
  return n->nout;
}
size_t mro_naive_tokens(const mro_naive_t *n) { return n->tokens; }
void mro_naive_free(mro_naive_t *n) {
  mro_map_abort(n->m);
  mro_free(n->o);
  free(n->out);
  free(n);
}
/* examples/WordCount/mapfn.lua:3-9: for each line, for w in line:gmatch("[^%s]+") emit(w,1) */
int mro_map_wordcount(mro_map_t *m, const void *text, size_t len) {
  const unsigned char *s = (const unsigned char *)text;
  size_t i = 0;
  while (i < len) {
    while (i < len && lua_isspace(s[i])) i++;
    size_t b = i;
    while (i < len && !lua_isspace(s[i])) i++;
    if (i > b) mro_emit_str(m, s + b, i - b, 1.0);
  }
  return 0;
}

/* ======================================================================== */
/* synthetic streams                                        */
/* ======================================================================== */
uint64_t mro_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
size_t mro_rank_to_key(uint64_t rank, char *out) {
  char tmp[16];
  int n = 0;
  uint64_t r = rank;
  while (r > 0) { /* bijective base 26 over a..z */
    r -= 1;
    tmp[n++] = (char)('a' + (r % 26));
    r /= 26;
  }
  size_t o = 0;
  while (n > 0) out[o++] = tmp[--n];
  uint64_t h = mro_splitmix64(rank ^ 0xA5A5A5A5A5A5A5A5ull);
  unsigned l = (unsigned)(h % 8);
  if (((h >> 8) % 64) == 0) l = 8 + (unsigned)((h >> 16) % 15);
  for (unsigned j = 0; j < l; j++) out[o++] = (char)('A' + (mro_splitmix64(h + j) % 26));
  return o;
}
uint64_t mro_zipf_rank(const uint64_t *T, uint64_t V, uint64_t u) {
  uint64_t lo = 0, hi = V; /* count of T[i] < u, i in [0,V) */
  while (lo < hi) {
    uint64_t mid = (lo + hi) / 2;
    if (T[mid] < u)
      lo = mid + 1;
    else
      hi = mid;
  }
  uint64_t r = 1 + lo;
  return r > V ? V : r;
}
void mro_gen_u64(uint64_t seed, uint64_t start, size_t n, uint64_t *keys, uint32_t *vals) {
  for (size_t i = 0; i < n; i++) {
    keys[i] = mro_splitmix64(seed + start + i);
    vals[i] = (uint32_t)(mro_splitmix64(seed + (1ull << 40) + start + i) >> 32);
  }
}
void mro_gen_zipf_rec32(uint64_t seed, uint64_t start, size_t n, const uint64_t *table,
                        uint64_t V, void *outv) {
  unsigned char *out = (unsigned char *)outv;
  for (size_t i = 0; i < n; i++) {
    uint64_t u = mro_splitmix64(seed + (1ull << 41) + start + i);
    uint64_t rank = mro_zipf_rank(table, V, u);
    unsigned char *r = out + 32 * i;
    memset(r, 0, 32);
    mro_rank_to_key(rank, (char *)r);
    uint32_t one = 1;
    memcpy(r + 28, &one, 4);
  }
}

/* ======================================================================== */
/* flat group-by oracles                                                    */
/* ======================================================================== */
typedef struct {
  uint32_t part;
  uint64_t val;
  uint64_t key;
} flat64_t;
static int flat64_cmp(const void *a, const void *b) {
  const flat64_t *x = (const flat64_t *)a, *y = (const flat64_t *)b;
  if (x->part != y->part) return x->part < y->part ? -1 : 1;
  return (x->key > y->key) - (x->key < y->key);
}
static size_t groupby_u64_any(const uint64_t *keys, const void *vals, int val_bytes, size_t n, int partitioner,
                              uint32_t nparts, uint64_t *out_keys, uint64_t *out_sums, uint64_t *part_off) {
  flat64_t *a = (flat64_t *)malloc((n ? n : 1) * sizeof *a);
  for (size_t i = 0; i < n; i++) {
    a[i].key = keys[i];
    a[i].val = val_bytes == 8 ? ((const uint64_t *)vals)[i] : ((const uint32_t *)vals)[i];
    if (partitioner == MRO_PART_MULHASH)
      a[i].part = mro_part_mulhash(keys[i], nparts);
    else {
      unsigned char be[8];
      for (int j = 0; j < 8; j++) be[j] = (unsigned char)(keys[i] >> (56 - 8 * j));
      a[i].part = partitioner == MRO_PART_FNV_LUA ? mro_part_fnv_lua(be, 8, nparts)
                                                  : mro_part_fnv64(be, 8, nparts);
    }
  }
  qsort(a, n, sizeof *a, flat64_cmp);
  size_t g = 0;
  memset(part_off, 0, (nparts + 1) * sizeof(uint64_t));
  for (size_t i = 0; i < n;) {
    size_t j = i;
    uint64_t s = 0;
    while (j < n && a[j].part == a[i].part && a[j].key == a[i].key) s += a[j++].val;
    if (s >= (1ull << 53)) { /* Lua sums are doubles */
      fprintf(stderr, "mr_oracle: sum exceeds 2^53\n");
      abort();
    }
    out_keys[g] = a[i].key;
    out_sums[g] = s;
    part_off[a[i].part + 1]++;
    g++;
    i = j;
  }
  for (uint32_t p = 0; p < nparts; p++) part_off[p + 1] += part_off[p];
  free(a);
  return g;
}
size_t mro_groupby_u64(const uint64_t *keys, const uint32_t *vals, size_t n, int partitioner,
                       uint32_t nparts, uint64_t *out_keys, uint64_t *out_sums,
                       uint64_t *part_off) {
  return groupby_u64_any(keys, vals, 4, n, partitioner, nparts, out_keys, out_sums, part_off);
}
size_t mro_groupby_u64_v64(const uint64_t *keys, const uint64_t *vals, size_t n, int partitioner,
                           uint32_t nparts, uint64_t *out_keys, uint64_t *out_sums,
                           uint64_t *part_off) {
  return groupby_u64_any(keys, vals, 8, n, partitioner, nparts, out_keys, out_sums, part_off);
}
static uint32_t g_rec_bytes;
typedef struct {
  uint32_t part;
  uint32_t idx;
} flatr_t;
static const unsigned char *g_recs;
static int flatr_cmp(const void *a, const void *b) {
  const flatr_t *x = (const flatr_t *)a, *y = (const flatr_t *)b;
  if (x->part != y->part) return x->part < y->part ? -1 : 1;
  /* zero padded slots without inner NULs: memcmp == bytewise, shorter first */
  return memcmp(g_recs + (size_t)x->idx * g_rec_bytes, g_recs + (size_t)y->idx * g_rec_bytes,
                g_rec_bytes - 4);
}
size_t mro_groupby_rec(const void *recs, size_t n, uint32_t rec_bytes, int partitioner,
                       uint32_t nparts, void *out_keys, uint64_t *out_sums, uint64_t *part_off) {
  const unsigned char *r = (const unsigned char *)recs;
  uint32_t kb = rec_bytes - 4;
  flatr_t *a = (flatr_t *)malloc((n ? n : 1) * sizeof *a);
  for (size_t i = 0; i < n; i++) {
    const unsigned char *k = r + i * rec_bytes;
    size_t len = strnlen((const char *)k, kb);
    a[i].idx = (uint32_t)i;
    a[i].part = partitioner == MRO_PART_FNV_LUA ? mro_part_fnv_lua(k, len, nparts)
                                                : mro_part_fnv64(k, len, nparts);
  }
  g_recs = r;
  g_rec_bytes = rec_bytes;
  qsort(a, n, sizeof *a, flatr_cmp);
  size_t g = 0;
  memset(part_off, 0, (nparts + 1) * sizeof(uint64_t));
  for (size_t i = 0; i < n;) {
    size_t j = i;
    uint64_t s = 0;
    const unsigned char *ki = r + (size_t)a[i].idx * rec_bytes;
    while (j < n && a[j].part == a[i].part &&
           memcmp(r + (size_t)a[j].idx * rec_bytes, ki, kb) == 0) {
      uint32_t v;
      memcpy(&v, r + (size_t)a[j].idx * rec_bytes + kb, 4);
      s += v;
      j++;
    }
    memcpy((unsigned char *)out_keys + g * kb, ki, kb);
    out_sums[g] = s;
    part_off[a[i].part + 1]++;
    g++;
    i = j;
  }
  for (uint32_t p = 0; p < nparts; p++) part_off[p + 1] += part_off[p];
  free(a);
  return g;
}

/* ======================================================================== */
/* job-size flat oracles (bench.py / tests: parity at the BASELINE.json sizes) */
/* ======================================================================== */
/* Same result semantics as mro_groupby_u64 / mro_groupby_rec -- per partition the distinct keys in
 * ascending bytewise order with the sum of their values -- computed for the synthetic streams of
 * SURVEY App. B with several threads, so that 10^8..10^9 pairs are checked in seconds.  A rank of a
 * multi-GPU job checks the partitions it owns (p % world == rank, server.lua:316-323 ids kept); the
 * other partitions come out empty, exactly like mrhbm_result_copy's part_off. */
typedef struct {
  uint64_t key, val;
} kv64_t;
static int kv64_cmp(const void *a, const void *b) {
  const kv64_t *x = (const kv64_t *)a, *y = (const kv64_t *)b;
  return (x->key > y->key) - (x->key < y->key);
}
typedef struct {
  int tid, nthreads, pass;
  uint64_t seed, start;
  size_t n;
  int partitioner;
  uint32_t nparts, world, rank;
  size_t *cnt;  /* [nthreads][nparts] counts, later write cursors */
  kv64_t *buf;
  size_t *pstart; /* [nparts+1] */
  size_t *groups; /* [nparts] */
  uint32_t *next_part;
  pthread_mutex_t *mu;
  uint64_t *out_keys, *out_sums;
  const size_t *gstart;
} gb64_t;
static uint32_t part_of_u64(int partitioner, uint64_t key, uint32_t nparts) {
  if (partitioner == MRO_PART_MULHASH) return mro_part_mulhash(key, nparts);
  unsigned char be[8];
  for (int j = 0; j < 8; j++) be[j] = (unsigned char)(key >> (56 - 8 * j));
  return partitioner == MRO_PART_FNV_LUA ? mro_part_fnv_lua(be, 8, nparts) : mro_part_fnv64(be, 8, nparts);
}
static void *gb64_worker(void *arg) {
  gb64_t *g = (gb64_t *)arg;
  size_t i0 = g->n / g->nthreads * g->tid, i1 = g->tid + 1 == g->nthreads ? g->n : g->n / g->nthreads * (g->tid + 1);
  size_t *cnt = g->cnt + (size_t)g->tid * g->nparts;
  if (g->pass == 0 || g->pass == 1) {
    for (size_t i = i0; i < i1; i++) {
      uint64_t k = mro_splitmix64(g->seed + g->start + i);
      uint32_t p = part_of_u64(g->partitioner, k, g->nparts);
      if (p % g->world != g->rank) continue;
      if (g->pass == 0) {
        cnt[p]++;
      } else {
        kv64_t *d = g->buf + cnt[p]++;
        d->key = k;
        d->val = (uint32_t)(mro_splitmix64(g->seed + (1ull << 40) + g->start + i) >> 32);
      }
    }
  } else if (g->pass == 2) { /* sort + reduce each partition in place */
    for (;;) {
      pthread_mutex_lock(g->mu);
      uint32_t p = (*g->next_part)++;
      pthread_mutex_unlock(g->mu);
      if (p >= g->nparts) break;
      kv64_t *a = g->buf + g->pstart[p];
      size_t m = g->pstart[p + 1] - g->pstart[p], w = 0;
      qsort(a, m, sizeof *a, kv64_cmp);
      for (size_t i = 0; i < m;) {
        size_t j = i;
        uint64_t sum = 0;
        while (j < m && a[j].key == a[i].key) sum += a[j++].val;
        if (sum >= (1ull << 53)) { /* Lua sums are doubles */
          fprintf(stderr, "mr_oracle: sum exceeds 2^53\n");
          abort();
        }
        a[w].key = a[i].key;
        a[w].val = sum;
        w++;
        i = j;
      }
      g->groups[p] = w;
    }
  } else { /* pass 3: compact */
    for (;;) {
      pthread_mutex_lock(g->mu);
      uint32_t p = (*g->next_part)++;
      pthread_mutex_unlock(g->mu);
      if (p >= g->nparts) break;
      const kv64_t *a = g->buf + g->pstart[p];
      size_t o = g->gstart[p];
      for (size_t i = 0; i < g->groups[p]; i++) {
        g->out_keys[o + i] = a[i].key;
        g->out_sums[o + i] = a[i].val;
      }
    }
  }
  return NULL;
}
static void gb64_run(gb64_t *proto, int nthreads, int pass) {
  pthread_t th[256];
  gb64_t args[256];
  for (int t = 0; t < nthreads; t++) {
    args[t] = *proto;
    args[t].tid = t;
    args[t].pass = pass;
    pthread_create(&th[t], NULL, gb64_worker, &args[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}
size_t mro_groupby_u64_stream(uint64_t seed, uint64_t start, size_t n, int partitioner, uint32_t nparts,
                              uint32_t world, uint32_t rank, int nthreads, uint64_t *out_keys,
                              uint64_t *out_sums, size_t out_cap, uint64_t *part_off) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (world < 1) world = 1;
  gb64_t g;
  memset(&g, 0, sizeof g);
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  uint32_t next = 0;
  g.nthreads = nthreads;
  g.seed = seed;
  g.start = start;
  g.n = n;
  g.partitioner = partitioner;
  g.nparts = nparts;
  g.world = world;
  g.rank = rank;
  g.mu = &mu;
  g.next_part = &next;
  g.cnt = (size_t *)calloc((size_t)nthreads * nparts, sizeof(size_t));
  g.pstart = (size_t *)calloc(nparts + 1, sizeof(size_t));
  g.groups = (size_t *)calloc(nparts, sizeof(size_t));
  gb64_run(&g, nthreads, 0);
  size_t run = 0; /* counts -> write cursors, partition-major then thread-major */
  for (uint32_t p = 0; p < nparts; p++) {
    g.pstart[p] = run;
    for (int t = 0; t < nthreads; t++) {
      size_t c = g.cnt[(size_t)t * nparts + p];
      g.cnt[(size_t)t * nparts + p] = run;
      run += c;
    }
  }
  g.pstart[nparts] = run;
  g.buf = (kv64_t *)malloc((run ? run : 1) * sizeof(kv64_t));
  gb64_run(&g, nthreads, 1);
  next = 0;
  gb64_run(&g, nthreads, 2);
  size_t *gstart = (size_t *)calloc(nparts + 1, sizeof(size_t));
  for (uint32_t p = 0; p < nparts; p++) gstart[p + 1] = gstart[p] + g.groups[p];
  size_t total = gstart[nparts];
  for (uint32_t p = 0; p <= nparts; p++) part_off[p] = gstart[p];
  if (total <= out_cap) {
    g.gstart = gstart;
    g.out_keys = out_keys;
    g.out_sums = out_sums;
    next = 0;
    gb64_run(&g, nthreads, 3);
  }
  free(gstart);
  free(g.buf);
  free(g.cnt);
  free(g.pstart);
  free(g.groups);
  return total; /* > out_cap: nothing was written */
}

/* occurrences of every Zipf rank in pairs [start, start+n) of the word stream: counts[r-1] for rank r */
typedef struct {
  int tid, nthreads;
  uint64_t seed, start;
  size_t n;
  const uint64_t *table;
  uint64_t V;
  uint32_t *priv;
} zc_t;
static void *zc_worker(void *arg) {
  zc_t *z = (zc_t *)arg;
  size_t i0 = z->n / z->nthreads * z->tid, i1 = z->tid + 1 == z->nthreads ? z->n : z->n / z->nthreads * (z->tid + 1);
  for (size_t i = i0; i < i1; i++) {
    uint64_t u = mro_splitmix64(z->seed + (1ull << 41) + z->start + i);
    z->priv[mro_zipf_rank(z->table, z->V, u) - 1]++;
  }
  return NULL;
}
int mro_zipf_counts(uint64_t seed, uint64_t start, size_t n, const uint64_t *table, uint64_t V, int nthreads,
                    uint64_t *counts) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (n / (size_t)nthreads >= 0xffffffffull) return -1; /* private counters are 32 bit */
  pthread_t th[256];
  zc_t args[256];
  for (int t = 0; t < nthreads; t++) {
    zc_t a = {t, nthreads, seed, start, n, table, V, (uint32_t *)calloc(V, sizeof(uint32_t))};
    args[t] = a;
    pthread_create(&th[t], NULL, zc_worker, &args[t]);
  }
  memset(counts, 0, V * sizeof(uint64_t));
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    for (uint64_t r = 0; r < V; r++) counts[r] += args[t].priv[r];
    free(args[t].priv);
  }
  return 0;
}
/* word count of a stream given its rank counts: keys = rank -> string (App. B), 28-byte zero padded slots */
typedef struct {
  uint32_t part;
  unsigned char key[28];
  uint64_t count;
} wcr_t;
static int wcr_cmp(const void *a, const void *b) {
  const wcr_t *x = (const wcr_t *)a, *y = (const wcr_t *)b;
  if (x->part != y->part) return x->part < y->part ? -1 : 1;
  return memcmp(x->key, y->key, 28);
}
size_t mro_wordcount_from_counts(const uint64_t *counts, uint64_t V, int partitioner, uint32_t nparts,
                                 uint32_t world, uint32_t rank, void *out_keys, uint64_t *out_sums,
                                 uint64_t *part_off) {
  if (world < 1) world = 1;
  wcr_t *a = (wcr_t *)malloc((V ? V : 1) * sizeof *a);
  size_t g = 0;
  for (uint64_t r = 1; r <= V; r++) {
    if (!counts[r - 1]) continue;
    wcr_t *e = a + g;
    memset(e->key, 0, 28);
    size_t len = mro_rank_to_key(r, (char *)e->key);
    e->part = partitioner == MRO_PART_FNV_LUA ? mro_part_fnv_lua(e->key, len, nparts) : mro_part_fnv64(e->key, len, nparts);
    if (e->part % world != rank) continue;
    e->count = counts[r - 1];
    g++;
  }
  qsort(a, g, sizeof *a, wcr_cmp);
  memset(part_off, 0, (nparts + 1) * sizeof(uint64_t));
  for (size_t i = 0; i < g; i++) {
    memcpy((unsigned char *)out_keys + 28 * i, a[i].key, 28);
    out_sums[i] = a[i].count;
    part_off[a[i].part + 1]++;
  }
  for (uint32_t p = 0; p < nparts; p++) part_off[p + 1] += part_off[p];
  free(a);
  return g;
}

/* ======================================================================== */
/* reference-shaped baseline runner                                         */
/* ======================================================================== */
#include <time.h>
static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
typedef struct {
  mro_t *o;
  int kind;
  uint64_t seed, start, per;
  uint32_t njobs, next;
  const uint64_t *table;
  uint64_t V;
  pthread_mutex_t mu;
} sctx_t;
static void *synthetic_worker(void *arg) {
  sctx_t *c = (sctx_t *)arg;
  for (;;) {
    pthread_mutex_lock(&c->mu);
    uint32_t j = c->next++;
    pthread_mutex_unlock(&c->mu);
    if (j >= c->njobs) break;
    char name[32];
    snprintf(name, sizeof name, "%u", j + 1);
    mro_map_t *m = mro_map_begin(c->o, name);
    uint64_t base = c->start + (uint64_t)j * c->per;
    for (uint64_t i = 0; i < c->per; i++) { /* mapfn: emit(k, v) per pair */
      if (c->kind == 0) {
        uint64_t k = mro_splitmix64(c->seed + base + i);
        uint32_t v = (uint32_t)(mro_splitmix64(c->seed + (1ull << 40) + base + i) >> 32);
        unsigned char be[8];
        for (int b = 0; b < 8; b++) be[b] = (unsigned char)(k >> (56 - 8 * b));
        mro_emit_str(m, be, 8, (double)v);
      } else {
        char key[32];
        uint64_t u = mro_splitmix64(c->seed + (1ull << 41) + base + i);
        size_t l = mro_rank_to_key(mro_zipf_rank(c->table, c->V, u), key);
        mro_emit_str(m, key, l, 1.0);
      }
    }
    mro_map_commit(m);
  }
  return NULL;
}
int mro_run_synthetic(mro_t *o, int kind, uint64_t seed, uint64_t start, uint64_t pairs_per_job,
                      uint32_t njobs, int nthreads, const uint64_t *table, uint64_t V,
                      double *map_seconds, double *reduce_seconds) {
  sctx_t c = {o, kind, seed, start, pairs_per_job, njobs, 0, table, V, PTHREAD_MUTEX_INITIALIZER};
  pthread_t th[64];
  if (nthreads > 64) nthreads = 64;
  if (nthreads < 1) nthreads = 1;
  double t0 = now_s();
  for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, synthetic_worker, &c);
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  double t1 = now_s();
  int r = mro_reduce_all(o, nthreads);
  double t2 = now_s();
  *map_seconds = t1 - t0;
  *reduce_seconds = t2 - t1;
  return r < 0 ? -1 : 0;
}

/* the same runner over a TEXT: njobs map jobs = the text cut at line ends into njobs slices (one "file" per
 * job, examples/WordCountBig/taskfn.lua:6-12), mapfn = examples/WordCount/mapfn.lua:3-9 */
typedef struct {
  mro_t *o;
  const char *text;
  const size_t *cut; /* njobs+1 slice boundaries */
  uint32_t njobs, next;
  pthread_mutex_t mu;
} tctx_t;
static void *text_worker(void *arg) {
  tctx_t *c = (tctx_t *)arg;
  for (;;) {
    pthread_mutex_lock(&c->mu);
    uint32_t j = c->next++;
    pthread_mutex_unlock(&c->mu);
    if (j >= c->njobs) break;
    char name[32];
    snprintf(name, sizeof name, "%u", j + 1);
    mro_map_t *m = mro_map_begin(c->o, name);
    mro_map_wordcount(m, c->text + c->cut[j], c->cut[j + 1] - c->cut[j]);
    mro_map_commit(m);
  }
  return NULL;
}
int mro_run_text(mro_t *o, const void *textv, size_t len, uint32_t njobs, int nthreads, double *map_seconds,
                 double *reduce_seconds) {
  const char *text = (const char *)textv;
  if (njobs < 1) njobs = 1;
  size_t *cut = (size_t *)calloc(njobs + 1, sizeof(size_t));
  for (uint32_t j = 1; j < njobs; j++) {
    size_t p = len / njobs * j;
    if (p < cut[j - 1]) p = cut[j - 1];
    while (p < len && p > 0 && text[p - 1] != '\n') p++; /* next line start */
    cut[j] = p;
  }
  cut[njobs] = len;
  tctx_t c = {o, text, cut, njobs, 0, PTHREAD_MUTEX_INITIALIZER};
  pthread_t th[64];
  if (nthreads > 64) nthreads = 64;
  if (nthreads < 1) nthreads = 1;
  double t0 = now_s();
  for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, text_worker, &c);
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  double t1 = now_s();
  int r = mro_reduce_all(o, nthreads);
  double t2 = now_s();
  free(cut);
  *map_seconds = t1 - t0;
  *reduce_seconds = t2 - t1;
  return r < 0 ? -1 : 0;
}

