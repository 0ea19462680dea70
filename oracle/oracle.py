"""ctypes face of the CPU oracle (oracle/mr_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg -- never from the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmr_oracle.so")

PART_FNV_LUA, PART_MULHASH, PART_FNV64 = 0, 1, 2
RED_SUM, RED_IDENTITY = 0, 1


def build(force=False):
    src = os.path.join(_HERE, "mr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _WC(C.Structure):
    _fields_ = [("key", C.c_void_p), ("klen", C.c_size_t), ("count", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, sz, u64, u32, dbl, i = C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_double, C.c_int

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("mro_escape_num", sz, dbl, C.c_char_p)
    sig("mro_escape_str", sz, C.c_char_p, sz, C.c_char_p)
    sig("mro_fnv_lua", dbl, C.c_char_p, sz)
    sig("mro_part_fnv_lua", u32, C.c_char_p, sz, u32)
    sig("mro_part_mulhash", u32, u64, u32)
    sig("mro_part_fnv64", u32, C.c_char_p, sz, u32)
    sig("mro_count_digits", i, C.c_long)
    sig("mro_heap_sort", i, vp, sz, vp)
    sig("mro_new", vp, i, u32, i, i, i)
    sig("mro_free", None, vp)
    sig("mro_error", C.c_char_p, vp)
    sig("mro_map_begin", vp, vp, C.c_char_p)
    sig("mro_emit_str", i, vp, C.c_char_p, sz, dbl)
    sig("mro_emit_num", i, vp, dbl, dbl)
    sig("mro_map_commit", i, vp)
    sig("mro_map_abort", None, vp)
    sig("mro_nfiles", sz, vp)
    sig("mro_file_name", C.c_char_p, vp, sz)
    sig("mro_file_data", vp, vp, sz, C.POINTER(sz))
    sig("mro_add_file", i, vp, C.c_char_p, C.c_char_p, sz)
    sig("mro_reduce_all", i, vp, i)
    sig("mro_nresults", sz, vp)
    sig("mro_result_name", C.c_char_p, vp, sz)
    sig("mro_result_data", vp, vp, sz, C.POINTER(sz))
    sig("mro_result_part", C.c_long, vp, sz)
    sig("mro_final_open", vp, vp)
    sig("mro_final_next", i, vp, C.POINTER(i), C.POINTER(dbl), C.POINTER(vp), C.POINTER(sz),
        C.POINTER(vp), C.POINTER(sz), C.POINTER(C.c_long))
    sig("mro_final_close", None, vp)
    sig("mro_naive_new", vp)
    sig("mro_naive_feed", None, vp, C.c_char_p, sz)
    sig("mro_naive_finish", sz, vp, C.POINTER(C.POINTER(_WC)))
    sig("mro_naive_tokens", sz, vp)
    sig("mro_naive_free", None, vp)
    sig("mro_map_wordcount", i, vp, C.c_char_p, sz)
    sig("mro_splitmix64", u64, u64)
    sig("mro_rank_to_key", sz, u64, C.c_char_p)
    sig("mro_zipf_rank", u64, vp, u64, u64)
    sig("mro_gen_u64", None, u64, u64, sz, vp, vp)
    sig("mro_gen_zipf_rec32", None, u64, u64, sz, vp, u64, vp)
    sig("mro_groupby_u64", sz, vp, vp, sz, i, u32, vp, vp, vp)
    sig("mro_groupby_u64_v64", sz, vp, vp, sz, i, u32, vp, vp, vp)
    sig("mro_groupby_rec", sz, vp, sz, u32, i, u32, vp, vp, vp)
    sig("mro_groupby_u64_stream", sz, u64, u64, sz, i, u32, u32, u32, i, vp, vp, sz, vp)
    sig("mro_zipf_counts", i, u64, u64, sz, vp, u64, i, vp)
    sig("mro_wordcount_from_counts", sz, vp, u64, i, u32, u32, u32, vp, vp, vp)
    sig("mro_run_text", i, vp, vp, sz, u32, i, C.POINTER(dbl), C.POINTER(dbl))
    sig("mro_run_synthetic", i, vp, i, u64, u64, u64, u32, i, vp, u64, C.POINTER(dbl), C.POINTER(dbl))
    _lib = L
    return L


# ---- scalar helpers --------------------------------------------------------
def escape(v):
    """mapreduce/utils.lua:100-112"""
    L = lib()
    if isinstance(v, (int, float)):
        buf = C.create_string_buffer(64)
        n = L.mro_escape_num(float(v), buf)
        return buf.raw[:n]
    b = bytes(v)
    buf = C.create_string_buffer(4 * len(b) + 8)
    n = L.mro_escape_str(b, len(b), buf)
    return buf.raw[:n]


def serialize_table_ipairs(vals):
    """mapreduce/utils.lua:114-120"""
    return b"{" + b",".join(escape(v) for v in vals) + b"}"


def fnv_lua(key: bytes) -> int:
    return int(lib().mro_fnv_lua(key, len(key)))


def part_fnv_lua(key: bytes, nparts=15) -> int:
    return lib().mro_part_fnv_lua(key, len(key), nparts)


def part_mulhash(key: int, nparts) -> int:
    return lib().mro_part_mulhash(key, nparts)


def part_fnv64(key: bytes, nparts) -> int:
    return lib().mro_part_fnv64(key, len(key), nparts)


def heap_sort(xs):
    a = np.asarray(xs, dtype=np.float64)
    out = np.empty_like(a)
    lib().mro_heap_sort(a.ctypes.data, a.size, out.ctypes.data)
    return out.tolist()


def splitmix64(x):
    return lib().mro_splitmix64(x & (2**64 - 1))


def rank_to_key(rank) -> bytes:
    buf = C.create_string_buffer(32)
    n = lib().mro_rank_to_key(rank, buf)
    return buf.raw[:n]


# ---- engine ----------------------------------------------------------------
class Engine:
    """One map/reduce task run through the reference-shaped CPU path."""

    def __init__(self, partitioner=PART_FNV_LUA, nparts=15, combiner=-1, reducer=RED_SUM, aci=True):
        self.L = lib()
        self.h = self.L.mro_new(partitioner, nparts, combiner, reducer, int(bool(aci)))

    def close(self):
        if self.h:
            self.L.mro_free(self.h)
            self.h = None

    __del__ = close

    def map_job(self, map_key, pairs=None, text=None):
        m = self.L.mro_map_begin(self.h, str(map_key).encode())
        if text is not None:
            self.L.mro_map_wordcount(m, text, len(text))
        for k, v in pairs or ():
            if isinstance(k, (int, float)):
                self.L.mro_emit_num(m, float(k), float(v))
            else:
                self.L.mro_emit_str(m, k, len(k), float(v))
        if self.L.mro_map_commit(m) != 0:
            raise RuntimeError(self.L.mro_error(self.h).decode())

    def add_file(self, name, data: bytes):
        assert self.L.mro_add_file(self.h, name.encode(), data, len(data)) == 0

    def files(self):
        out = {}
        for i in range(self.L.mro_nfiles(self.h)):
            n = C.c_size_t()
            p = self.L.mro_file_data(self.h, i, C.byref(n))
            out[self.L.mro_file_name(self.h, i).decode()] = C.string_at(p, n.value) if n.value else b""
        return out

    def reduce_all(self, nthreads=1):
        r = self.L.mro_reduce_all(self.h, nthreads)
        if r < 0:
            raise RuntimeError(self.L.mro_error(self.h).decode())
        return r

    def results(self):
        out = []
        for i in range(self.L.mro_nresults(self.h)):
            n = C.c_size_t()
            p = self.L.mro_result_data(self.h, i, C.byref(n))
            out.append((self.L.mro_result_name(self.h, i).decode(), self.L.mro_result_part(self.h, i),
                        C.string_at(p, n.value) if n.value else b""))
        return out

    def final_pairs(self):
        """finalfn's pairs_iterator: yields (part, key, [values]) in reference order."""
        it = self.L.mro_final_open(self.h)
        isnum, knum, kp, kl = C.c_int(), C.c_double(), C.c_void_p(), C.c_size_t()
        vp, vn, part = C.c_void_p(), C.c_size_t(), C.c_long()
        try:
            while True:
                r = self.L.mro_final_next(it, C.byref(isnum), C.byref(knum), C.byref(kp), C.byref(kl),
                                          C.byref(vp), C.byref(vn), C.byref(part))
                if r == 0:
                    return
                if r < 0:
                    raise RuntimeError("oracle: cannot parse result line")
                key = knum.value if isnum.value else (C.string_at(kp, kl.value) if kl.value else b"")
                vals = list((C.c_double * vn.value).from_address(vp.value)) if vn.value else []
                yield part.value, key, vals
        finally:
            self.L.mro_final_close(it)


def naive_wordcount(chunks):
    """misc/naive.lua: returns (tokens, [(key bytes, count)] sorted bytewise)."""
    L = lib()
    n = L.mro_naive_new()
    for c in chunks:
        L.mro_naive_feed(n, c, len(c))
    out = C.POINTER(_WC)()
    k = L.mro_naive_finish(n, C.byref(out))
    res = [(C.string_at(out[i].key, out[i].klen), int(out[i].count)) for i in range(k)]
    tokens = L.mro_naive_tokens(n)
    L.mro_naive_free(n)
    return tokens, res


# ---- synthetic streams + flat group-by --------------------------------------
def gen_u64(seed, start, n):
    keys = np.empty(n, dtype=np.uint64)
    vals = np.empty(n, dtype=np.uint32)
    lib().mro_gen_u64(seed, start, n, keys.ctypes.data, vals.ctypes.data)
    return keys, vals


def gen_zipf_rec32(seed, start, n, table):
    out = np.empty((n, 32), dtype=np.uint8)
    t = np.ascontiguousarray(table, dtype=np.uint64)
    lib().mro_gen_zipf_rec32(seed, start, n, t.ctypes.data, t.size, out.ctypes.data)
    return out


def groupby_u64(keys, vals, partitioner, nparts):
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    wide = np.asarray(vals).dtype.itemsize == 8  # 64-bit values (u64-key records carry them since ABI 2)
    vals = np.ascontiguousarray(vals, dtype=np.uint64 if wide else np.uint32)
    n = keys.size
    ok = np.empty(max(n, 1), dtype=np.uint64)
    os_ = np.empty(max(n, 1), dtype=np.uint64)
    po = np.empty(nparts + 1, dtype=np.uint64)
    fn = lib().mro_groupby_u64_v64 if wide else lib().mro_groupby_u64
    g = fn(keys.ctypes.data, vals.ctypes.data, n, partitioner, nparts, ok.ctypes.data, os_.ctypes.data, po.ctypes.data)
    return ok[:g].copy(), os_[:g].copy(), po


def groupby_rec(recs, partitioner, nparts):
    recs = np.ascontiguousarray(recs, dtype=np.uint8)
    n, rb = recs.shape
    okeys = np.empty((max(n, 1), rb - 4), dtype=np.uint8)
    os_ = np.empty(max(n, 1), dtype=np.uint64)
    po = np.empty(nparts + 1, dtype=np.uint64)
    g = lib().mro_groupby_rec(recs.ctypes.data, n, rb, partitioner, nparts, okeys.ctypes.data,
                              os_.ctypes.data, po.ctypes.data)
    return okeys[:g].copy(), os_[:g].copy(), po


def groupby_u64_stream(seed, start, n, partitioner, nparts, world=1, rank=0, nthreads=None, max_groups=None):
    """job-size oracle for the uniform u64 stream: (keys, sums, part_off) of the partitions rank owns"""
    nthreads = nthreads or min(os.cpu_count() or 1, 128)
    cap = int(max_groups if max_groups is not None else n)
    ok = np.empty(max(cap, 1), dtype=np.uint64)
    os_ = np.empty(max(cap, 1), dtype=np.uint64)
    po = np.empty(nparts + 1, dtype=np.uint64)
    g = lib().mro_groupby_u64_stream(seed, start, n, partitioner, nparts, world, rank, nthreads,
                                     ok.ctypes.data, os_.ctypes.data, cap, po.ctypes.data)
    if g > cap:
        raise RuntimeError("groupby_u64_stream: %d groups exceed the output capacity %d" % (g, cap))
    return ok[:g], os_[:g], po


def zipf_counts(seed, start, n, table, nthreads=None):
    """occurrences of every Zipf rank (index r-1) in pairs [start, start+n) of the word stream"""
    nthreads = nthreads or min(os.cpu_count() or 1, 128)
    t = np.ascontiguousarray(table, dtype=np.uint64)
    counts = np.zeros(t.size, dtype=np.uint64)
    if lib().mro_zipf_counts(seed, start, n, t.ctypes.data, t.size, nthreads, counts.ctypes.data) != 0:
        raise RuntimeError("zipf_counts: n too large")
    return counts


def wordcount_from_counts(counts, partitioner, nparts, world=1, rank=0):
    """(keys[g,28] uint8, sums, part_off) of the word count the rank counts mean, for the partitions rank owns"""
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    V = counts.size
    okeys = np.empty((max(V, 1), 28), dtype=np.uint8)
    os_ = np.empty(max(V, 1), dtype=np.uint64)
    po = np.empty(nparts + 1, dtype=np.uint64)
    g = lib().mro_wordcount_from_counts(counts.ctypes.data, V, partitioner, nparts, world, rank,
                                        okeys.ctypes.data, os_.ctypes.data, po.ctypes.data)
    return okeys[:g], os_[:g], po


def run_synthetic(engine, kind, seed, start, pairs_per_job, njobs, nthreads, table=None):
    """bench.py --impl reference: njobs map jobs + all reduce jobs on nthreads workers.
    Returns (map_seconds, reduce_seconds)."""
    t = np.ascontiguousarray(table, dtype=np.uint64) if table is not None else None
    ms, rs = C.c_double(), C.c_double()
    r = lib().mro_run_synthetic(engine.h, kind, seed, start, pairs_per_job, njobs, nthreads,
                                t.ctypes.data if t is not None else None, t.size if t is not None else 0,
                                C.byref(ms), C.byref(rs))
    if r != 0:
        raise RuntimeError(lib().mro_error(engine.h).decode())
    return ms.value, rs.value


def run_text(engine, text, njobs, nthreads):
    """word count of a text (bytes or uint8 array) as njobs map jobs + all reduce jobs on nthreads workers.
    Returns (map_seconds, reduce_seconds)."""
    a = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text, dtype=np.uint8)
    ms, rs = C.c_double(), C.c_double()
    if lib().mro_run_text(engine.h, a.ctypes.data, a.size, njobs, nthreads, C.byref(ms), C.byref(rs)) != 0:
        raise RuntimeError(lib().mro_error(engine.h).decode())
    return ms.value, rs.value
