"""Synthetic stream constants of SURVEY App. B shared by bench.py, tests and the device
generators (mrhbm_map_gen_u64 / mrhbm_map_gen_zipf).  Host side only builds the Zipf
threshold table -- the table bytes, not the formula, are the shared artefact."""
import numpy as np

SEED = 0x5EED20260921
ZIPF_V = 1 << 20
ZIPF_S = 1.1


def zipf_table(V=ZIPF_V, s=ZIPF_S):
    """T[r] = floor(2^64 * sum_{j<=r} j^-s / sum_{j<=V} j^-s), r=1..V-1; T[V] = 2^64-1.
    Ascending sequential summation in doubles.  Returned as uint64[V] (T[1] at index 0)."""
    w = np.arange(1, V + 1, dtype=np.float64) ** (-s)
    cum = np.cumsum(w)
    frac = cum / cum[-1]
    t = np.empty(V, dtype=np.uint64)
    scaled = np.floor(frac[:-1] * 18446744073709551616.0)
    t[:-1] = np.minimum(scaled, 18446744073709549568.0).astype(np.uint64)
    t[-1] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return t


def splitmix64_np(x):
    """vectorised splitmix64 over a uint64 array (wraparound arithmetic)."""
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        z = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))
