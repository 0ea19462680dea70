"""Randomised cross-checks (hypothesis) between three independent statements of the same reference
semantics: the C oracle, the Python mirror's helpers and plain Python / numpy written here.
CPU only; sized to run in a few seconds."""
import os
import sys

import numpy as np
from hypothesis import given, settings, strategies as st

import mrhbm_loader

mrhbm_loader.load()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import oracle as O  # noqa: E402
from lua_mapreduce_b200.mapreduce import utils as mirror  # noqa: E402

keys_nonul = st.binary(min_size=0, max_size=40).map(lambda b: bytes(c for c in b if c))


def fnv_in_doubles(key: bytes) -> int:
    """examples/WordCount/partitionfn.lua:8-16 with Lua 5.2 numbers = Python floats (IEEE doubles)"""
    h = 2166136261.0
    for b in key:
        h = (h * 16777619.0) % 4294967296.0  # rounded to 53 bits BEFORE the modulo, like Lua
        h = float(int(h) ^ b)
    return int(h)


@settings(max_examples=300, deadline=None)
@given(keys_nonul)
def test_oracle_fnv_equals_lua_double_arithmetic(key):
    assert O.fnv_lua(key) == fnv_in_doubles(key)
    assert O.part_fnv_lua(key, 15) == fnv_in_doubles(key) % 15


@settings(max_examples=300, deadline=None)
@given(st.binary(min_size=0, max_size=60))
def test_mirror_escape_equals_oracle_escape_for_strings(s):
    assert mirror.escape(s) == O.escape(s)  # Lua 5.2 %q incl. \\ddd, \\n rewrite (utils.lua:100-112)


@settings(max_examples=300, deadline=None)
@given(st.one_of(st.integers(min_value=-2**53, max_value=2**53),
                 st.floats(allow_nan=False, allow_infinity=False, width=64)))
def test_mirror_escape_equals_oracle_escape_for_numbers(x):
    assert mirror.escape(x) == O.escape(float(x))  # "%.14g"


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 40), st.integers(0, 2**32 - 1)), min_size=0, max_size=400),
       st.sampled_from([1, 3, 16, 1024]))
def test_flat_groupby_equals_numpy(pairs, P):
    """mro_groupby_u64 (what the GPU parity tests compare against) vs numpy: per partition ascending keys,
    sums of equal keys"""
    keys = np.array([k * 0x9E3779B97F4A7C15 % 2**64 for k, _ in pairs], dtype=np.uint64)
    vals = np.array([v for _, v in pairs], dtype=np.uint32)
    ok, osum, po = O.groupby_u64(keys, vals, O.PART_MULHASH, P)
    part = np.array([O.part_mulhash(int(k), P) for k in keys], dtype=np.int64)
    want_k, want_s, want_po = [], [], [0]
    for p in range(P):
        sel = part == p
        uk = np.unique(keys[sel])
        want_k += uk.tolist()
        want_s += [int(vals[sel][keys[sel] == k].astype(np.uint64).sum()) for k in uk]
        want_po.append(len(want_k))
    assert ok.tolist() == want_k and osum.tolist() == want_s and po.tolist() == want_po
