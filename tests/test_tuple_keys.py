"""Composite (tuple) keys (mapreduce/tuple.lua): the order-preserving key codec
against the restated reference order (tuple.lua:183-195), the tuple() constructor semantics
(tuple.lua:252-301, utest 309-327) and the host flow over the stand-in ctx."""
import itertools
import random

import pytest

import mrhbm_loader

mrhbm_loader.load()
from lua_mapreduce_b200.mapreduce import tuple as T  # noqa: E402


def test_constructor_semantics():  # tuple.lua:309-327
    assert T.tuple_(1) == 1 and T.tuple_("a") == "a"          # scalars pass through unchanged
    a, b = T.tuple_(2, [4, 5], "a"), T.tuple_(2, [4, 5], "a")
    assert a == b and hash(a) == hash(b) and a == (2, (4, 5), "a")  # interning == equality by value
    assert T.tuple_([1, 2]) == (1, 2) and T.tuple_((1, 2)) == (1, 2)
    assert {a: 1}[b] == 1


def rand_component(rng, kind):
    if kind == "n":
        return rng.choice([0, 1, -1, 2, 3.5, -2.25, 1e9, -1e9, 2**53, -(2**53), 1e-300, -1e-300, 7, 8, 255, 256, float("inf")])
    return bytes(rng.choice(b"abcz\x02\xff") for _ in range(rng.randint(0, 4)))


def rand_tuple(rng, shape):
    return tuple(rand_component(rng, k) for k in shape)


def test_encoding_is_a_linear_extension_of_the_reference_order():
    rng = random.Random(7)
    shapes = ["n", "s", "nn", "ns", "sn", "ss", "nns", "sss"]
    keys = {rand_tuple(rng, sh) for sh in shapes for _ in range(60)}
    enc = {k: T.encode(k) for k in keys}
    assert len(set(enc.values())) == len(keys)                      # injective
    assert all(b"\0" not in e and len(e) <= T.MAX_KEY_BYTES for e in enc.values())
    for a, b in itertools.combinations(keys, 2):
        same_shape = len(a) == len(b) and all(type(x) is type(y) or {type(x), type(y)} <= {int, float} for x, y in zip(a, b))
        if len(a) != len(b) or same_shape:                           # Lua cannot compare a number with a string
            if T.lt(a, b):
                assert enc[a] < enc[b], (a, b)
            if T.lt(b, a):
                assert enc[b] < enc[a], (a, b)
    for k, e in enc.items():                                         # round trip
        assert T.decode(e) == k


def test_component_order_details():
    e = T.encode
    assert e((1,)) < e((2,)) < e((1, 1))                              # shorter tuples first
    assert e((-2.5,)) < e((-1,)) < e((0,)) == e((-0.0,)) < e((1e-300,)) < e((1,)) < e((float("inf"),))
    assert e((b"ab",)) < e((b"abc",)) < e((b"b",))                   # a proper prefix sorts first (C-locale bytewise)
    assert e(5) < e(6) and T.decode(e(5)) == 5 and T.decode(e(b"w")) == b"w"   # scalar keys
    assert e((1, (2, 3))) < e((1, (2, 4))) and T.decode(e((1, (2, b"x")))) == (1, (2, b"x"))
    assert e(("text", 3)) == e((b"text", 3))
    with pytest.raises(ValueError):
        e((b"a\0b",))
    with pytest.raises(ValueError):
        e((float("nan"),))
    with pytest.raises(ValueError):
        e(tuple(range(40)))
    with pytest.raises(ValueError):
        e((b"x" * 200,))
    with pytest.raises(OverflowError):
        e((2**60,))


# ---- host flow: a bigram count with tuple keys through server:loop() over the stand-in ctx
import sys
import types

from lua_mapreduce_b200.mapreduce import server  # noqa: E402
from test_host_api import StandInCtx  # noqa: E402

DOCS = {"d1": "a b a b c", "d2": "b c a b", "d3": "c"}
RESULT = {}


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.init = lambda args: None
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return name


def _taskfn(emit):
    for k, v in DOCS.items():
        emit(k, v)


def _mapfn(key, value, emit):
    w = value.split()
    for x, y in zip(w, w[1:]):
        emit(T.tuple_(x, y), 1)        # composite key
    emit(T.tuple_(len(w)), 1)          # tuple(x) of one scalar IS the scalar (tuple.lua:255-257)
    emit(T.tuple_([len(w), "len"]), 1)   # a table argument becomes a tuple
    emit("words", len(w))              # scalar key next to tuples


def _reducefn(key, values, emit):
    emit(sum(values))


def _finalfn(it):
    for k, v in it:
        RESULT[k] = v[0]
    return True


def test_bigram_count_with_tuple_keys_over_the_task_api():
    RESULT.clear()
    mods = dict(
        taskfn=_module("tk_taskfn", taskfn=_taskfn), mapfn=_module("tk_mapfn", mapfn=_mapfn),
        partitionfn=_module("tk_partitionfn", partitionfn=lambda k: 0, NUM_REDUCERS=15, hbm_partitionfn="wordhash"),
        reducefn=_module("tk_reducefn", reducefn=_reducefn, hbm_reducefn="sum", associative_reducer=True,
                         commutative_reducer=True, idempotent_reducer=True),
        finalfn=_module("tk_finalfn", finalfn=_finalfn))
    s = server.new("hbm://local", "tuple-keys")
    s.ctx_factory = StandInCtx
    s.configure(dict(mods, storage="hbm", hbm=dict(key_kind="tuple")))
    s.loop()
    want = {}
    for v in DOCS.values():
        w = v.split()
        for x, y in zip(w, w[1:]):
            want[(x.encode(), y.encode())] = want.get((x.encode(), y.encode()), 0) + 1
        want[len(w)] = want.get(len(w), 0) + 1
        want[(len(w), b"len")] = want.get((len(w), b"len"), 0) + 1
        want[b"words"] = want.get(b"words", 0) + len(w)
    assert RESULT == want


@pytest.mark.gpu
def test_bigram_count_with_tuple_keys_on_the_device():
    """same task over the CUDA path: encoded keys are opaque byte strings (128-byte record class); the
    device's ascending key order is the codec's order, so finalfn sees shorter tuples first"""
    RESULT.clear()
    seen = []

    def finalfn(it):
        for k, v in it:
            seen.append(k)
            RESULT[k] = v[0]
        return True

    mods = dict(
        taskfn=_module("tkg_taskfn", taskfn=_taskfn), mapfn=_module("tkg_mapfn", mapfn=_mapfn),
        partitionfn=_module("tkg_partitionfn", partitionfn=lambda k: 0, NUM_REDUCERS=1, hbm_partitionfn="wordhash"),
        reducefn=_module("tkg_reducefn", reducefn=_reducefn, hbm_reducefn="sum", associative_reducer=True,
                         commutative_reducer=True, idempotent_reducer=True),
        finalfn=_module("tkg_finalfn", finalfn=finalfn))
    s = server.new("hbm://local", "gpu-tuple-keys")
    s.configure(dict(mods, storage="hbm", hbm=dict(key_kind="tuple")))
    s.loop()
    want = {}
    for v in DOCS.values():
        w = v.split()
        for x, y in zip(w, w[1:]):
            want[(x.encode(), y.encode())] = want.get((x.encode(), y.encode()), 0) + 1
        want[len(w)] = want.get(len(w), 0) + 1
        want[(len(w), b"len")] = want.get((len(w), b"len"), 0) + 1
        want[b"words"] = want.get(b"words", 0) + len(w)
    assert RESULT == want
    assert [T.encode(k) for k in seen] == sorted(T.encode(k) for k in want)  # one partition: ascending codec order
