"""The reference's WordCount example (mapreduce/examples/WordCount/init.lua) as ONE module
providing all six functions -- test.sh:54-71 "INIT SCRIPT" configuration.  The per-role modules
next to this file (taskfn.py, mapfn.py, ...) mirror the reference's per-file plugins."""
import re

NUM_REDUCERS = 15
hbm_partitionfn = "fnv_lua"   # == partitionfn below, evaluated on the device
hbm_reducefn = "sum"          # == reducefn below
_TOKEN = re.compile(rb"[^ \t\n\v\f\r]+")  # Lua "[^%s]+" in the C locale
FILES = []


def init(arg):
    pass


def taskfn(emit):
    for i, f in enumerate(FILES):
        emit(i + 1, f)


def mapfn(key, value, emit):
    with open(value, "rb") as fh:
        for line in fh:
            for w in _TOKEN.findall(line):
                emit(w, 1)


def partitionfn(key):
    """examples/WordCount/partitionfn.lua:8-16 (host copy, for documentation and tests)"""
    import math
    h = 2166136261.0
    for b in key:
        h = math.fmod(h * 16777619.0, 4294967296.0)
        h = float(int(h) ^ b)
    return int(h) % NUM_REDUCERS


def reducefn(key, values, emit):
    count = 0
    for v in values:
        count += v
    emit(count)


combinerfn = reducefn
RESULT = {}


def finalfn(pairs_iterator):
    for key, values in pairs_iterator:
        RESULT[key] = values[0]
    return True


associative_reducer = True
commutative_reducer = True
idempotent_reducer = True
