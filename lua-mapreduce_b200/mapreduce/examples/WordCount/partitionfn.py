"""examples/WordCount/partitionfn.lua"""
from .init import partitionfn, NUM_REDUCERS  # noqa: F401

hbm_partitionfn = "fnv_lua"


def init(arg=None):
    pass
