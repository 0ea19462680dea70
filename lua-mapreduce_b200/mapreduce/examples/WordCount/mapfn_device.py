"""examples/WordCount/mapfn.lua with the tokeniser declared as the device built-in."""
from .init import mapfn  # noqa: F401  (host definition, used when the storage has no device tokeniser)

hbm_mapfn = "wordcount_file"


def init(arg=None):
    pass
