"""examples/WordCount/mapfn.lua"""
from .init import mapfn  # noqa: F401


def init(arg=None):
    pass
