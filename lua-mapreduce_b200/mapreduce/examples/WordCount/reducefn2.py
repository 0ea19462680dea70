"""examples/WordCount/reducefn2.lua: the same sum WITHOUT the algebra flags (general reducer path)"""
from .init import reducefn  # noqa: F401

combinerfn = reducefn
hbm_reducefn = "sum"


def init(arg=None):
    pass
