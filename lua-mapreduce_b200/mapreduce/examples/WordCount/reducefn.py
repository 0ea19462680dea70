"""examples/WordCount/reducefn.lua: sum with the associative/commutative/idempotent flags"""
from .init import reducefn  # noqa: F401

combinerfn = reducefn
hbm_reducefn = "sum"
associative_reducer = True
commutative_reducer = True
idempotent_reducer = True


def init(arg=None):
    pass
