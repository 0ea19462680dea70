"""examples/WordCount/taskfn.lua"""
from .init import FILES


def init(arg):
    pass


def taskfn(emit):
    for i, f in enumerate(FILES):
        emit(i + 1, f)
