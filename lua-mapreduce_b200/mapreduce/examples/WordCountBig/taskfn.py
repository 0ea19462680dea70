"""examples/WordCountBig/taskfn.lua:6-12: one map job per split file, key = 1-based position in the listing,
value = the file's path.  The reference hard-codes its corpus directory in an `ls` call; here the directory comes
from init's argument (server.configure{init_args = {dir = ...}}), listed in the same sorted order `ls` prints."""
import os

DIR = None


def init(arg):
    global DIR
    if arg and arg.get("dir"):
        DIR = arg["dir"]


def taskfn(emit):
    assert DIR, "WordCountBig.taskfn needs init_args = {'dir': <directory of text splits>}"
    for i, name in enumerate(sorted(os.listdir(DIR))):
        emit(i + 1, os.path.join(DIR, name))
