"""Host-side mirror of lua-mapreduce's task API (mapreduce/init.lua:26-40) with the Mongo-backed
shuffle replaced by the in-HBM path behind include/mrhbm.h (storage = "hbm").

  server.new(cnn, dbname[, auth]) / :configure{...} / :loop()     mapreduce/server.lua:419-624
  worker.new(cnn, dbname[, auth]) / :configure{...} / :execute()  mapreduce/worker.lua:112-167
  plugin modules: init + taskfn / mapfn / partitionfn / reducefn / combinerfn / finalfn

Written in Python because the image has no Lua toolchain; lua/ holds the Lua 5.2 layer a
maintainer would ship (same call sequence into the same C ABI).
"""
from . import utils, server, worker  # noqa: F401

_VERSION = "0.4.0-hbm"
_NAME = "mapreduce"
