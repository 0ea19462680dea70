"""lua-mapreduce_b200 -- B200-native shuffle/sort/reduce core behind lua-mapreduce's task API.

The directory name carries a hyphen (it mirrors the reference's name), so import it through
the root-level loader:  ``import mrhbm_loader; pkg = mrhbm_loader.load()`` which registers
this package as ``lua_mapreduce_b200``.

  csrc/        hand-written sm_100a kernels + the C-ABI runtime (include/mrhbm.h)
  build.py     nvcc build of lib/libmrhbm.so
  mrhbm.py     ctypes binding of the C ABI (what the Lua C module binds)
  mapreduce/   host-side mirror of the reference's server/worker/job API with storage="hbm"
  lua/         the Lua 5.2 host layer + C module source (not executable in this image)
"""
__version__ = "0.1.0"
