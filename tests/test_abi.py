"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every
symbol include/mrhbm.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import mrhbm_loader

mrhbm_loader.load()
from lua_mapreduce_b200 import mrhbm  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mrhbm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrhbm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = mrhbm.load()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libmrhbm.so does not export %s" % n
    assert sorted(mrhbm.EXPORTS) == names
    assert L.mrhbm_abi_version() == 2


def test_struct_layouts_match_header():
    assert C.sizeof(mrhbm.Config) == 48
    assert C.sizeof(mrhbm.ResultInfo) == 40
    assert C.sizeof(mrhbm.Stats) == 88


def test_init_rejects_bad_struct_size():
    L = mrhbm.load()
    cfg = mrhbm.Config()
    cfg.struct_size = 4
    h = C.c_void_p()
    assert L.mrhbm_init(C.byref(cfg), C.byref(h)) == -1 and not h


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="GPU present")
def test_no_cpu_fallback():
    with pytest.raises(mrhbm.MrhbmError) as e:
        mrhbm.Ctx()
    assert e.value.code == mrhbm.E_NODEVICE and "no CPU path" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lua-mapreduce_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".c", ".lua")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "mr_oracle" not in txt and "import oracle" not in txt, os.path.join(d, f)


def test_zipf_table_shape():
    from lua_mapreduce_b200 import synth
    t = synth.zipf_table(1 << 12)
    assert t.dtype.name == "uint64" and t.size == 1 << 12 and int(t[-1]) == 2**64 - 1
    assert (t[1:] >= t[:-1]).all()
    # top-1 mass of Zipf(1.1) over a small vocabulary
    assert 0.15 < int(t[0]) / 2**64 < 0.25


def test_fnv_in_doubles_integer_emulation_matches_real_doubles(tmp_path):
    """csrc/mrhbm_dev.cuh fnv_lua_step (32-bit integer emulation of the example partitionfn's
    double arithmetic, examples/WordCount/partitionfn.lua:8-16) against real IEEE doubles on
    3e7 random (h, byte) pairs + edge values: host-compiled from the same header."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "fnvcheck")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "lua-mapreduce_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "native", "fnv_emulation_check.cu")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout
