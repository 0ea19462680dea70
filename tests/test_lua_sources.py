"""The Lua host layer (lua-mapreduce_b200/lua/) cannot run in this image (no Lua).  What can be checked
without an interpreter: a scope lint over the source (a name assigned before its `local` declaration is a
GLOBAL write -- the round-1 bug that left hbm.lua's tuple-key switch dead), a ban on library calls Lua 5.2
does not have, and the double <-> words codec of hbm.lua mirrored line by line and checked against struct."""
import math
import os
import re
import struct

from hypothesis import given, settings, strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LUA_DIR = os.path.join(ROOT, "lua-mapreduce_b200", "lua")
HBM = os.path.join(LUA_DIR, "mapreduce", "hbm.lua")
LUA_KEYWORDS = {"and", "break", "do", "else", "elseif", "end", "false", "for", "function", "goto", "if", "in", "local",
                "nil", "not", "or", "repeat", "return", "then", "true", "until", "while"}


def strip_comments_and_strings(src):
    src = re.sub(r"--\[\[.*?\]\]", "", src, flags=re.S)
    src = re.sub(r"--[^\n]*", "", src)
    src = re.sub(r'"(?:\\.|[^"\\])*"', '""', src)
    src = re.sub(r"'(?:\\.|[^'\\])*'", "''", src)
    return src


def global_writes(src):
    """names assigned (plain `name = ...` statements, incl. multiple assignment) that no `local`, function
    parameter or for-variable declared EARLIER in the file -- i.e. writes to globals"""
    code = strip_comments_and_strings(src)
    declared, bad, depth = set(), [], 0
    for line_no, line in enumerate(code.split("\n"), 1):
        inside_table, depth = depth > 0, depth + line.count("{") - line.count("}")
        if inside_table:  # `field = value` lines of a table constructor are not assignments
            continue
        m = re.match(r"\s*local\s+function\s+([A-Za-z_]\w*)", line)
        if m:
            declared.add(m.group(1))
        m = re.match(r"\s*local\s+([A-Za-z_][\w\s,]*?)(=|$)", line)
        if m and not line.strip().startswith("local function"):
            declared.update(n.strip() for n in m.group(1).split(",") if n.strip())
        for m in re.finditer(r"function\s*[\w.:]*\s*\(([^)]*)\)", line):
            declared.update(n.strip() for n in m.group(1).split(",") if n.strip() and n.strip() != "...")
        m = re.match(r"\s*for\s+([\w\s,]+?)\s*(=|in)\s", line)
        if m:
            declared.update(n.strip() for n in m.group(1).split(","))
        m = re.match(r"\s*([A-Za-z_]\w*(?:\s*,\s*[A-Za-z_]\w*)*)\s*=(?!=)", line)
        if m and not line.strip().startswith("local"):
            for n in (x.strip() for x in m.group(1).split(",")):
                if n not in declared and n not in LUA_KEYWORDS:
                    bad.append((line_no, n))
    return bad


def test_lint_catches_the_round1_bug():
    src = "local ctx\nfunction f()\n  flag = true\nend\nlocal flag = false\nfunction g() return flag end\n"
    assert global_writes(src) == [(3, "flag")]
    assert global_writes("local flag = false\nfunction f()\n  flag = true\nend\n") == []


def test_lua_sources_write_no_globals_and_use_only_lua52_library():
    for dirpath, _, names in os.walk(LUA_DIR):
        for n in names:
            if not n.endswith(".lua"):
                continue
            src = open(os.path.join(dirpath, n)).read()
            assert global_writes(src) == [], (n, global_writes(src))
            code = strip_comments_and_strings(src)
            for banned in ("string.pack", "string.unpack", "math.tointeger", "utf8.", "table.move", "//"):
                assert banned not in code, "%s uses %s (not in Lua 5.2)" % (n, banned)


def test_hbm_lua_shares_one_tuple_switch():
    src = open(HBM).read()
    decl = src.index("local tuple_keys")
    assert decl < src.index("function hbm.configure") < src.index("function hbm.map_job") < src.index("function hbm.groups")
    assert src.count("local tuple_keys") == 1


# ---- hbm.lua's double_to_words / words_to_double, statement for statement
def double_to_words(x):
    if x == 0:
        return 0, 0
    sign = 0
    if x < 0:
        sign, x = 0x80000000, -x
    if x == math.inf:
        return sign + 0x7FF00000, 0
    m, e = math.frexp(x)
    e = e + 1022
    if e <= 0:
        mant, e = m * 2.0 ** (52 + e), 0
    else:
        mant = (m * 2 - 1) * 2.0 ** 52
    hi_m = math.floor(mant / 2.0 ** 32)
    return sign + e * 2 ** 20 + hi_m, mant - hi_m * 2.0 ** 32


def words_to_double(hi, lo):
    neg = hi >= 0x80000000
    if neg:
        hi -= 0x80000000
    e = math.floor(hi / 2 ** 20)
    mant = (hi - e * 2 ** 20) * 2.0 ** 32 + lo
    if e == 0:
        x = math.ldexp(mant, -1074)
    elif e == 2047:
        x = math.inf
    else:
        x = math.ldexp(mant + 2.0 ** 52, e - 1075)
    return -x if neg else x


def test_mirror_matches_the_lua_text():
    """the mirror above is only worth something while it follows the Lua: pin the load-bearing lines"""
    src = open(HBM).read()
    for frag in ("local m, e = math.frexp(x)", "e = e + 1022", "mant, e = m * 2 ^ (52 + e), 0", "mant = (m * 2 - 1) * 2 ^ 52",
                 "return sign + e * 2 ^ 20 + hi_m, mant - hi_m * 2 ^ 32", "x = math.ldexp(mant, -1074)",
                 "x = math.ldexp(mant + 2 ^ 52, e - 1075)"):
        assert frag in src, frag


@settings(max_examples=400, deadline=None)
@given(st.floats(allow_nan=False))
def test_double_codec_is_ieee754(x):
    if x == 0:
        x = 0.0  # hbm.lua folds -0 into 0 before encoding
    hi, lo = double_to_words(x)
    assert (int(hi), int(lo)) == struct.unpack(">II", struct.pack(">d", x))
    assert words_to_double(int(hi), int(lo)) == x


def test_double_codec_edge_values():
    for x in (5e-324, 2.2250738585072014e-308, 2.225073858507201e-308, 1.0, -1.0, 1.7976931348623157e308, math.inf, -math.inf,
              2.0 ** 53, 0.1, -123456.789e-200):
        hi, lo = double_to_words(x)
        assert (int(hi), int(lo)) == struct.unpack(">II", struct.pack(">d", x)), x
        assert words_to_double(int(hi), int(lo)) == x


def test_lua_c_module_type_checks_against_lua52_prototypes():
    """mrhbm_lua.c cannot be built here (no lua.h).  With DECLARATIONS of the Lua 5.2 C API subset it uses
    (tests/native/lua_stub, written from the reference manual) gcc -fsyntax-only checks every call against
    include/mrhbm.h and the Lua prototypes: argument counts, types, undeclared names."""
    import subprocess
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-std=gnu99",
                        "-I" + os.path.join(ROOT, "tests", "native", "lua_stub"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(LUA_DIR, "mrhbm_lua.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
