#!/usr/bin/env python3
"""Generates tests/golden/* from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box).

Independent of oracle/: everything here is derived with plain Python from
  - the reference's test.sh corpus (test.sh:12-15; examples/WordCount/taskfn.lua:8-11),
    tokenised like examples/WordCount/mapfn.lua:3-9 / misc/naive.lua:2-5 ("[^%s]+"), and
  - the example partitioner evaluated in Python floats == IEEE doubles
    (examples/WordCount/partitionfn.lua:8-16),
so the fixtures pin the C oracle instead of echoing it.  Only derived data (word -> count ->
partition) is written; no reference source text is copied.
"""
import hashlib, json, math, os, re, sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
CORPUS = ["mapreduce/server.lua", "mapreduce/worker.lua", "mapreduce/test.lua", "mapreduce/utils.lua"]
TOKEN = re.compile(rb"[^ \t\n\v\f\r]+")  # Lua %s in the C locale


def fnv_lua(key: bytes) -> int:
    h = 2166136261.0
    for b in key:
        h = math.fmod(h * 16777619.0, 4294967296.0)  # exact: power-of-two modulus
        h = float(int(h) ^ b)
    return int(h)


def main():
    counts, per_file = {}, []
    for i, f in enumerate(CORPUS):
        toks = TOKEN.findall(open(os.path.join(REF, f), "rb").read())
        per_file.append(len(toks))
        for t in toks:
            c = counts.setdefault(t, [0, 0, 0, 0])
            c[i] += 1
    rows = sorted(counts.items())
    with open(os.path.join(HERE, "wordcount_testsh.tsv"), "w") as out:
        out.write("# key_hex\tpartition(FNV-lua mod 15)\tcount_in_job1\tjob2\tjob3\tjob4\n")
        for k, c in rows:
            out.write("%s\t%d\t%d\t%d\t%d\t%d\n" % (k.hex(), fnv_lua(k) % 15, *c))
    lines = sorted(b"%d %s\n" % (sum(c), k) for k, c in rows)
    meta = {
        "corpus": CORPUS,
        "tokens_per_file": per_file,
        "tokens": sum(per_file),
        "distinct": len(rows),
        "max_key_len": max(len(k) for k in counts),
        "sha256_sorted_count_word_lines": hashlib.sha256(b"".join(lines)).hexdigest(),
        "distinct_per_partition": [sum(1 for k in counts if fnv_lua(k) % 15 == p) for p in range(15)],
        "partitionfn_known_answers": {k.decode(): [fnv_lua(k) % 15, fnv_lua(k)]
                                      for k in (b"a", b"the", b"local", b"function", b"mapreduce")},
        # mapreduce/utils.lua:345-349
        "escape": [[120, "120"], ["30", "\"30\""], ["30\n", "\"30\\n\""]],
        "serialize_table_ipairs": [[[1, 2, 3, "hola"], "{1,2,3,\"hola\"}"]],
        "keys_sorted": [[["c", "a", "b"], ["a", "b", "c"]]],
        # mapreduce/utils.lua:360-380
        "merge": {"files": {"f1": ["return 1,{1,1}", "return 2,{1}", "return 3,{1}"],
                            "f2": ["return 1,{1,1,1,1}", "return 3,{1}", "return 4,{1}"]},
                  "result": [[1, [1, 1, 1, 1, 1, 1]], [2, [1]], [3, [1, 1]], [4, [1]]]},
        # mapreduce/heap.lua:99-118
        "heap": {"push": [20, 10, 15, 1], "pop": [1, 10, 15, 20]},
        # mapreduce/server.lua:630-636
        "count_digits": [[0, 1], [1, 1], [9, 1], [10, 2], [99, 2], [111, 3], [1111, 4]],
        # SURVEY App. B self-tests (spec of the synthetic streams, not reference data)
        "splitmix64": {"seed": 0x5EED20260921,
                       "out": ["0xe8c7281a3e72a16c", "0xcb1236c3aceab168", "0x0befa0aa81cb0cac"]},
        "rank_to_key": {"1": "aFJFE", "2": "bV", "3": "cLPDLKZM", "27": "aaWYJCDM",
                        "1000": "allJRH", "1048576": "bgqcvYKV"},
    }
    json.dump(meta, open(os.path.join(HERE, "vectors.json"), "w"), indent=1)
    print(json.dumps({k: meta[k] for k in ("tokens", "distinct", "max_key_len",
                                           "sha256_sorted_count_word_lines")}))


if __name__ == "__main__":
    main()
