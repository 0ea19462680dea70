import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_vectors():
    import json
    return json.load(open(os.path.join(GOLDEN, "vectors.json")))


@pytest.fixture(scope="session")
def golden_wordcount():
    """[(key bytes, partition, [count per map job])] from tests/golden/wordcount_testsh.tsv"""
    rows = []
    for line in open(os.path.join(GOLDEN, "wordcount_testsh.tsv")):
        if line.startswith("#"):
            continue
        f = line.rstrip("\n").split("\t")
        rows.append((bytes.fromhex(f[0]), int(f[1]), [int(x) for x in f[2:6]]))
    return rows


def expand_tokens(rows, job, seed=7):
    """Deterministic token stream of one map job rebuilt from the golden table
    (word counts are emission-order independent)."""
    import random
    toks = []
    for k, _, c in rows:
        toks.extend([k] * c[job])
    random.Random(seed + job).shuffle(toks)
    return toks
