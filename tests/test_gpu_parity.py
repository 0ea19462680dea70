"""GPU parity: the CUDA shuffle/sort/reduce path (through the C ABI) against the CPU oracle on
the same seeded inputs, against the golden word count, and -- at larger sizes -- through
size-independent properties (linearity of the sum, sortedness, partition membership)."""
import hashlib

import numpy as np
import pytest

import mrhbm_loader
import oracle as O
from conftest import expand_tokens

mrhbm_loader.load()
from lua_mapreduce_b200 import mrhbm, synth  # noqa: E402

pytestmark = pytest.mark.gpu
SEED = synth.SEED


def u64_records(keys, vals):
    recs = np.zeros(keys.size, dtype=mrhbm.record_dtype(mrhbm.KEY_U64))
    recs["key"], recs["val"] = keys, vals
    return recs


def check_vs_oracle_u64(ctx, keys, vals, P, partitioner=O.PART_MULHASH):
    ok, osum, po = O.groupby_u64(keys, vals, partitioner, P)
    info = ctx.result_info()
    assert info.groups == ok.size and info.pairs_in == keys.size
    gk, gs, gpo = ctx.result_copy()
    assert (gpo == po).all()
    if info.sorted:
        assert (gk == ok).all() and (gs == osum).all()
    else:  # partition = several ascending runs: compare as sorted sets per partition
        for p in range(P):
            a, b = int(po[p]), int(po[p + 1])
            order = np.argsort(gk[a:b], kind="stable")
            assert (gk[a:b][order] == ok[a:b]).all() and (gs[a:b][order] == osum[a:b]).all()
    # the iterator always yields ascending keys (utils.merge_iterator order), 8-byte BE keys
    nonempty = [p for p in range(P) if po[p + 1] > po[p]]
    assert ctx.partitions() == nonempty
    for p in nonempty[:3] + nonempty[-2:]:
        got = list(ctx.groups(p))
        a, b = int(po[p]), int(po[p + 1])
        assert [int.from_bytes(k, "big") for k, _ in got] == ok[a:b].tolist()
        assert [v[0] for _, v in got] == osum[a:b].tolist()
    cin, cout = ctx.checksum_input(), ctx.checksum_result()
    assert cin[:3] == cout[:3] and cin[3] == keys.size and cout[3] == ok.size and cout[4:] == [0, 0]


@pytest.mark.parametrize("flags", [0, mrhbm.F_NO_OPTIMISTIC])  # single-pass and two-pass partition layouts
@pytest.mark.parametrize("n,P", [(1, 1), (37, 4), (5000, 16), (200_000, 16), (1_000_000, 1024), (6_000_000, 1024)])
def test_u64_uniform_vs_oracle(n, P, flags):
    keys, vals = O.gen_u64(SEED, 0, n)
    with mrhbm.Ctx(mrhbm.KEY_U64, P, flags=flags) as ctx:
        m = ctx.map_begin("m1")
        m.emit_batch(u64_records(keys, vals))
        m.commit()
        ctx.shuffle()
        assert ctx.stats()["attempts"] == 1
        check_vs_oracle_u64(ctx, keys, vals, P)


def test_device_generator_matches_oracle_stream():
    n, P = 300_000, 64
    keys, vals = O.gen_u64(SEED, 12345, n)
    with mrhbm.Ctx(mrhbm.KEY_U64, P) as ctx:
        for j in range(3):  # three map jobs over disjoint counter ranges
            m = ctx.map_begin(j)
            m.gen_u64(SEED, 12345 + j * 100_000, 100_000)
            m.commit()
        ctx.shuffle()
        check_vs_oracle_u64(ctx, keys, vals, P)


def test_u64_duplicates_and_hot_key():
    """few distinct keys + one hot key: key-ordered bins overflow -> hash sub-bins -> big-bin path"""
    rng = np.random.default_rng(3)
    n, P = 400_000, 8
    keys = O.gen_u64(SEED, 0, 2000)[0][rng.integers(0, 2000, n)]
    keys[rng.random(n) < 0.3] = np.uint64(0xDEADBEEFCAFEF00D)
    vals = rng.integers(0, 1 << 20, n).astype(np.uint32)
    with mrhbm.Ctx(mrhbm.KEY_U64, P) as ctx:
        m = ctx.map_begin("dups")
        m.emit_batch(u64_records(keys, vals))
        m.commit()
        ctx.shuffle()
        st = ctx.stats()
        assert st["big_bins"] >= 1
        check_vs_oracle_u64(ctx, keys, vals, P)


def test_u64_sparse_duplicates_inside_key_ordered_bins():
    """mostly unique keys plus a few keys repeated 9..40 times: the bins stay balanced (one pass, key-ordered
    sub-bins, pipelined sort kernel) but some buckets exceed the rank-among-mates limit, so those bins take
    the general sort inside the same kernel"""
    rng = np.random.default_rng(11)
    n0, P = 500_000, 16
    keys, vals = O.gen_u64(SEED, 4242, n0)
    reps = rng.integers(9, 41, 300)
    dup = np.repeat(keys[rng.integers(0, n0, 300)], reps)
    keys = np.concatenate([keys, dup])
    vals = np.concatenate([vals, rng.integers(0, 1 << 16, dup.size).astype(np.uint32)])
    perm = rng.permutation(keys.size)
    keys, vals = keys[perm], vals[perm]
    for flags in (0, mrhbm.F_NO_OPTIMISTIC):
        with mrhbm.Ctx(mrhbm.KEY_U64, P, flags=flags) as ctx:
            m = ctx.map_begin("sparse-dups")
            m.emit_batch(u64_records(keys, vals))
            m.commit()
            ctx.shuffle()
            st = ctx.stats()
            assert st["attempts"] == 1 and st["sub_bins"] > 1 and ctx.result_info().sorted == 1
            check_vs_oracle_u64(ctx, keys, vals, P)


def test_u64_neighbouring_keys_and_small_repeats_in_key_ordered_bins():
    """The u64 sort kernel ranks bucket mates on the top 32 significant bits of key - bin start and compares whole
    keys only on a tie.  Keys that differ in their LOWEST bits only (k, k+1, k+2, k+5: same bin, same bucket, same
    32-bit sort key) and keys repeated 2..4 times (real duplicates: the head sums its equals in place) must come
    out in full 64-bit order with exact sums; the bins stay balanced, so everything stays on the one-pass path."""
    rng = np.random.default_rng(29)
    n0, P = 400_000, 4
    base, vals = O.gen_u64(SEED, 777, n0)
    base = base & np.uint64(0xFFFFFFFFFFFFFFF0)  # room for the small offsets
    near = np.concatenate([base[:60_000] + np.uint64(d) for d in (1, 2, 5)])
    reps = np.repeat(base[60_000:90_000], rng.integers(1, 4, 30_000))  # 2..4 copies with the original
    keys = np.concatenate([base, near, reps])
    vals = np.concatenate([vals, rng.integers(0, 1 << 31, near.size + reps.size).astype(np.uint32)])
    perm = rng.permutation(keys.size)
    keys, vals = keys[perm], vals[perm]
    for flags in (0, mrhbm.F_NO_OPTIMISTIC):
        with mrhbm.Ctx(mrhbm.KEY_U64, P, flags=flags) as ctx:
            m = ctx.map_begin("near")
            m.emit_batch(u64_records(keys, vals))
            m.commit()
            ctx.shuffle()
            st = ctx.stats()
            assert st["attempts"] == 1 and st["sub_bins"] > 1 and ctx.result_info().sorted == 1
            check_vs_oracle_u64(ctx, keys, vals, P)
    with mrhbm.Ctx(mrhbm.KEY_U64, P, reducer=mrhbm.RED_NONE) as ctx:  # group-only: equal keys keep all their rows
        m = ctx.map_begin("near")
        m.emit_batch(u64_records(keys, vals))
        m.commit()
        ctx.shuffle()
        want = {}
        for k, v in zip(keys.tolist(), vals.tolist()):
            want.setdefault(k, []).append(v)
        got = {int.from_bytes(k, "big"): sorted(v) for p in ctx.partitions() for k, v in ctx.groups(p)}
        assert got == {k: sorted(v) for k, v in want.items()}


@pytest.mark.parametrize("combiner", [False, True])
def test_u64_values_wider_than_32_bits(combiner):
    """u64-key records carry 64-bit values (ABI 2): any integer-valued Lua number < 2^53 (job.lua:83-97 emits
    arbitrary values; reducefn.lua:1-5 adds doubles).  Sums here pass 2^32 many times over."""
    n, P = 1_200_000 if combiner else 300_000, 16
    rng = np.random.default_rng(21)
    keys = O.gen_u64(SEED, 99, 4000)[0][rng.integers(0, 4000, n)]
    vals = rng.integers(0, 1 << 40, n).astype(np.uint64)
    vals[::7] = rng.integers(0, 60000, vals[::7].size).astype(np.uint64)  # small ones go through the shared-memory tables
    with mrhbm.Ctx(mrhbm.KEY_U64, P, combiner=combiner) as ctx:
        m = ctx.map_begin("wide")
        m.emit_batch(u64_records(keys[: n - 3], vals[: n - 3]))
        for k, v in zip(keys[n - 3:].tolist(), vals[n - 3:].tolist()):
            m.emit(k, v)
        with pytest.raises(mrhbm.MrhbmError):
            m.emit(5, 1 << 53)
        with pytest.raises(mrhbm.MrhbmError):
            m.emit(5, -1)
        m.commit()
        ctx.shuffle()
        check_vs_oracle_u64(ctx, keys, vals, P)
        assert int(ctx.result_copy()[1].max()) > 1 << 40


def test_keys_with_nul_and_control_bytes():
    """any byte string is a key (utils.lua:104-110 escapes them for the spill; SURVEY A.3: "a" < "a\\0"): bytes 0x00 and
    0x01 cross the boundary escaped, order and partition (FNV over the ORIGINAL bytes) as the reference has them"""
    rng = np.random.default_rng(12)
    alphabet = [b"\0", b"\1", b"\2", b"a", b"b", b"\xff", b" "]
    keys = [b"", b"a", b"a\0", b"a\0\0", b"a\1", b"a\1\1", b"a\2", b"ab", b"\0", b"\1", b"\0a", b"\1\0\1"]
    keys += [b"".join(alphabet[j] for j in rng.integers(0, len(alphabet), rng.integers(1, 12))) for _ in range(3000)]
    pairs = [(keys[j], int(v)) for j, v in zip(rng.integers(0, len(keys), 40_000), rng.integers(0, 1000, 40_000))]
    pairs += [(b"\0" * 14, 3), (b"\0" * 13, 4), (b"\0" * 14, 5)]  # 14 NULs need 28 slot bytes: one too many for the record
    eng = O.Engine(O.PART_FNV_LUA, 15, combiner=-1, reducer=O.RED_SUM, aci=True)
    for job in range(4):
        eng.map_job(job + 1, pairs=pairs[job::4])
    eng.reduce_all()
    want = [(p, k, [int(x) for x in v]) for p, k, v in eng.final_pairs()]
    with mrhbm.Ctx(mrhbm.KEY_STR, 15, mrhbm.PART_FNV_LUA, max_key_bytes=27) as ctx:
        for job in range(4):
            m = ctx.map_begin(job + 1)
            for k, v in pairs[job::4]:
                m.emit(k, v)
            m.commit()
        ctx.shuffle()
        got = [(p, k, v) for p in ctx.partitions() for k, v in ctx.groups(p)]
    assert got == want


def test_keys_longer_than_a_record_slot():
    """The reference takes keys of any length (utils.lua:104-110).  Pairs whose key does not fit the ctx record class
    stay on the host, are partitioned with the same FNV-in-doubles partitioner, grouped at the barrier and merged
    into the reduce-side iteration at their place in the bytewise key order -- also next to device keys that are
    their proper prefixes, with NUL bytes inside, across map jobs, after a re-commit and an abort."""
    rng = np.random.default_rng(31)
    P = 15
    short = [O.rank_to_key(int(r)) for r in rng.integers(1, 3000, 20_000)]
    stem = b"abcdefghijklmnopqrstuvwxyz0"  # 27 bytes: the longest key a 32-byte record holds
    longs = [stem + b"X", stem + b"X" * 40, stem + b"\x00tail", stem + b"Y" * 200, b"q" * 5000, b"\x01" * 20, b"z" * 28,
             bytes(range(1, 200)), stem[:20] + b"\x00" * 8]
    jobs = []
    for j in range(3):
        pairs = [(w, int(v)) for w, v in zip(short[j::3], rng.integers(1, 1000, len(short[j::3])))]
        pairs += [(stem, 7 + j), (stem[:26], 1)]  # device keys that are proper prefixes of long ones
        pairs += [(k, int(v)) for k, v in zip(longs, rng.integers(1, 1 << 31, len(longs)))] * (j + 1)
        rng.shuffle(pairs)
        jobs.append(pairs)
    e = O.Engine(O.PART_FNV_LUA, P, combiner=-1, reducer=O.RED_SUM, aci=True)
    for j, pairs in enumerate(jobs):
        e.map_job(j, pairs=pairs)
    e.reduce_all()
    want = [(p, k, [int(x) for x in v]) for p, k, v in e.final_pairs()]
    for combiner in (False, True):
        with mrhbm.Ctx(mrhbm.KEY_STR, P, mrhbm.PART_FNV_LUA, max_key_bytes=27, combiner=combiner) as ctx:
            m = ctx.map_begin(1)  # a first version of job 1 that the re-commit below replaces, long keys included
            m.emit(b"gone" * 20, 5)
            m.emit(b"gone", 5)
            m.commit()
            for j, pairs in enumerate(jobs):
                m = ctx.map_begin(j)
                for k, v in pairs:
                    m.emit(k, v)
                m.commit()
            m = ctx.map_begin("broken")
            m.emit(b"never" * 30, 1)
            m.abort()
            ctx.shuffle()
            got = [(p, k, v) for p in ctx.partitions() for k, v in ctx.groups(p)]
            assert got == want
            with pytest.raises(mrhbm.MrhbmError) as ei:  # fixed-width rows cannot hold them
                ctx.result_copy()
            assert ei.value.code == -4
            ctx.reset()  # the next task iteration starts empty on the host side too
            m = ctx.map_begin(0)
            m.emit(b"only", 1)
            m.commit()
            ctx.shuffle()
            assert [(k, v) for p in ctx.partitions() for k, v in ctx.groups(p)] == [(b"only", [1])]
            ctx.reset()  # nothing but a long key: the device side of the shuffle is empty
            m = ctx.map_begin(0)
            m.emit(b"L" * 100, 2)
            m.emit(b"L" * 100, 3)
            m.commit()
            ctx.shuffle()
            assert [(k, v) for p in ctx.partitions() for k, v in ctx.groups(p)] == [(b"L" * 100, [5])]
            # the device tokeniser meets a word that does not fit: that piece of the text is tokenised on the host
            ctx.reset()
            url = b"http://" + b"x" * 300
            text = b"aa bb " + url + b"\ncc aa\t" + url + b" " + b"y" * 28
            m = ctx.map_begin("text")
            assert m.wordcount(text) == 7
            m.commit()
            ctx.shuffle()
            assert sorted((k, v[0]) for p in ctx.partitions() for k, v in ctx.groups(p)) == \
                sorted({b"aa": 2, b"bb": 1, b"cc": 1, url: 2, b"y" * 28: 1}.items())
    # general reducer: every value of a long key; the other partitioner: each key in exactly one partition, ascending
    with mrhbm.Ctx(mrhbm.KEY_STR, 4, mrhbm.PART_WORDHASH, max_key_bytes=27, reducer=mrhbm.RED_NONE) as ctx:
        for j, pairs in enumerate(jobs):
            m = ctx.map_begin(j)
            for k, v in pairs:
                m.emit(k, v)
            m.commit()
        ctx.shuffle()
        allv = {}
        for pairs in jobs:
            for k, v in pairs:
                allv.setdefault(k, []).append(v)
        seen = {}
        for p in ctx.partitions():
            ks = [k for k, _ in ctx.groups(p)]
            assert ks == sorted(ks)
            for k, v in ctx.groups(p):
                assert k not in seen
                seen[k] = sorted(v)
        assert seen == {k: sorted(v) for k, v in allv.items()}


def test_u64_clustered_keys_fall_back_to_runs():
    """sequential integers: top key bits are constant, so key-ordered sub-bins cannot balance"""
    n, P = 300_000, 4
    keys = np.arange(n, dtype=np.uint64) * np.uint64(3) + np.uint64(17)
    vals = np.ones(n, dtype=np.uint32)
    with mrhbm.Ctx(mrhbm.KEY_U64, P) as ctx:
        m = ctx.map_begin("seq")
        m.emit_batch(u64_records(keys, vals))
        m.commit()
        ctx.shuffle()
        # the key sample taken before the shuffle sees the clustering: hash sub-bins from the start, no discarded attempt
        assert ctx.stats()["attempts"] == 1 and ctx.result_info().sorted == 0
        check_vs_oracle_u64(ctx, keys, vals, P)


def test_u64_scalar_emit_path_and_staging_flush():
    n, P = 150_000, 16  # > 2 staging buffers of 65536 records
    keys, vals = O.gen_u64(SEED, 777, n)
    with mrhbm.Ctx(mrhbm.KEY_U64, P, flags=mrhbm.F_FORCE_RUNS) as ctx:
        m = ctx.map_begin("scalar")
        for k, v in zip(keys.tolist(), vals.tolist()):
            m.emit(k, v)
        m.commit()
        ctx.shuffle()
        check_vs_oracle_u64(ctx, keys, vals, P)


def str_records(words, vals, max_key_bytes):
    recs = np.zeros(len(words), dtype=mrhbm.record_dtype(mrhbm.KEY_STR, max_key_bytes))
    recs["key"] = words
    recs["val"] = vals
    return recs


def check_vs_oracle_str(ctx, recs, P, partitioner):
    raw = recs.view(np.uint8).reshape(len(recs), -1)
    okeys, osum, po = O.groupby_rec(raw, partitioner, P)
    info = ctx.result_info()
    assert info.groups == osum.size
    gk, gs, gpo = ctx.result_copy()
    assert (gpo == po).all()
    want = [bytes(k).rstrip(b"\0") for k in okeys]
    for p in range(P):
        a, b = int(po[p]), int(po[p + 1])
        got = list(ctx.groups(p))
        assert [k for k, _ in got] == want[a:b]
        assert [v[0] for _, v in got] == osum[a:b].tolist()
        assert sorted(zip(gk[a:b].tolist(), gs[a:b].tolist())) == list(zip(want[a:b], osum[a:b].tolist()))
    cin, cout = ctx.checksum_input(), ctx.checksum_result()
    assert cin[:3] == cout[:3] and cout[3] == osum.size and cout[4:] == [0, 0]


def test_wordcount_golden_through_c_abi(golden_vectors, golden_wordcount):
    """test.sh:8-53 on the reference's own corpus: 4 map jobs, FNV-in-doubles partitioner mod 15,
    sum reducer; keys up to 84 bytes -> 128-byte records.  Must equal misc/naive.lua."""
    with mrhbm.Ctx(mrhbm.KEY_STR, 15, mrhbm.PART_FNV_LUA, max_key_bytes=golden_vectors["max_key_len"]) as ctx:
        assert ctx.record_bytes == 128
        for job in range(4):
            m = ctx.map_begin(job + 1)
            for t in expand_tokens(golden_wordcount, job):
                m.emit(t, 1)
            m.commit()
        ctx.shuffle()
        assert ctx.partitions() == list(range(15))
        got = {}
        for p in ctx.partitions():
            ks = []
            for k, v in ctx.groups(p):
                assert len(v) == 1
                got[k] = (p, v[0])
                ks.append(k)
            assert ks == sorted(ks)
        assert got == {k: (p, sum(c)) for k, p, c in golden_wordcount}
        per = [sum(1 for k in got if got[k][0] == p) for p in range(15)]
        assert per == golden_vectors["distinct_per_partition"]
        lines = sorted(b"%d %s\n" % (c, k) for k, (_, c) in got.items())
        assert hashlib.sha256(b"".join(lines)).hexdigest() == golden_vectors["sha256_sorted_count_word_lines"]


@pytest.mark.parametrize("partitioner", [mrhbm.PART_FNV_LUA, mrhbm.PART_WORDHASH])
def test_zipf_strings_vs_oracle(partitioner):
    """Zipf(1.1) word stream generated on the device == the oracle's stream; hot keys exercise
    the big-bin path."""
    V, n, P = 1 << 14, 400_000, 15
    table = synth.zipf_table(V)
    recs = O.gen_zipf_rec32(SEED, 0, n, table)
    opart = O.PART_FNV_LUA if partitioner == mrhbm.PART_FNV_LUA else O.PART_FNV64
    with mrhbm.Ctx(mrhbm.KEY_STR, P, partitioner, max_key_bytes=27) as ctx:
        for j in range(2):
            m = ctx.map_begin(j)
            m.gen_zipf(SEED, j * (n // 2), n // 2, table)
            m.commit()
        ctx.shuffle()
        assert ctx.stats()["big_bins"] >= 1
        check_vs_oracle_str(ctx, recs.view(mrhbm.record_dtype(mrhbm.KEY_STR, 27)).reshape(-1), P, opart)


@pytest.mark.parametrize("mkb", [27, 59, 123])
def test_shared_prefix_keys_take_the_full_key_path(mkb):
    """keys that agree on their first 8+ bytes force the multi-pass (whole key) ordering"""
    rng = np.random.default_rng(11)
    n, P = 60_000, 5
    stems = [b"internationalisation", b"internat", b"interna", b"x" * (mkb - 6), b""]
    words = []
    for i in rng.integers(0, 3000, n):
        s = stems[i % len(stems)]
        words.append((s + b"%d" % (i // len(stems)))[:mkb])
    vals = rng.integers(1, 100, n).astype(np.uint32)
    recs = str_records(words, vals, mkb)
    with mrhbm.Ctx(mrhbm.KEY_STR, P, mrhbm.PART_WORDHASH, max_key_bytes=mkb) as ctx:
        m = ctx.map_begin("p")
        m.emit_batch(recs)
        m.commit()
        ctx.shuffle()
        check_vs_oracle_str(ctx, recs, P, O.PART_FNV64)


@pytest.mark.parametrize("mkb", [27, 59, 123])
@pytest.mark.parametrize("flags", [0, mrhbm.F_NO_OPTIMISTIC])
def test_string_keys_many_partitions_use_the_tile_split(mkb, flags):
    """>= 2048 bins: the partition stage is the two-level tile split for every record class (32/64/128 B),
    in the fixed-stride and in the exact layout"""
    rng = np.random.default_rng(mkb)
    n, P, V = 150_000, 2048, 40_000
    vocab = [(b"k%d-" % i + b"abcdefghij" * 12)[: 3 + (i * 7) % (mkb - 2)] for i in range(V)]
    words = [vocab[i] for i in rng.integers(0, V, n)]
    vals = rng.integers(1, 1000, n).astype(np.uint32)
    recs = str_records(words, vals, mkb)
    with mrhbm.Ctx(mrhbm.KEY_STR, P, mrhbm.PART_WORDHASH, max_key_bytes=mkb, flags=flags) as ctx:
        m = ctx.map_begin("many-partitions")
        m.emit_batch(recs)
        m.commit()
        ctx.shuffle()
        assert ctx.stats()["bins"] >= 2048
        check_vs_oracle_str(ctx, recs, P, O.PART_FNV64)


def test_small_bins_flag_exercises_overflow_paths():
    keys, vals = O.gen_u64(SEED, 0, 50_000)
    keys[::7] = keys[0]
    with mrhbm.Ctx(mrhbm.KEY_U64, 4, flags=mrhbm.F_SMALL_BINS) as ctx:
        m = ctx.map_begin("s")
        m.emit_batch(u64_records(keys, vals))
        m.commit()
        ctx.shuffle()
        assert ctx.stats()["big_bins"] >= 1
        check_vs_oracle_u64(ctx, keys, vals, 4)


def test_commit_replaces_abort_discards_and_empty_shuffle():
    """job.lua:217-221 (remove_file + build), worker.lua:120-127 (BROKEN job), server.lua:300-324"""
    with mrhbm.Ctx(mrhbm.KEY_STR, 15, mrhbm.PART_FNV_LUA) as ctx:
        ctx.shuffle()  # nothing committed
        assert ctx.partitions() == [] and ctx.result_info().groups == 0
        m = ctx.map_begin(1)
        m.emit(b"a", 1)
        m.emit(b"a", 1)
        m.commit()
        m = ctx.map_begin(2)
        m.emit(b"zzz", 9)
        m.abort()
        m = ctx.map_begin(1)  # re-execution of job 1 replaces its output
        m.emit(b"a", 5)
        m.commit()
        m = ctx.map_begin(3)  # a job that emits nothing
        m.commit()
        ctx.shuffle()
        assert ctx.partitions() == [10]  # "a" -> partition 10
        assert list(ctx.groups(10)) == [(b"a", [5])]
        m = ctx.map_begin(4)
        m.emit(b"x" * 28, 1)  # does not fit the 32-byte record class: kept on the host (test_keys_longer_than_a_record_slot)
        with pytest.raises(mrhbm.MrhbmError):
            m.emit(b"x" * ((1 << 20) + 1), 1)  # beyond the 1 MB limit
        m.abort()
        ctx.shuffle()
        assert list(ctx.groups(10)) == [(b"a", [5])]  # the aborted job left nothing, on the device or beside it
        ctx.reset()
        ctx.shuffle()
        assert ctx.partitions() == []


def test_second_shuffle_after_overflow_skips_the_optimistic_layout():
    """a hot key overfills its fixed-capacity bin: the first shuffle pays a discarded optimistic attempt and
    lands on the exact layout (k_big_bins), later shuffles of the ctx go there directly"""
    keys, vals = O.gen_u64(SEED, 5, 200_000)
    keys = keys.copy()
    keys[::3] = np.uint64(0x1234567890ABCDEF)
    with mrhbm.Ctx(mrhbm.KEY_U64, 4) as ctx:
        for it in range(2):  # "loop" iterations of one task (server.lua:386-404)
            ctx.reset()
            m = ctx.map_begin("hot")
            m.emit_batch(u64_records(keys, vals))
            m.commit()
            ctx.shuffle()
            st = ctx.stats()
            assert st["attempts"] == (2 if it == 0 else 1) and st["big_bins"] >= 1
            check_vs_oracle_u64(ctx, keys, vals, 4)


def test_combiner_with_many_distinct_keys():
    """duplicate-heavy strings, 2^20 possible keys: shared-memory tables + the L2-resident global table collapse the
    stream to about one record per distinct key before the partition / sort / reduce stages (the config 3 path)"""
    n, P = 6_000_000, 15
    table = synth.zipf_table(1 << 20)
    recs = O.gen_zipf_rec32(SEED, 0, n, table).view(mrhbm.record_dtype(mrhbm.KEY_STR, 27)).reshape(-1)
    with mrhbm.Ctx(mrhbm.KEY_STR, P, mrhbm.PART_FNV_LUA, combiner=True) as ctx:
        m = ctx.map_begin(1)
        m.gen_zipf(SEED, 0, n, table)
        m.commit()
        ctx.shuffle()
        st, info = ctx.stats(), ctx.result_info()
        assert st["attempts"] == 1 and st["ms_combine"] > 0 and st["ms_hist"] < 0.05, st
        assert info.groups <= info.pairs_recv <= info.groups + info.groups // 50  # (a key may own two table entries)
        check_vs_oracle_str(ctx, recs, P, O.PART_FNV_LUA)


def test_properties_at_10_pow_7():
    """size-independent parity: sum linearity, strict ascending order, partition membership"""
    n, P = 10_000_000, 1024
    with mrhbm.Ctx(mrhbm.KEY_U64, P, reserve_pairs=n) as ctx:
        m = ctx.map_begin("big")
        m.gen_u64(SEED, 0, n)
        m.commit()
        ctx.shuffle()
        cin, cout = ctx.checksum_input(), ctx.checksum_result()
        assert cin[:3] == cout[:3] and cin[3] == n and cout[4:] == [0, 0]
        keys, vals = O.gen_u64(SEED, 0, n)
        assert cout[3] == np.unique(keys).size and cin[2] == int(vals.astype(np.uint64).sum())
        assert ctx.result_info().sorted == 1


@pytest.mark.parametrize("kind", ["zipf", "zipfbig", "u64dup"])
def test_map_side_combiner_keeps_the_result(kind):
    """combinerfn == reducefn (job.lua:92-96,198-202): the map-side combine changes nothing"""
    n, P = 1_500_000, 15
    if kind in ("zipf", "zipfbig"):
        table = synth.zipf_table(1 << 16)
        recs = O.gen_zipf_rec32(SEED, 0, n, table).view(mrhbm.record_dtype(mrhbm.KEY_STR, 27)).reshape(-1)
        if kind == "zipfbig":  # pairs x largest value >= 2^32: the combiner must switch to checked adds; zeros and
            recs = recs.copy()  # values > 0xffff bypass the shared-memory tables
            recs["val"] = np.random.default_rng(3).integers(0, 3000, n).astype(np.uint32)
            recs["val"][::97] = 70000
        with mrhbm.Ctx(mrhbm.KEY_STR, P, mrhbm.PART_FNV_LUA, combiner=True) as ctx:
            m = ctx.map_begin(1)
            if kind == "zipf":
                m.gen_zipf(SEED, 0, n, table)
            else:
                m.emit_batch(recs)
            m.commit()
            ctx.shuffle()
            st = ctx.stats()
            info = ctx.result_info()
            assert info.pairs_in == n and info.pairs_recv < n // 2  # hot keys collapsed before partitioning
            assert st["ms_combine"] > 0
            check_vs_oracle_str(ctx, recs, P, O.PART_FNV_LUA)
    else:
        rng = np.random.default_rng(9)
        keys = O.gen_u64(SEED, 0, 5000)[0][rng.integers(0, 5000, n)]
        vals = rng.integers(0, 70000, n).astype(np.uint32)  # values > 0xffff and 0 bypass the table
        vals[::11] = 0
        with mrhbm.Ctx(mrhbm.KEY_U64, P, combiner=True) as ctx:
            m = ctx.map_begin(1)
            m.emit_batch(u64_records(keys, vals))
            m.commit()
            ctx.shuffle()
            assert ctx.result_info().pairs_recv < n
            check_vs_oracle_u64(ctx, keys, vals, P)


def test_group_only_mode_for_general_reducers():
    """MRHBM_RED_NONE (job.lua:264-284 with an arbitrary reducefn): the device sorts and groups, the
    caller sees every value of a key.  Checked against the oracle's identity reducer."""
    rng = np.random.default_rng(21)
    n, P = 120_000, 7
    words = [O.rank_to_key(int(r)) for r in rng.integers(1, 40_000, n)]
    vals = rng.integers(0, 1 << 31, n).astype(np.uint32)
    e = O.Engine(O.PART_FNV_LUA, P, combiner=-1, reducer=O.RED_IDENTITY, aci=False)
    for j in range(3):
        e.map_job(j, pairs=[(w, int(v)) for w, v in zip(words[j::3], vals[j::3])])
    e.reduce_all()
    want = {(p, k): sorted(int(x) for x in v) for p, k, v in e.final_pairs()}
    with mrhbm.Ctx(mrhbm.KEY_STR, P, mrhbm.PART_FNV_LUA, reducer=mrhbm.RED_NONE) as ctx:
        for j in range(3):
            m = ctx.map_begin(j)
            m.emit_batch(str_records(words[j::3], vals[j::3], 27))
            m.commit()
        ctx.shuffle()
        got = {}
        for p in ctx.partitions():
            ks = []
            for k, v in ctx.groups(p):
                got[(p, k)] = sorted(v)
                ks.append(k)
            assert ks == sorted(ks) and len(set(ks)) == len(ks)
        assert got == want
        cin, cout = ctx.checksum_input(), ctx.checksum_result()
        assert cin[:3] == cout[:3] and cout[3] == n and cout[4:] == [0, 0]
    with pytest.raises(mrhbm.MrhbmError):  # a general reducer cannot be combined on the device
        mrhbm.Ctx(mrhbm.KEY_STR, P, mrhbm.PART_FNV_LUA, reducer=mrhbm.RED_NONE, combiner=True)
    with mrhbm.Ctx(mrhbm.KEY_U64, 2, reducer=mrhbm.RED_NONE) as ctx:  # more values for one key than a bin holds
        keys = np.concatenate([np.full(5000, 7, dtype=np.uint64), np.arange(100, 400, dtype=np.uint64)])
        m = ctx.map_begin(1)
        m.emit_batch(u64_records(keys, np.arange(keys.size, dtype=np.uint32)))
        m.commit()
        ctx.shuffle()
        got = {int.from_bytes(k, "big"): sorted(v) for p in ctx.partitions() for k, v in ctx.groups(p)}
        assert got[7] == list(range(5000)) and len(got) == 301 and got[100] == [5000]


def test_device_tokeniser_wordcount_golden(golden_vectors, golden_wordcount):
    """device-side mapfn (examples/WordCount/mapfn.lua:3-9): text -> (word, 1) with C-locale isspace"""
    seps = [b" ", b"\n", b"\t", b"  \n", b"\r\n", b"\v", b"\f "]
    with mrhbm.Ctx(mrhbm.KEY_STR, 15, mrhbm.PART_FNV_LUA, max_key_bytes=golden_vectors["max_key_len"]) as ctx:
        total = 0
        for job in range(4):
            toks = expand_tokens(golden_wordcount, job)
            text = b"\n \t" + b"".join(t + seps[i % len(seps)] for i, t in enumerate(toks[:-1])) + toks[-1]
            m = ctx.map_begin(job + 1)
            total += m.wordcount(text)
            m.wordcount(b"")
            m.wordcount(b" \n\t ")
            m.commit()
        assert total == golden_vectors["tokens"]
        ctx.shuffle()
        got = {k: (p, v[0]) for p in ctx.partitions() for k, v in ctx.groups(p)}
        assert got == {k: (p, sum(c)) for k, p, c in golden_wordcount}
    # a longer stream against the oracle's tokeniser + naive count
    table = synth.zipf_table(1 << 12)
    recs = O.gen_zipf_rec32(SEED, 0, 300_000, table)
    words = [bytes(r[:28]).rstrip(b"\0") for r in recs]
    text = b"".join(w + (b"\n" if i % 25 == 24 else b" ") for i, w in enumerate(words))
    ntok, wc = O.naive_wordcount([text])
    with mrhbm.Ctx(mrhbm.KEY_STR, 15, mrhbm.PART_FNV_LUA) as ctx:
        m = ctx.map_begin("t")
        assert m.wordcount(text) == ntok == 300_000
        m.commit()
        ctx.shuffle()
        assert sorted((k, v[0]) for p in ctx.partitions() for k, v in ctx.groups(p)) == wc
        m = ctx.map_begin("long")
        m.emit(b"ok", 1)
        assert m.wordcount(b"fits " + b"x" * 28 + b" tail") == 3  # a word longer than the slot: this piece goes through the host
        m.commit()
        ctx.shuffle()
        got = dict((k, v[0]) for p in ctx.partitions() for k, v in ctx.groups(p))
        assert got[b"ok"] == 1 and got[b"x" * 28] == 1 and got[b"fits"] == 1 and got[b"tail"] == 1
