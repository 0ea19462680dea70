"""The reference-facing task API (server / worker / plugin modules) mirrored in
lua-mapreduce_b200/mapreduce: test.sh's four plugin configurations against misc/naive.lua's
answer (the golden word count).  CPU tests run the host logic over a stand-in ctx (tests only);
the gpu-marked tests run the same scripts over the CUDA path."""
import threading
import time

import pytest

import mrhbm_loader
from conftest import expand_tokens

mrhbm_loader.load()
from lua_mapreduce_b200.mapreduce import server, worker, utils, task  # noqa: E402
from lua_mapreduce_b200.mapreduce.examples.WordCount import init as WordCount  # noqa: E402

WC = "lua_mapreduce_b200.mapreduce.examples.WordCount"
WCI = WC + ".init"


class StandInCtx:
    """Dict-based stand-in for mrhbm.Ctx so that the HOST logic is testable without a GPU.
    Lives in tests/ only; the product has no CPU path."""

    class _Map:
        def __init__(self, ctx, job):
            self.ctx, self.job, self.buf = ctx, job, []

        def emit(self, k, v=1):
            self.buf.append((k, v))

        def commit(self):
            self.ctx.jobs[self.job] = self.buf

        def abort(self):
            self.buf = None

    def __init__(self, h):
        self.P, self.jobs, self.parts = h["num_partitions"], {}, {}

    def map_begin(self, job):
        return self._Map(self, str(job))

    def reset(self):
        self.jobs, self.parts = {}, {}

    def shuffle(self):
        self.parts = {}
        for buf in self.jobs.values():
            for k, v in buf:
                d = self.parts.setdefault(WordCount.partitionfn(k), {})
                d[k] = d.get(k, 0) + v

    def partitions(self):
        return sorted(self.parts)

    def groups(self, p):
        for k in sorted(self.parts.get(p, {})):
            yield k, [self.parts[p][k]]

    def stats(self):
        return {}


@pytest.fixture
def corpus(tmp_path, golden_wordcount):
    files = []
    for job in range(4):
        toks = expand_tokens(golden_wordcount, job)
        p = tmp_path / ("file%d.txt" % job)
        lines = [b" ".join(toks[i:i + 7]) for i in range(0, len(toks), 7)]
        p.write_bytes(b"\n".join(lines) + b"\n\t \n")
        files.append(str(p))
    WordCount.FILES[:] = files
    WordCount.RESULT.clear()
    return files


class BulkStandInCtx(StandInCtx):
    """A stand-in that also offers the bulk result calls; `long_keys` makes result_copy answer the way the library
    does when the result holds keys longer than a slot (MRHBM_E_KEY = -4)."""
    long_keys = False
    copies = 0

    def result_info(self):
        class Info:
            sorted = 1
        return Info()

    def result_copy(self):
        import numpy as np
        type(self).copies += 1
        if self.long_keys:
            e = RuntimeError("mrhbm error -4: groups have keys longer than the slots of mrhbm_result_copy")
            e.code = -4
            raise e
        keys, sums, po = [], [], [0]
        for p in range(self.P):
            for k in sorted(self.parts.get(p, {})):
                keys.append(k)
                sums.append(self.parts[p][k])
            po.append(len(keys))
        return np.array(keys, dtype="S123"), np.array(sums, dtype=np.uint64), np.array(po, dtype=np.uint64)


CONFIGS = {  # test.sh:9-71
    "combiner+aci": dict(reducefn=WC + ".reducefn", combinerfn=WC + ".reducefn"),
    "aci": dict(reducefn=WC + ".reducefn"),
    "general": dict(reducefn=WC + ".reducefn2"),
    "init-script": dict(taskfn=WCI, mapfn=WCI, partitionfn=WCI, reducefn=WCI, finalfn=WCI, combinerfn=WCI),
}


def run_config(name, dbname, ctx_factory=None, with_worker=False):
    params = dict(taskfn=WC + ".taskfn", mapfn=WC + ".mapfn", partitionfn=WC + ".partitionfn",
                  finalfn=WC + ".finalfn", storage="hbm")
    params.update(CONFIGS[name])
    s = server.new("hbm://local", dbname)
    s.ctx_factory = ctx_factory
    s.configure(params)
    th = None
    if with_worker:
        w = worker.new("hbm://local", dbname)
        w.configure({"max_iter": 2000, "max_tasks": 1})
        th = threading.Thread(target=w.execute, daemon=True)
        th.start()
        for _ in range(2000):  # the worker registers itself on the board (worker.execute); wait for that
            if s.board.workers > 0:
                break
            time.sleep(0.001)
        assert s.board.workers == 1
    s.loop()
    if th:
        th.join(timeout=60)
        assert not th.is_alive()
    return s


@pytest.mark.parametrize("name", list(CONFIGS))
def test_wordcount_configs_host_logic(name, corpus, golden_wordcount):
    s = run_config(name, "cpu-" + name, ctx_factory=StandInCtx)
    assert WordCount.RESULT == {k: sum(c) for k, _, c in golden_wordcount}
    assert s.stats["map_count"] == 4 and s.stats["reduce_count"] == 15 and s.stats["failed_maps"] == 0
    assert [j["value"]["result"] for j in s.results] == ["result.P%02d" % p for p in range(15)]


def test_worker_thread_takes_the_jobs(corpus, golden_wordcount):
    s = run_config("aci", "cpu-worker", ctx_factory=StandInCtx, with_worker=True)
    assert WordCount.RESULT == {k: sum(c) for k, _, c in golden_wordcount}
    assert {j["worker"] for j in s.board.jobs["map_jobs"]} != {"server-inline"}


def test_configure_errors_and_storage_strings():
    s = server.new("hbm://local", "cfg")
    with pytest.raises(AssertionError, match="mandatory"):
        s.configure({"taskfn": WC})
    with pytest.raises(ValueError, match="Given incorrect storage"):
        utils.get_storage_from("floppy")
    assert utils.get_storage_from("hbm:/tmp/x") == ("hbm", "/tmp/x")
    assert utils.get_storage_from(None)[0] == "hbm"
    with pytest.raises(AssertionError, match="Call to server:configure"):
        server.new("hbm://local", "cfg2").loop()
    w = worker.new("hbm://local", "cfg")
    with pytest.raises(AssertionError, match="Unknown parameter"):
        w.configure({"bogus": 1})


def test_wire_format_helpers(golden_vectors):  # utils.lua:345-349
    for v, want in golden_vectors["escape"]:
        assert utils.escape(v) == want.encode()
    for vals, want in golden_vectors["serialize_table_ipairs"]:
        assert utils.serialize_table_ipairs(vals) == want.encode()
    assert utils.result_line(b"a", [3]) == b'return "a",{3}\n'
    for n, d in golden_vectors["count_digits"]:
        assert utils.count_digits(n) == d


def test_broken_map_job_is_retried_then_failed(corpus):
    """worker.lua:112-138 / server.lua:194-213: a job that raises is BROKEN, retried, and
    counted FAILED after MAX_JOB_RETRIES; its partial emits never become visible."""
    import types, sys
    mod = types.ModuleType("bad_mapfn")
    calls = []

    def mapfn(key, value, emit):
        calls.append(key)
        emit(b"ghost", 1)
        if key == "2":
            raise RuntimeError("boom")
        emit(b"ok", 1)
    mod.mapfn, mod.init = mapfn, lambda a=None: None
    sys.modules["bad_mapfn"] = mod
    s = server.new("hbm://local", "cpu-broken")
    s.ctx_factory = StandInCtx
    s.configure(dict(taskfn=WC + ".taskfn", mapfn="bad_mapfn", partitionfn=WC + ".partitionfn",
                     reducefn=WC + ".reducefn", finalfn=WC + ".finalfn", storage="hbm"))
    s.loop()
    assert calls.count("2") == utils.MAX_JOB_RETRIES and s.stats["failed_maps"] == 1
    assert WordCount.RESULT == {b"ghost": 3, b"ok": 3}


@pytest.mark.gpu
def test_declared_device_mapfn(corpus, golden_wordcount):
    """mapfn module declaring hbm_mapfn = "wordcount_file": the device tokenises the files"""
    CONFIGS["device-mapfn"] = dict(mapfn=WC + ".mapfn_device", reducefn=WC + ".reducefn")
    try:
        s = run_config("device-mapfn", "gpu-device-mapfn")
    finally:
        del CONFIGS["device-mapfn"]
    assert WordCount.RESULT == {k: sum(c) for k, _, c in golden_wordcount}
    assert s.stats["shuffle"]["pairs"] == 4989
    s.board.ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CONFIGS))
def test_wordcount_configs_on_gpu(name, corpus, golden_wordcount):
    s = run_config(name, "gpu-" + name)
    assert WordCount.RESULT == {k: sum(c) for k, _, c in golden_wordcount}
    assert s.stats["reduce_count"] == 15 and s.stats["shuffle"]["pairs"] == 4989
    s.board.ctx.close()


def _general_reducer_module():
    import sys
    import types
    mod = types.ModuleType("max_reducefn")

    def reducefn(key, values, emit):  # not a sum: needs every value
        emit(max(values))
        emit(len(values))
    mod.reducefn, mod.init = reducefn, lambda a=None: None
    sys.modules["max_reducefn"] = mod
    return "max_reducefn"


@pytest.mark.gpu
def test_general_reducer_runs_on_the_host_over_device_groups(corpus, golden_wordcount):
    """a reducefn without hbm_reducefn: device groups, host reduces (job.lua:275-284)"""
    s = server.new("hbm://local", "gpu-max-reducer")
    s.configure(dict(taskfn=WC + ".taskfn", mapfn=WC + ".mapfn", partitionfn=WC + ".partitionfn",
                     reducefn=_general_reducer_module(), finalfn=WC + ".finalfn", storage="hbm"))
    assert s.config["hbm"]["reducer"] == 1
    got = {}
    s.finalfn = type("F", (), {"finalfn": staticmethod(lambda it: got.update({k: v for k, v in it}) or True)})
    s.loop()
    assert got == {k: [1, sum(c)] for k, _, c in golden_wordcount}
    s.board.ctx.close()


def test_undeclared_combinerfn_is_skipped_not_refused(corpus, golden_wordcount):
    """a combinerfn that is not the device built-in (no hbm_reducefn = 'sum'): configure() no longer raises; the
    combiner -- an optimisation whose contract is "reduce(combined) == reduce(raw)" (job.lua:92-96,198-202) -- is
    skipped and the reducefn sees the raw values"""
    import sys
    import types
    mod = types.ModuleType("my_combinerfn")
    mod.init = lambda a=None: None
    mod.combinerfn = lambda key, values, emit: emit(sum(values))
    sys.modules["my_combinerfn"] = mod
    s = server.new("hbm://local", "cpu-undeclared-combiner")
    s.ctx_factory = StandInCtx
    s.configure(dict(taskfn=WC + ".taskfn", mapfn=WC + ".mapfn", partitionfn=WC + ".partitionfn", reducefn=WC + ".reducefn",
                     combinerfn="my_combinerfn", finalfn=WC + ".finalfn", storage="hbm"))
    assert s.config["hbm"]["combiner"] is False and s.config["hbm"]["reducer"] == 0
    s.loop()
    assert WordCount.RESULT == {k: sum(c) for k, _, c in golden_wordcount}


def test_exported_result_files_equal_the_oracles(corpus, golden_wordcount, tmp_path):
    """server:export_results writes result.P<kk> files byte-identical to what the reference's reduce jobs
    write (oracle restatement of job.lua:272-273 / server.lua:313-321) for the test.sh corpus"""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle as O
    keep = WordCount.finalfn

    def finalfn_keep(it):  # finalfn returning false keeps the results (server.lua:386-404)
        keep(it)
        return False
    WordCount.finalfn = finalfn_keep
    try:
        s = run_config("init-script", "export-results", ctx_factory=StandInCtx)
    finally:
        WordCount.finalfn = keep
    out = tmp_path / "results"
    out.mkdir()
    names = s.export_results(str(out))
    e = O.Engine(O.PART_FNV_LUA, 15, combiner=O.RED_SUM, reducer=O.RED_SUM, aci=True)
    for job, path in enumerate(corpus):
        e.map_job(job + 1, text=open(path, "rb").read())
    e.reduce_all()
    want = {n: data for n, _, data in e.results()}
    e.close()
    assert sorted(names) == sorted(want)
    for n in names:
        assert (out / n).read_bytes() == want[n], n


def test_finalfn_loop_runs_another_iteration(corpus, golden_wordcount):
    """finalfn returning "loop" re-runs taskfn/map/reduce on a reset ctx (server.lua:386-404); results of the
    previous iteration are dropped, the final answer is the last iteration's"""
    calls = []
    keep = WordCount.finalfn

    def finalfn_loop_once(it):
        pairs = {k: v[0] for k, v in it}
        calls.append(pairs)
        return "loop" if len(calls) == 1 else True
    WordCount.finalfn = finalfn_loop_once
    try:
        s = run_config("init-script", "loop-task", ctx_factory=StandInCtx)
    finally:
        WordCount.finalfn = keep
    assert len(calls) == 2 and calls[0] == calls[1] and s.stats["iteration"] == 2 and s.finished
    want = {}
    for job in range(4):
        for t in expand_tokens(golden_wordcount, job):
            want[t] = want.get(t, 0) + 1
    assert calls[1] == want
    assert all("pairs" not in j["value"] for j in s.results)  # true / "loop": results removed (server.lua:395-401)


@pytest.mark.parametrize("long_keys", [False, True])
def test_reduce_falls_back_to_the_iterator_when_the_bulk_copy_cannot_hold_the_keys(corpus, golden_wordcount, long_keys):
    """job.py takes the whole result with ONE bulk copy when the built-in reducer + ACI flags allow it; a result with
    keys longer than a slot makes the library answer MRHBM_E_KEY, and the reduce jobs then iterate group by group
    (job.lua:264-284) -- same final result either way."""
    BulkStandInCtx.long_keys, BulkStandInCtx.copies = long_keys, 0
    s = run_config("aci", "bulk-%s" % long_keys, ctx_factory=BulkStandInCtx)
    want = {}
    for k, p, c in golden_wordcount:
        want[k] = sum(c)
    assert {k: v for k, v in WordCount.RESULT.items()} == want
    assert BulkStandInCtx.copies == 1  # once per shuffle: cached on the board, also when it failed

