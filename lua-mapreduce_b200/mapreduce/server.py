"""mapreduce/server.lua mirrored: configure (server.lua:419-462), loop (466-611),
server_prepare_map (249-276), server_prepare_reduce (279-329), server_final (348-413).
The MongoDB task/job collections are an in-process board; GridFS spill files are the HBM
shuffle (storage "hbm").  Jobs run on attached workers or, when none is attached, inline."""
import importlib
import sys
import time

from . import task as _task
from .job import Job, get_func
from .utils import STATUS, TASK_STATUS, get_storage_from, count_digits

_BUILTIN_PART = {"fnv_lua": 0, "mulhash": 1, "wordhash": 2}


class server:
    def __init__(self, connection_string, dbname, auth_table=None):
        self.cnn_string, self.dbname = connection_string, dbname
        self.board = _task.board(connection_string, dbname)
        self.configured = False
        self.finished = False
        self.ctx_factory = None  # tests inject a stand-in; default = the CUDA ctx

    @staticmethod
    def new(connection_string, dbname, auth_table=None):
        return server(connection_string, dbname, auth_table)

    # ---- server.lua:419-462
    def configure(self, params):
        params = dict(params)
        storage, path = get_storage_from(params.get("storage"), True)
        params["storage"] = "%s:%s" % (storage, path)
        self.result_ns = params.get("result_ns") or "result"
        assert all(params.get(k) for k in ("taskfn", "mapfn", "partitionfn", "reducefn")), \
            "Fields taskfn, mapfn, partitionfn and reducefn are mandatory"
        mods = {}
        for name in ("taskfn", "mapfn", "partitionfn", "reducefn", "finalfn", "combinerfn"):
            v = params.get(name)
            assert (isinstance(v, str)) or (not v and name in ("finalfn", "combinerfn")), \
                "Needs a %s module with %s function" % (name, name)
            if v:
                aux = importlib.import_module(v)
                assert hasattr(aux, name), "Module %s must return a table with the field %s" % (name, name)
                assert hasattr(aux, "init"), "Init function is needed: %s" % name
                mods[name] = aux
        self.init_args = params.get("init_args")
        self.taskfn = mods["taskfn"]
        self.finalfn = mods.get("finalfn")
        for init in {m.init for m in (self.taskfn, self.finalfn) if m is not None}:
            init(self.init_args)
        # ---- what the device evaluates: declared on the plugin modules (same pattern as the
        # associative/commutative/idempotent flags, examples/WordCount/reducefn.lua:10-14)
        red, part = mods["reducefn"], mods["partitionfn"]
        builtin_red = getattr(red, "hbm_reducefn", None)
        if builtin_red not in (None, "sum"):
            raise NotImplementedError("unknown hbm_reducefn %r (built-ins: 'sum')" % (builtin_red,))
        # no declaration = general reducer: the device partitions, sorts and groups, reducefn runs
        # on the host per group (job.lua:264-284); values are unsigned integers < 2^32
        # A combinerfn only ever runs on the device, as the declared built-in (it must equal the reducefn there).
        # Any other combinerfn is SKIPPED: the reference's contract for it (job.lua:92-96,198-202) is that reducing
        # combined values equals reducing the raw ones, so the reducefn -- on the device if it is the built-in, per
        # group on the host otherwise -- sees the raw values and produces the same result, just without the saving.
        device_combiner = "combinerfn" in mods and builtin_red == "sum" and \
            getattr(mods["combinerfn"], "hbm_reducefn", None) == "sum"
        if "combinerfn" in mods and not device_combiner:
            sys.stderr.write("# WARNING: combinerfn %s is not a device built-in (hbm_reducefn = 'sum'); it is skipped, "
                             "the reducefn sees the uncombined values\n" % params["combinerfn"])
        pname = getattr(part, "hbm_partitionfn", None)
        if pname not in _BUILTIN_PART:
            raise NotImplementedError("partitionfn module must declare hbm_partitionfn in %s" % sorted(_BUILTIN_PART))
        nparts = getattr(part, "NUM_REDUCERS", None) or getattr(part, "hbm_num_partitions", None)
        assert isinstance(nparts, int) and nparts >= 1, "partitionfn module must expose NUM_REDUCERS"
        hbm = dict(key_kind="str", max_key_bytes=123, device=-1)
        hbm.update(params.get("hbm") or {})
        hbm.update(partitioner=_BUILTIN_PART[pname], num_partitions=nparts, combiner=device_combiner,
                   reducer=0 if builtin_red == "sum" else 1)
        self.config = dict(mapfn=params["mapfn"], reducefn=params["reducefn"], partitionfn=params["partitionfn"],
                           combinerfn=params.get("combinerfn"), init_args=self.init_args,
                           storage=params["storage"], hbm=hbm)
        self.configured = True

    def _make_ctx(self):
        h = self.config["hbm"]
        if self.ctx_factory is not None:
            return self.ctx_factory(h)
        from .. import mrhbm
        kind = mrhbm.KEY_U64 if h["key_kind"] == "u64" else mrhbm.KEY_STR
        return mrhbm.Ctx(kind, h["num_partitions"], h["partitioner"], max_key_bytes=h["max_key_bytes"],
                         combiner=h["combiner"], device=h["device"], reducer=h["reducer"])

    # ---- polling (server.lua:186-234) with inline execution when no worker is attached
    def _wait(self, ns):
        b = self.board
        while True:
            with b.cv:
                pend = b.pending(ns)
                if not pend:
                    break
                runnable = [j for j in pend if j["status"] in (STATUS.WAITING, STATUS.BROKEN)]
                if b.workers > 0 or not runnable:
                    b.cv.wait(timeout=0.05)
                    continue
            ns2, doc = b.take_next_job("server-inline")
            if doc is None:
                continue
            job = Job(b, ns2, doc, self.config)
            try:
                job.execute()
            except Exception as e:  # worker.lua:116-131
                b.mark_as_broken(doc, repr(e))
                sys.stderr.write("Error executing a job: %r\n" % (e,))
        return sum(1 for j in b.jobs[ns] if j["status"] == STATUS.FAILED)

    # ---- server.lua:466-611
    def loop(self):
        assert self.configured, "Call to server:configure(...) method is mandatory"
        b = self.board
        it = 0
        while True:
            it += 1
            t0 = time.time()
            with b.cv:
                if b.ctx is None or getattr(b.ctx, "h", True) is None:  # none yet, or closed by its owner
                    b.ctx = self._make_ctx()
                else:
                    b.ctx.reset()
                b.config, b.iteration = self.config, it
                b.jobs = {"map_jobs": [], "red_jobs": []}
                # server_prepare_map (server.lua:249-276)
                seen = set()

                def emit(key, value):
                    assert key not in seen, "Duplicate key: %s" % (key,)
                    seen.add(key)
                    b.jobs["map_jobs"].append(_task.make_job(key, value))
                self.taskfn.taskfn(emit)
                b.status = TASK_STATUS.MAP
                b.cv.notify_all()
            sys.stderr.write("# Iteration %d\n# \t Preparing MAP\n" % it)
            failed_maps = self._wait("map_jobs")
            # server_prepare_reduce (server.lua:279-329): one reduce job per NON-EMPTY partition
            b.ctx.shuffle()
            parts = b.ctx.partitions()
            digits = count_digits(max(parts) if parts else 0)
            with b.cv:
                for p in parts:
                    b.jobs["red_jobs"].append(_task.make_job(p, {
                        "file": "map_results.P%d" % p, "result": "%s.P%0*d" % (self.result_ns, digits, p)}))
                b.status = TASK_STATUS.REDUCE
                b.cv.notify_all()
            sys.stderr.write("# \t Preparing REDUCE\n")
            failed_reds = self._wait("red_jobs")
            self.stats = {"iteration": it, "map_count": len(b.jobs["map_jobs"]), "reduce_count": len(parts),
                          "failed_maps": failed_maps, "failed_reds": failed_reds,
                          "server_time": time.time() - t0, "shuffle": b.ctx.stats()}
            sys.stderr.write("# Server time %f\n" % self.stats["server_time"])
            reply = self._final()
            if reply != "loop":
                with b.cv:
                    b.status = TASK_STATUS.FINISHED
                    b.cv.notify_all()
                self.finished = True
                return

    # ---- server.lua:348-413
    def _final(self):
        b = self.board
        files = sorted(b.jobs["red_jobs"], key=lambda j: j["value"]["result"])  # server.lua:367

        def pair_iterator():
            for j in files:
                for key, values in j["value"].get("pairs", ()):
                    yield key, values

        self.results = files
        reply = self.finalfn.finalfn(pair_iterator()) if self.finalfn is not None else None
        if reply not in ("loop", True, False, None):
            sys.stderr.write("# WARNING!!! INCORRECT FINAL RETURN: %s\n" % (reply,))
        if reply == "loop":
            sys.stderr.write("# LOOP again\n")
        if reply is True or reply == "loop":  # remove results
            for j in files:
                j["value"].pop("pairs", None)
        return reply


    # ---- interchange with a stock lua-mapreduce deployment
    def export_results(self, directory):
        """Writes the kept results (finalfn returned false/nil) as the reference's result files: one file per
        non-empty partition named `<result_ns>.P<kk>` (server.lua:313-321), one line `return <key>,{<values>}`
        per key in ascending key order (job.lua:272-273).  Returns the file names."""
        import os
        from .utils import result_line
        names = []
        for j in getattr(self, "results", []):
            pairs = j["value"].get("pairs")
            if pairs is None:
                continue
            name = j["value"]["result"]
            with open(os.path.join(directory, name), "wb") as fh:
                for key, values in pairs:
                    fh.write(result_line(key, values))
            names.append(name)
        return names


new = server.new
