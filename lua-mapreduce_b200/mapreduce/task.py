"""In-process task board replacing the MongoDB `task` singleton and job collections
(mapreduce/task.lua:96-116,258-343).  Job documents keep the reference's fields and state
machine (WAITING -> RUNNING -> FINISHED -> WRITTEN | BROKEN -> FAILED, utils.lua:33-40)."""
import threading
import time

from .utils import STATUS, TASK_STATUS, MAX_JOB_RETRIES

_boards = {}
_boards_lock = threading.Lock()


def board(connection_string, dbname):
    """One board per (connection string, db) -- what `cnn(...)` + dbname select in the reference."""
    with _boards_lock:
        return _boards.setdefault((connection_string, dbname), Board())


def make_job(key, value):
    """mapreduce/utils.lua:87-98"""
    assert key is not None and value is not None, "Needs a key and a value"
    return {"_id": str(key), "value": value, "worker": "<unknown>", "tmpname": "<NONE>",
            "creation_time": time.time(), "status": STATUS.WAITING, "repetitions": 0}


class Board:
    def __init__(self):
        self.cv = threading.Condition()
        self.status = TASK_STATUS.WAIT
        self.iteration = 0
        self.jobs = {"map_jobs": [], "red_jobs": []}
        self.errors = []
        self.config = None   # what server:configure stored in the task document
        self.ctx = None      # the HBM shuffle context shared by server and workers
        self.workers = 0

    def current_ns(self):
        return {TASK_STATUS.MAP: "map_jobs", TASK_STATUS.REDUCE: "red_jobs"}.get(self.status)

    def take_next_job(self, worker_name):
        """task.lua:258-343: claim a WAITING (or BROKEN) job, mark it RUNNING."""
        with self.cv:
            ns = self.current_ns()
            if ns is None:
                return None, None
            for j in self.jobs[ns]:
                if j["status"] in (STATUS.WAITING, STATUS.BROKEN) and j["repetitions"] < MAX_JOB_RETRIES:
                    j.update(status=STATUS.RUNNING, worker=worker_name, started_time=time.time())
                    return ns, j
            return ns, None

    def mark(self, job, status, **extra):
        with self.cv:
            job["status"] = status
            job.update(extra)
            self.cv.notify_all()

    def mark_as_broken(self, job, msg):
        """job.lua:322-342 + cnn.lua:62-78 + server.lua:194-213 (BROKEN x3 -> FAILED)"""
        with self.cv:
            job["repetitions"] += 1
            job["status"] = STATUS.FAILED if job["repetitions"] >= MAX_JOB_RETRIES else STATUS.BROKEN
            self.errors.append({"worker": job.get("worker"), "msg": msg})
            self.cv.notify_all()

    def pending(self, ns):
        return [j for j in self.jobs[ns] if j["status"] not in (STATUS.WRITTEN, STATUS.FAILED)]
