"""mapreduce/worker.lua mirrored: poll loop (worker.lua:42-105) + xpcall retry wrapper
(worker.lua:112-138).  Workers of one (connection string, db) share the server's HBM ctx."""
import socket
import sys
import time
import traceback

from . import task as _task
from .job import Job
from .utils import MAX_WORKER_RETRIES, TASK_STATUS


class worker:
    def __init__(self, connection_string, dbname, auth_table=None):
        self.board = _task.board(connection_string, dbname)
        self.max_iter, self.max_sleep, self.max_tasks = 20, 20, 1  # worker.lua:160-162
        self.name = "%s:%d" % (socket.gethostname(), id(self) & 0xffff)
        self.current_job = None
        self.poll_sleep = 0.01  # the reference polls MongoDB every 1 s (utils.lua:28)

    @staticmethod
    def new(connection_string, dbname, auth_table=None):
        return worker(connection_string, dbname, auth_table)

    def configure(self, t):
        for k, v in t.items():  # worker.lua:142-148
            assert k in ("max_iter", "max_sleep", "max_tasks"), "Unknown parameter: %s\n" % k
            setattr(self, k, v)

    def _execute_once(self):
        """worker.lua:42-105: take jobs until the task is FINISHED (or max_iter idle polls)."""
        b = self.board
        idle, ntasks, job_done = 0, 0, False
        while idle < self.max_iter and ntasks < self.max_tasks:
            with b.cv:
                status, cfg = b.status, b.config
            if status == TASK_STATUS.FINISHED:
                if job_done:
                    ntasks += 1
                    job_done = False
                    continue
                break
            ns, doc = b.take_next_job(self.name) if cfg else (None, None)
            if doc is None:
                idle += 1
                time.sleep(min(self.max_sleep, self.poll_sleep * idle))
                continue
            idle = 0
            self.current_job = Job(b, ns, doc, cfg)
            self.current_job.execute()
            self.current_job = None
            job_done = True

    def execute(self):
        """worker.lua:112-138.  The worker registers itself on the board for the time it runs: the server only
        executes jobs inline while no worker is attached, so the two never drive the shared ctx at once (the C
        handle additionally serialises its callers with a per-ctx lock)."""
        with self.board.cv:
            self.board.workers += 1
        try:
            self._execute_with_retries()
        finally:
            with self.board.cv:
                self.board.workers -= 1
                self.board.cv.notify_all()

    def _execute_with_retries(self):
        failed = set()
        while True:
            try:
                self._execute_once()
                break
            except Exception:
                msg = traceback.format_exc()
                if self.current_job is not None:
                    self.board.mark_as_broken(self.current_job.doc, msg)
                    failed.add(self.current_job.get_id())
                    self.current_job = None
                sys.stderr.write("Error executing a job: %s\n" % msg)
                if len(failed) >= MAX_WORKER_RETRIES:
                    print("# Worker retries: %d" % len(failed))
                    raise RuntimeError("Maximum number of retries achieved")
        print("# Worker retries: %d" % len(failed))


new = worker.new
