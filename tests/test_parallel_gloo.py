"""N>1 host path on CPU: two gloo ranks (see tests/gloo_check.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_gloo_host_logic():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tests", "gloo_check.py")],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "GLOO_CHECK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_ownership_helpers():
    import mrhbm_loader
    mrhbm_loader.load()
    from lua_mapreduce_b200 import parallel
    assert parallel.owned_partitions(1, 4, 10) == [1, 5, 9]
    assert sorted(p for r in range(8) for p in parallel.owned_partitions(r, 8, 1024)) == list(range(1024))
    assert parallel.owned_partitions(5, 8, 3) == []
