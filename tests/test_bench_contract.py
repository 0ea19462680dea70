"""bench.py contract pieces that run without a GPU: the reference arm (oracle on the host cores), its
behaviour under torchrun-style environments, and that the product arm has no CPU fallback."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "100000"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "kv_pairs_per_sec_shuffle_sort_reduce" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "config3" in d["config"]["workload"] and d["warmup"] == 1


def test_reference_arm_other_ranks_exit_quietly():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-sample", "50000"],
            env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            return  # GPU box: covered by the gpu-marked tests and the bench itself
    except Exception:
        pass
    r = run(["--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--e2e-steps", "0"], timeout=600)
    assert r.returncode != 0
    assert r.stdout.strip() == ""  # no JSON line, no silent CPU path


def test_bench_main_dry_run_assembles_the_contract_line():
    """host logic of the product arm over stand-ins (no measurement meaning): every contract key is present"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_dryrun.py")], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "gpu_launches", "clocks", "e2e", "roofline", "cpu_baseline",
              "parity_vs_oracle", "configs", "run"):
        assert k in d, k
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic", "kernels", "pipeline", "fusion_headroom"}
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert d["config"]["workload"].startswith("config3") and d["vs_baseline"] is None
    assert set(d["configs"]) == {"config2_u64_1e8", "config4_u64_1e9_total", "config1_wordcount_197x10k"}
    for blk in (d["configs"]["config2_u64_1e8"], d["configs"]["config4_u64_1e9_total"]):
        assert set(blk) >= {"value", "ms_per_step", "roofline", "e2e", "config", "parity_vs_oracle"}
