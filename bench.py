#!/usr/bin/env python3
"""bench.py -- key-value pairs/s through shuffle + sort + reduce (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload u64|zipf32] [--pairs M]
  python bench.py --impl reference ...   # the reference-shaped CPU path (oracle/) on host cores

One "step" = one pass of the hot path (hash-partition + sort + segmented reduce) over one batch
of synthetic pairs.  `value` times the step with the batch already resident in HBM; `e2e` times
the same step through the public C-ABI calls a user makes, from pinned HOST buffers
(emit_batch -> commit -> shuffle -> result_copy), host<->device copies inside the timed region.
Workloads:
  u64    : config 2, 10^8 uniform u64 keys / u32 values, 16 B records, 1024 partitions (default)
  zipf32 : config 3, Zipf(1.1) words <= 27 B, 32 B records, value 1, 15 partitions
Inputs are far larger than the 126 MB L2, so no flush is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC = "kv_pairs_per_sec_shuffle_sort_reduce"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="u64", choices=["u64", "zipf32"])
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step (default: config size)")
    ap.add_argument("--partitions", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-serial", action="store_true", help="one worker, no overlap between steps")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="pairs in the CPU sample (default: 10^7 for cpu_baseline = ~10 s on 4 threads; "
                         "2*10^6 per step for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_params(a):
    if a.workload == "u64":
        return dict(name="config2: 1e8 uniform u64 keys / u32 values, 16 B records", rb=16,
                    pairs=a.pairs or 100_000_000, P=a.partitions or 1024)
    return dict(name="config3: Zipf(1.1) string keys <=27 B, 32 B records, value 1", rb=32,
                pairs=a.pairs or 1_000_000_000, P=a.partitions or 15)


def host_u64_records(seed, start, n, out):
    """numpy splitmix64 stream into a pinned record array."""
    from lua_mapreduce_b200.synth import splitmix64_np
    step = 1 << 22
    for a in range(0, n, step):
        b = min(n, a + step)
        i = np.arange(start + a, start + b, dtype=np.uint64)
        out["key"][a:b] = splitmix64_np(np.uint64(seed) + i)
        out["val"][a:b] = (splitmix64_np(np.uint64(seed) + np.uint64(1 << 40) + i) >> np.uint64(32)).astype(np.uint32)
        out["pad"][a:b] = 0


def cpu_baseline(a, wp, table, threads, sample):
    """Reference-shaped CPU path (oracle engine: emit table, sort, text spill, heap merge,
    sum) on a bounded sample of the same stream.  kind = "port" (the reference is Lua+MongoDB
    and cannot run in this image)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from lua_mapreduce_b200 import synth
    njobs = max(threads, 8)
    per = max(1, sample // njobs)
    if a.workload == "u64":
        eng = O.Engine(O.PART_MULHASH, wp["P"], combiner=O.RED_SUM, reducer=O.RED_SUM, aci=True)
        ms, rs = O.run_synthetic(eng, 0, synth.SEED, 0, per, njobs, threads)
    else:
        eng = O.Engine(O.PART_FNV_LUA, wp["P"], combiner=O.RED_SUM, reducer=O.RED_SUM, aci=True)
        ms, rs = O.run_synthetic(eng, 1, synth.SEED, 0, per, njobs, threads, table)
    n = per * njobs
    eng.close()
    return {"value": n / (ms + rs), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d pairs of the same stream as %d map jobs on %d worker threads "
                      "(map %.2f s + reduce %.2f s)" % (n, njobs, threads, ms, rs)}


def run_reference(a):
    wp = workload_params(a)
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import mrhbm_loader
    mrhbm_loader.load()
    from lua_mapreduce_b200 import synth
    table = synth.zipf_table() if a.workload == "zipf32" else None
    threads = os.cpu_count() or 1
    vals = []
    sample = a.cpu_sample or 2_000_000
    for _ in range(a.warmup and 1):
        cpu_baseline(a, wp, table, threads, sample)
    t0 = time.perf_counter()
    last = None
    for _ in range(max(1, a.steps)):
        last = cpu_baseline(a, wp, table, threads, sample)
        vals.append(last["value"])
    dt = time.perf_counter() - t0
    v = float(np.mean(vals))
    last["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
        "steps": max(1, a.steps), "warmup": a.warmup and 1, "ms_per_step": 1e3 * dt / max(1, a.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (Lua numbers) / bytes",
        "data": "synthetic", "config": {"workload": wp["name"], "partitions": wp["P"],
                                        "note": "reference-shaped CPU restatement (oracle/), not Lua+MongoDB"},
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    a = parse()
    # Libraries (NCCL prints its version) must not pollute stdout: the driver reads ONE JSON line.
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if a.impl == "reference":
        return run_reference(a)
    import torch
    import mrhbm_loader
    mrhbm_loader.load()
    from lua_mapreduce_b200 import mrhbm, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus > 1 or world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                                timeout=datetime.timedelta(seconds=180))
    else:
        dist = None
    torch.cuda.set_device(local)
    wp = workload_params(a)
    n, P, rb = wp["pairs"], wp["P"], wp["rb"]
    table = synth.zipf_table() if a.workload == "zipf32" else None
    kind = mrhbm.KEY_U64 if a.workload == "u64" else mrhbm.KEY_STR
    part = mrhbm.PART_MULHASH if a.workload == "u64" else mrhbm.PART_FNV_LUA
    ctx = mrhbm.Ctx(kind, P, part, max_key_bytes=27, device=local, reserve_pairs=n,
                    combiner=(a.workload == "zipf32"))  # the reference's headline run uses combiner = reducer
    if world > 1:
        uid = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident input (weak scaling: every GPU owns n pairs of disjoint counters)
    m = ctx.map_begin("resident")
    if a.workload == "u64":
        m.gen_u64(synth.SEED, rank * n, n)
    else:
        m.gen_zipf(synth.SEED, rank * n, n, table)
    m.commit()
    for _ in range(max(3, a.warmup)):
        ctx.shuffle()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(3):  # every rank (the shuffle is collective): lets nvidia-smi attach before the timed region
        ctx.shuffle()
    barrier()
    t0 = time.perf_counter()
    agg = {}
    launches = 0
    for _ in range(a.steps):
        ctx.shuffle()
        st = ctx.stats()
        launches += st["launches"]
        for k, v in st.items():
            if k.startswith("ms_"):
                agg[k] = agg.get(k, 0.0) + v
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    st = ctx.stats()
    groups = ctx.result_info().groups
    cin, cout = ctx.checksum_input(), ctx.checksum_result()
    if dist is not None:  # the sum is linear over the whole job: add the per-rank digests mod 2^64
        box = [None] * world
        dist.all_gather_object(box, (cin, cout))
        cin = [sum(b[0][i] for b in box) % 2**64 for i in range(4)]
        cout = [sum(b[1][i] for b in box) % 2**64 for i in range(6)]
        groups = cout[3]
    parity_ok = cin[:3] == cout[:3] and cout[4:] == [0, 0]
    ms_step = 1e3 * dt / a.steps
    dev_ms = {k: v / a.steps for k, v in agg.items()}
    value = world * n / (dt / a.steps)

    # ---- roofline (SURVEY 8d accounting: one read of its input + one write of its output per stage)
    peak, peak_src = peaks()
    R = rb
    info = ctx.result_info()
    g_local = info.groups
    combined = a.workload == "zipf32"  # ctx runs the map-side combiner for this workload
    n2 = info.pairs_recv if combined else n  # N' = pairs that survive the combiner
    # per GPU, SURVEY 8d: combine N R + N' R, hash-partition N' R + N' R, sort N' R + N' R, reduce N' R + U R.
    # The hash-partition stage is one kernel (k_scatter) or, in the default single-GPU layout, the two
    # levels of k_split, each charged half of the stage.
    sort_name = "k_agg_bins" if combined else "k_sort_reduce"
    split = world == 1 and not combined and dev_ms["ms_plan"] > 0.2
    kernels = {}
    if combined:
        kernels["k_combine"] = (n * R + n2 * R, dev_ms["ms_combine"])
    if split:
        kernels["k_split_level1"] = (n2 * R, dev_ms["ms_plan"])
        kernels["k_split_level2"] = (n2 * R, dev_ms["ms_scatter"])
    else:
        kernels["k_scatter"] = (2 * n2 * R, dev_ms["ms_scatter"])
        if dev_ms["ms_hist"] > 0.05:
            kernels["k_hist(overhead)"] = (0, dev_ms["ms_hist"])
    kernels[sort_name] = (3 * n2 * R + g_local * R, dev_ms["ms_sort_reduce"] + dev_ms["ms_bigbins"])
    stage_bytes = {k: v[0] for k, v in kernels.items()}
    k_ms = {k: v[1] for k, v in kernels.items()}
    pipe_bytes = sum(stage_bytes.values())
    dom = max((k for k in k_ms if stage_bytes[k]), key=k_ms.get)
    ach = stage_bytes[dom] / (k_ms[dom] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per launch from the committed ncu capture
    if os.path.exists(tp) and world == 1 and a.workload == "u64" and n == 100_000_000:
        traffic = json.load(open(tp)).get(dom)
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
        "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes": stage_bytes[dom],
        "kernel_ms": k_ms[dom],
        "pipeline": {"algorithmic_bytes": pipe_bytes, "ms": dev_ms["ms_total"],
                     "achieved": pipe_bytes / (dev_ms["ms_total"] * 1e-3) / 1e9,
                     "frac": pipe_bytes / (dev_ms["ms_total"] * 1e-3) / 1e9 / peak},
        "stages_ms": dev_ms,
        # SURVEY 8d "fusion headroom": the strict end-to-end lower bound N R + U R_out of a single-pass hash
        # aggregation over the step's time -- reported beside the graded fraction, not instead of it
        "fusion_headroom": {"lower_bound_bytes": n * R + g_local * R,
                            "frac_of_peak": (n * R + g_local * R) / (dev_ms["ms_total"] * 1e-3) / 1e9 / peak},
        "kernels": {k: {"algorithmic_bytes": stage_bytes[k], "ms": k_ms[k],
                        "achieved": (stage_bytes[k] / (k_ms[k] * 1e-3) / 1e9) if k_ms[k] > 0 else None,
                        "frac": (stage_bytes[k] / (k_ms[k] * 1e-3) / 1e9 / peak) if k_ms[k] > 0 else None}
                    for k in kernels},
    }

    # ---- e2e through the public API from pinned host buffers
    e2e = None
    if a.e2e_steps > 0:
        ctx.reset()
        rec_dt = mrhbm.record_dtype(kind, 27)
        host = ctx.pinned_array(n, rec_dt)
        if a.workload == "u64":
            host_u64_records(synth.SEED, rank * n, n, host)
        else:  # read the device-generated stream back once (outside the timed region)
            mm = ctx.map_begin("g")
            mm.gen_zipf(synth.SEED, rank * n, n, table)
            mm.commit()
            ctx.pool_read(0, n, host)
            ctx.reset()
        # a rank may own slightly more groups than the resident run showed (multi-GPU ownership)
        cap_out = min(n + n // 16, int(g_local) * 2) + 4096
        out_keys = ctx.pinned_array(cap_out, np.uint64 if kind == mrhbm.KEY_U64 else "S%d" % (rb - 4))
        out_sums = ctx.pinned_array(cap_out, np.uint64)
        chunk = 1 << 22

        def one_step(cx, ok_, os_, h2d=None, d2h=None):
            """one task iteration through the public API: emit from pinned host memory, shuffle, read the result"""
            cx.reset()  # every step is a fresh task iteration (server.lua:386-404); frees the pool
            if h2d:
                h2d.acquire()
            try:
                mm = cx.map_begin("e2e")
                for s0 in range(0, n, chunk):
                    c = min(chunk, n - s0)
                    mm.emit_batch_ptr(host.ctypes.data + s0 * rb, c)
                mm.commit()  # returns when the host buffers have been read
            finally:
                if h2d:
                    h2d.release()
            cx.shuffle()
            if d2h:
                d2h.acquire()
            try:
                cx.result_copy(ok_, os_)
            finally:
                if d2h:
                    d2h.release()

        def run_serial():
            dts = []
            for it in range(a.e2e_steps + 1):
                barrier()
                t1 = time.perf_counter()
                one_step(ctx, out_keys, out_sums)
                barrier()
                if it:
                    dts.append(time.perf_counter() - t1)
            return float(np.mean(dts)), ctx.result_info().groups, "one worker: emit -> shuffle -> result_copy, serial"

        def run_duplex():
            # Two worker threads, each with its own ctx and output buffers (the reference runs several workers per
            # host), alternate steps.  One lock per PCIe direction keeps a single H2D and a single D2H in flight, so
            # the upload of step k+1 overlaps the download of step k (full-duplex PCIe) and the device work of both.
            import threading
            ctx2 = mrhbm.Ctx(kind, P, part, max_key_bytes=27, device=local, reserve_pairs=n, combiner=False)
            try:
                outs = [(out_keys, out_sums),
                        (ctx2.pinned_array(cap_out, out_keys.dtype), ctx2.pinned_array(cap_out, np.uint64))]
                ctxs = [ctx, ctx2]
                for w in range(2):  # warm-up: allocations, first-touch of the pinned result buffers
                    one_step(ctxs[w], *outs[w])
                h2d, d2h = threading.Lock(), threading.Lock()
                total_steps = 2 * a.e2e_steps
                errors = []

                def work(w):
                    try:
                        torch.cuda.set_device(local)
                        for _ in range(w, total_steps, 2):
                            one_step(ctxs[w], *outs[w], h2d=h2d, d2h=d2h)
                    except BaseException as e:  # noqa: surfaced below
                        errors.append(e)

                barrier()
                t1 = time.perf_counter()
                th = [threading.Thread(target=work, args=(w,), daemon=True) for w in range(2)]
                for t in th:
                    t.start()
                for t in th:
                    t.join()
                barrier()
                t_step = (time.perf_counter() - t1) / total_steps
                if errors:
                    raise errors[0]
                g = ctx.result_info().groups
                if ctx2.result_info().groups != g:
                    raise RuntimeError("the two workers disagree on the number of groups")
            finally:
                ctx2.close()
            return t_step, g, ("two worker threads x own ctx alternate steps; the H2D of one step overlaps the D2H of "
                               "the previous one (full-duplex PCIe); %d steps timed as one region" % total_steps)

        workers = 2 if (world == 1 and n * rb <= 4_000_000_000 and not a.e2e_serial) else 1
        if workers == 2:
            try:
                e2e_t, g2, mode = run_duplex()
            except Exception as ex:  # keep the bench line: fall back to the serial measurement and say so
                sys.stderr.write("e2e: two-worker run failed (%r); falling back to the serial run\n" % (ex,))
                workers = 1
        if workers == 1:
            e2e_t, g2, mode = run_serial()
        if dist is not None:
            t = torch.tensor([e2e_t], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_t = float(t.item())
        e2e = {"value": world * n / e2e_t, "unit": UNIT, "h2d_bytes_per_step": n * rb,
               "d2h_bytes_per_step": int(g2) * (rb - 4 + 8 if kind == mrhbm.KEY_STR else 16) + 8 * (P + 1),
               "ms_per_step": 1e3 * e2e_t, "steps": a.e2e_steps * workers, "groups_match": bool(g2 == g_local),
               "mode": mode}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline(a, wp, table, min(4, os.cpu_count() or 1), a.cpu_sample or 10_000_000)
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(3, a.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64" if a.workload == "u64" else "u8", "data": "synthetic",
            "config": {"workload": wp["name"], "pairs_per_gpu": n, "partitions": P, "record_bytes": rb,
                       "seed": hex(synth.SEED), "l2": "inputs (%.1f GB) >> 126 MB L2, no flush needed" % (n * rb / 1e9),
                       "bins": st["bins"], "sub_bins": st["sub_bins"], "big_bins": st["big_bins"],
                       "groups": int(groups), "pairs_after_combine": int(n2) if combined else None,
                       "parity_properties_ok": bool(parity_ok)},
            "gpu_launches": int(launches), "device_ms_per_step": dev_ms["ms_total"],
            "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
        }))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
