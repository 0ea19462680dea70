"""Launched by torchrun (one rank per GPU) from tests/test_gpu_multi.py and profiles/*.sh:
the sharded shuffle (level-1 split into rank-aligned regions, level 2 of the owner pulling its regions from every
peer over NVLink, per-rank sort/reduce; the exact-layout NCCL path for skewed keys) must give the oracle's result
for the union of all ranks' pairs, with partition p living on rank p % world."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mrhbm_loader  # noqa: E402

mrhbm_loader.load()
from lua_mapreduce_b200 import mrhbm, parallel, synth  # noqa: E402


def check(name, cond):
    if not cond:
        raise AssertionError("rank %d: %s" % (dist.get_rank(), name))


def run_u64(rank, world, n_per, P, flags=0, dup=False):
    import oracle as O
    with mrhbm.Ctx(mrhbm.KEY_U64, P, device=int(os.environ["LOCAL_RANK"]), flags=flags) as ctx:
        parallel.init_comm(ctx, dist)
        m = ctx.map_begin("r%d" % rank)
        if dup:  # hot key + few distinct keys: big-bin path after the exchange
            keys, vals = O.gen_u64(synth.SEED, rank * n_per, n_per)
            keys = keys[np.arange(n_per) % 3000]
            keys[::5] = np.uint64(42)
            recs = np.zeros(n_per, dtype=mrhbm.record_dtype(mrhbm.KEY_U64))
            recs["key"], recs["val"] = keys, vals
            m.emit_batch(recs)
        else:
            m.gen_u64(synth.SEED, rank * n_per, n_per)
        m.commit()
        ctx.shuffle()
        info = ctx.result_info()
        gk, gs, gpo = ctx.result_copy()
        cin, cout = ctx.checksum_input(), ctx.checksum_result()
        st = ctx.stats()
        pairs = parallel.gather_final_pairs(ctx, dist) if n_per <= 200_000 else None
        box = [None] * world
        payload = (gk, gs, gpo, cin, cout, info.sorted, ctx.partitions(), st,
                   (keys, vals) if dup else None)
        dist.gather_object(payload, box if rank == 0 else None, dst=0)
        if rank != 0:
            return
        if dup:
            keys = np.concatenate([b[8][0] for b in box])
            vals = np.concatenate([b[8][1] for b in box])
        else:
            keys, vals = O.gen_u64(synth.SEED, 0, n_per * world)
        ok, osum, po = O.groupby_u64(keys, vals, O.PART_MULHASH, P)
        tot_in = [sum(b[3][i] for b in box) % 2**64 for i in range(3)]
        tot_out = [sum(b[4][i] for b in box) % 2**64 for i in range(3)]
        check("sum linearity across ranks", tot_in == tot_out)
        check("pairs", sum(b[3][3] for b in box) == keys.size)
        check("groups", sum(b[4][3] for b in box) == ok.size)
        check("order / membership", all(b[4][4:] == [0, 0] for b in box))
        for r, (gk, gs, gpo, _, _, srt, parts, st, _) in enumerate(box):
            check("ownership", all(p % world == r for p in parts))
            for p in range(P):
                a, b = int(gpo[p]), int(gpo[p + 1])
                if p % world != r:
                    check("foreign partition empty", a == b)
                    continue
                wa, wb = int(po[p]), int(po[p + 1])
                check("partition size", b - a == wb - wa)
                k, s = gk[a:b], gs[a:b]
                if not srt:
                    o = np.argsort(k, kind="stable")
                    k, s = k[o], s[o]
                check("partition %d keys" % p, (k == ok[wa:wb]).all() and (s == osum[wa:wb]).all())
        if pairs is not None:
            want = [(int(np.searchsorted(po, i, side="right")) - 1, int(ok[i]).to_bytes(8, "big"), [int(osum[i])])
                    for i in range(ok.size)]
            check("finalfn order over all ranks", pairs == want)
        print("u64 world=%d n/rank=%d P=%d dup=%s ok: groups=%d sorted=%s big_bins=%s attempts=%s bins=%d pulled %.1f MB"
              % (world, n_per, P, dup, ok.size, [b[5] for b in box], [b[7]["big_bins"] for b in box],
                 [b[7]["attempts"] for b in box], box[0][7]["bins"], box[0][7]["bytes_exchanged"] / 1e6), flush=True)


def run_zipf(rank, world, n_per, P, combiner=False):
    import oracle as O
    table = synth.zipf_table(1 << 14)
    with mrhbm.Ctx(mrhbm.KEY_STR, P, mrhbm.PART_FNV_LUA, device=int(os.environ["LOCAL_RANK"]), combiner=combiner) as ctx:
        parallel.init_comm(ctx, dist)
        m = ctx.map_begin("r%d" % rank)
        m.gen_zipf(synth.SEED, rank * n_per, n_per, table)
        m.commit()
        ctx.shuffle()
        pairs = parallel.gather_final_pairs(ctx, dist)
        st = ctx.stats()
        if rank != 0:
            return
        recs = O.gen_zipf_rec32(synth.SEED, 0, n_per * world, table)
        okeys, osum, po = O.groupby_rec(recs, O.PART_FNV_LUA, P)
        want = [(int(np.searchsorted(po, i, side="right")) - 1, bytes(okeys[i]).rstrip(b"\0"), [int(osum[i])])
                for i in range(osum.size)]
        check("zipf word count over all ranks", pairs == want)
        print("zipf world=%d n/rank=%d P=%d combiner=%s ok: groups=%d big_bins(rank0)=%d attempts=%d"
              % (world, n_per, P, combiner, osum.size, st["big_bins"], st["attempts"]), flush=True)


def run_group_only(rank, world, P):
    """MRHBM_RED_NONE on several GPUs (job.lua:275-284 hands reducefn EVERY value of a key): one key carries more
    values than a shared-memory bin holds and they come from all ranks -- the owner makes the bin contiguous,
    sorts it run by run, and the iterator merges the runs."""
    hot, nhot, nother = 7, 5000, 3000
    with mrhbm.Ctx(mrhbm.KEY_U64, P, device=int(os.environ["LOCAL_RANK"]), reducer=mrhbm.RED_NONE) as ctx:
        parallel.init_comm(ctx, dist)
        keys = np.concatenate([np.full(nhot, hot, dtype=np.uint64),
                               np.arange(1000 + rank * nother, 1000 + (rank + 1) * nother, dtype=np.uint64) // 3])
        vals = (np.arange(keys.size, dtype=np.uint64) + rank * keys.size)
        rec = np.zeros(keys.size, dtype=mrhbm.record_dtype(mrhbm.KEY_U64))
        rec["key"], rec["val"] = keys, vals
        m = ctx.map_begin("r%d" % rank)
        m.emit_batch(rec)
        m.commit()
        ctx.shuffle()
        pairs = parallel.gather_final_pairs(ctx, dist)
        st = ctx.stats()
        allk = [None] * world if rank == 0 else None
        dist.gather_object((keys.tolist(), vals.tolist()), allk, dst=0)
        if rank != 0:
            return
        want = {}
        for ks, vs in allk:
            for k, v in zip(ks, vs):
                want.setdefault(k, []).append(v)
        got = {int.from_bytes(k, "big"): sorted(v) for _, k, v in pairs}
        check("group-only: every value of every key, once", got == {k: sorted(v) for k, v in want.items()})
        check("group-only: the hot key has all its values", len(got[hot]) == nhot * world)
        per_part = {}
        for p_, k, _ in pairs:
            per_part.setdefault(p_, []).append(k)
        check("group-only: keys ascending and distinct inside a partition",
              all(v == sorted(v) and len(set(v)) == len(v) for v in per_part.values()))
        print("group-only world=%d P=%d ok: keys=%d hot key values=%d big_bins(rank0)=%d"
              % (world, P, len(got), len(got[hot]), st["big_bins"]), flush=True)


def run_long_keys(rank, world, P):
    """Keys that do not fit a record slot (host side store): every rank emits them, the barrier all-gathers them and
    the owner of a key's partition groups them; the job's result is the oracle engine's."""
    import oracle as O
    rng = np.random.default_rng(100 + rank)
    stem = b"abcdefghijklmnopqrstuvwxyz0"
    longs = [stem + b"X", stem + b"Y" * 300, b"q" * 4000, stem + b"\x00\x01z", b"w" * 28]
    pairs = [(O.rank_to_key(int(r)), int(v)) for r, v in zip(rng.integers(1, 500, 3000), rng.integers(1, 100, 3000))]
    pairs += [(k, rank + 1 + i) for i, k in enumerate(longs)] * 2
    pairs += [(b"only-rank-%d-" % rank + b"L" * 100, 5)]
    with mrhbm.Ctx(mrhbm.KEY_STR, P, mrhbm.PART_FNV_LUA, max_key_bytes=27, device=int(os.environ["LOCAL_RANK"])) as ctx:
        parallel.init_comm(ctx, dist)
        m = ctx.map_begin("r%d" % rank)
        for k, v in pairs:
            m.emit(k, v)
        m.commit()
        ctx.shuffle()
        got = parallel.gather_final_pairs(ctx, dist)
        allp = [None] * world if rank == 0 else None
        dist.gather_object(pairs, allp, dst=0)
        if rank != 0:
            return
        e = O.Engine(O.PART_FNV_LUA, P, combiner=-1, reducer=O.RED_SUM, aci=True)
        for j, pp in enumerate(allp):
            e.map_job(j, pairs=pp)
        e.reduce_all()
        want = [(p_, k, [int(x) for x in v]) for p_, k, v in e.final_pairs()]
        check("long keys over all ranks", got == want)
        print("long keys world=%d P=%d ok: groups=%d of them longer than a slot=%d"
              % (world, P, len(want), sum(len(k) > 27 for _, k, _ in want)), flush=True)


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    rank, world = dist.get_rank(), dist.get_world_size()
    run_u64(rank, world, 100_000, 16)
    run_u64(rank, world, 1_000_000, 1024)
    run_u64(rank, world, 6_000_000, 1024)                # several tiles per (region, source): the level-2 pulls in earnest
    run_u64(rank, world, 200_000, 7, dup=True)           # P not a multiple of world, hot key
    run_u64(rank, world, 150_000, 3, flags=mrhbm.F_FORCE_RUNS)
    run_zipf(rank, world, 200_000, 15)
    run_zipf(rank, world, 1_500_000, 15, combiner=True)  # local combine (global table) on every rank, then the exchange
    run_group_only(rank, world, 5)
    run_long_keys(rank, world, 15)
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK_OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
