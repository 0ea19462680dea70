"""Multi-GPU plumbing: one process per GPU (torch.distributed for rendezvous), one mrhbm ctx
per rank joined through mrhbm_comm_init.  Partition p is owned by rank p % world -- the
reference's reduce-job ids are kept (mapreduce/server.lua:316-323), only their placement
changes.  The data path itself (count all-gather + one NCCL all-to-all) lives in the library."""


def partition_owner(p, world):
    return p % world


def owned_partitions(rank, world, num_partitions):
    return list(range(rank, num_partitions, world))


def init_comm(ctx, dist):
    """Joins ctx to the job's communicator: rank 0 creates the NCCL unique id, every rank gets
    it through torch.distributed (any backend), then mrhbm_comm_init."""
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, world)
    return rank, world


def gather_final_pairs(ctx, dist, dst=0):
    """finalfn's view on the server (mapreduce/server.lua:360-385): all (partition, key, values)
    of the whole job in ascending partition id, then ascending key.  Returned on rank dst."""
    mine = [(p, k, v) for p in ctx.partitions() for k, v in ctx.groups(p)]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(mine, out, dst=dst)
    if out is None:
        return None
    merged = [t for part in out for t in part]
    merged.sort(key=lambda t: t[0])  # stable: keys stay ascending inside a partition
    return merged


def max_over_ranks(x, dist):
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
