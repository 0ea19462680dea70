"""world_size-2 gloo run of the N>1 host logic (no GPU): communicator hand-shake through
torch.distributed, partition ownership, finalfn-order gather, max-over-ranks timing.  The ctx is
a stand-in that records what the real mrhbm.Ctx would receive."""
import os
import sys

import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mrhbm_loader  # noqa: E402

mrhbm_loader.load()
from lua_mapreduce_b200 import parallel  # noqa: E402


class StandIn:
    def __init__(self, rank, world, P):
        self.rank, self.world, self.P = rank, world, P
        self.uid = None

    def comm_unique_id(self):
        return bytes(range(128))

    def comm_init(self, uid, rank, world):
        self.uid, self.rank_seen, self.world_seen = uid, rank, world

    def partitions(self):
        return parallel.owned_partitions(self.rank, self.world, self.P)

    def groups(self, p):
        for i in range(3):
            yield b"k%d_%d" % (p, i), [p * 10 + i]


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = StandIn(rank, world, 7)
    assert parallel.init_comm(ctx, dist) == (rank, world)
    assert ctx.uid == bytes(range(128)) and ctx.rank_seen == rank and ctx.world_seen == world
    assert all(parallel.partition_owner(p, world) == rank for p in ctx.partitions())
    pairs = parallel.gather_final_pairs(ctx, dist)
    if rank == 0:
        assert [p for p, _, _ in pairs] == sorted(p for p in range(7) for _ in range(3))
        assert pairs[0] == (0, b"k0_0", [0]) and pairs[-1] == (6, b"k6_2", [62])
    else:
        assert pairs is None
    assert parallel.max_over_ranks(1.0 + rank, dist) == float(world)
    dist.barrier()
    if rank == 0:
        print("GLOO_CHECK_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
