"""worker of tests/test_multiprocess.py: the N > 1 host logic of bench.py on the gloo backend (CPU)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # every rank encodes its own independent sequence: distinct seeds, no data-path collective
    seed = bench.rank_seed(1000, rank)
    seeds = [None] * world
    dist.all_gather_object(seeds, seed)
    # the step time reported is the max over ranks
    mine = 10.0 + rank
    mx = bench.max_over_ranks(mine, world, device="cpu")
    agg = bench.aggregate_fps(mx, world)
    if rank == 0:
        print(json.dumps({"seeds": seeds, "max_ms": mx, "fps": agg, "world": world}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
