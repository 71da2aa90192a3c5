"""Worker for tests/test_multirank_gloo.py: world_size-2 gloo rendezvous on CPU, the
bench's multi-rank plumbing (independent node-view per rank, barrier, max over ranks,
whole-job rate).  No engine calls (no GPU here)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "py-swirld_b200"))
import bench  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert world == 2
    # each rank builds its own (different) trace
    tr = bench.make_trace(dict(M=8, N=512, K=64), bench.rank_seed(rank))
    digest = torch.tensor([int(tr.p1.astype("int64").sum())], dtype=torch.int64)
    both = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(both, digest)
    assert both[0].item() != both[1].item(), "ranks must run independent node-views"
    dist.barrier()
    my_ms = [10.0 + 5 * rank, 20.0 - 3 * rank, 7.0]
    mx = bench.max_over_ranks(my_ms, dist, "cpu")
    assert mx == [15.0, 20.0, 7.0], mx
    rate = bench.whole_job_rate(world, 512, 3, mx[0])
    assert abs(rate - 2 * 512 * 3 / 0.015) < 1e-6
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("GLOO_OK")


if __name__ == "__main__":
    main()
