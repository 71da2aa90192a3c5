"""Worker for tests/test_multirank_gloo.py::test_two_rank_sharded_round_steps: the multi-GPU scheme of
k_rounds_wide on CPU -- ONE hashgraph, every rank holds the whole state, the P_r tests of a round step are sharded
by member chain (chain % nranks == rank) and the first hits of a step are exchanged (here: gloo all_gather_object;
on the GPUs: P2P stores into every peer's buffer from inside the kernel).  Every rank must end with the oracle's rounds."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "py-swirld_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import oracle as orc                       # noqa: E402
from swirld_b200 import traces             # noqa: E402
from test_rounds_model import RoundBatch   # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    tr = traces.gossip(70, 5000, 11)       # the same graph on every rank
    o = orc.Oracle(tr.M)
    o.append(tr)
    o.divide_rounds(0, tr.N)
    evaluated = [0]

    def exchange(mine):
        evaluated[0] += len(mine)
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        merged = {}
        for part in parts:
            merged.update(part)
        return merged

    rb = RoundBatch(tr, o.can_see(), None, 4, frontier=True, rank=rank, nranks=world, exchange=exchange)
    first = 0
    for n in (2000, 3000):
        rb.divide(first, n)
        first += n
    assert np.array_equal(rb.round, o.results()["round"]), "rank %d: rounds differ" % rank
    counts = [None] * world
    dist.all_gather_object(counts, evaluated[0])
    assert all(c > 0 for c in counts) and abs(counts[0] - counts[1]) < 0.2 * sum(counts), counts   # the work is really split
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("SHARD_OK steps=%d chain-steps per rank=%s" % (rb.steps, counts))


if __name__ == "__main__":
    main()
