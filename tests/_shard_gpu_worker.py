"""Worker for tests/test_gpu_multi.py: one hashgraph on WORLD_SIZE GPUs (sw_peer_connect).  Every rank feeds the same
trace through the same calls; the P_r tests of every round step are sharded by member chain and exchanged over NVLink
inside k_rounds_wide.  Every rank must end with the oracle's results (rank 0 checks against the oracle, all ranks
against each other)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "py-swirld_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import bench                                # noqa: E402
import oracle as orc                        # noqa: E402
from swirld_b200 import engine, traces      # noqa: E402
from swirld_b200.traces import chunks       # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [("gossip", 96, 20000, 3000), ("adversarial", 130, 30000, 8192), ("gossip_np", 256, 120000, 40000), ("gossip_np", 1024, 24000, 12000)]
    for gen, M, N, K in cases:
        tr = getattr(traces, gen)(M, N, 7)
        e = engine.Engine(M, N, device=local)
        bench.connect_peers(e, dist, rank, world)
        ncs = []
        for first, cnt in chunks(N, K):
            e.append_trace(tr, first, cnt)
            dist.barrier()
            e.divide_rounds(first, cnt)
            nc = e.decide_fame()
            e.find_order(nc)
            ncs.append(sorted(nc))
        r = e.results()
        digest = torch.tensor([int(r["round"].astype(np.int64).sum()), int(r["famous"].astype(np.int64).sum()),
                               int(r["transactions"].astype(np.int64).sum() % (1 << 40)), len(r["transactions"])], dtype=torch.int64, device="cuda")
        lo, hi = digest.clone(), digest.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks disagree on %s M=%d: %s vs %s" % (gen, M, lo.tolist(), hi.tolist())
        if rank == 0:
            o = orc.run_oracle(tr, K)
            for k in ("round", "witness_table", "famous", "consensus", "transactions"):
                assert np.array_equal(np.asarray(o[k]), np.asarray(r[k])), "%s differs from the oracle (%s M=%d)" % (k, gen, M)
            assert o["new_c_per_call"] == ncs
            print("SHARD_GPU_OK %s M=%d N=%d world=%d max_round=%d ordered=%d" % (gen, M, N, world, int(r["round"].max()), len(r["transactions"])), flush=True)
        e.close()
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
