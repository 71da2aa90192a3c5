"""GpuNode over the real engine (GPU): the gossip simulation with real signatures;
each node's device state equals the oracle's replay of that node's own trace and call
schedule (K = one sync per call: the streaming cadence of Node.main)."""
import pytest

import node_sim
from util import assert_same

pytestmark = pytest.mark.gpu
KEYS = ["round", "famous", "consensus", "transactions"]


def test_simulation_over_the_gpu_engine():
    sim = node_sim.run_sim(5, 400, capacity=128)      # small capacity: forces one growth replay
    txs = [n.transactions for n in sim]
    k = min(len(t) for t in txs)
    assert k > 100 and all(t[:k] == txs[0][:k] for t in txs)
    for nd in sim:
        tr, sizes = node_sim.node_trace(nd)
        assert_same(node_sim.replay_oracle(tr, sizes), node_sim.node_results(nd), KEYS, "GpuNode vs oracle replay")
        h = nd.head
        assert nd.can_see[h][nd.pk] == h and nd.round[h] >= 0
