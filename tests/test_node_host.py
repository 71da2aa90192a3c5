"""GpuNode's host logic on CPU: the reference's gossip simulation (real Ed25519
signatures, BLAKE2b ids, random peers) runs over GpuNode with the oracle standing in
for the device; every node's state must equal a replay of its own arrival trace and
call schedule through the oracle and -- where /root/reference is mounted -- through
the unmodified reference."""
import pytest

import node_sim
import ref_harness as rh
from oracle_engine import OracleEngine
from util import assert_same

KEYS = ["round", "famous", "consensus", "transactions"]


def _factory(M, cap, stake, C):
    return OracleEngine(M, cap, stake, C)


@pytest.fixture(scope="module")
def sim():
    return node_sim.run_sim(4, 400, engine_factory=_factory, capacity=64)   # small capacity: forces growth replay


def test_transactions_are_consistent(sim):
    # prefix agreement between nodes is NOT a property of the reference (its final order
    # depends on each node's call schedule); each node's own state is
    # pinned against the oracle / the reference below.
    for nd in sim:
        assert len(nd.transactions) > 20
        assert len(set(nd.transactions)) == len(nd.transactions)


def test_views_and_attributes(sim):
    nd = sim[0]
    h = nd.head
    assert nd.hg[h].c == nd.pk and nd.height[h] >= 1
    assert nd.can_see[h][nd.pk] == h
    assert nd.round[h] >= nd.round[nd.hg[h].p[0]]
    assert set(nd.idx) == set(nd.transactions)
    assert [nd.idx[x] for x in nd.transactions] == list(range(len(nd.transactions)))
    assert nd.tbd == set(nd.hg) - set(nd.transactions)
    r0 = nd.witnesses[0]
    assert set(r0) <= set(nd.stake) and all(nd.round[w] == 0 for w in r0.values())
    fam = [x for x in nd.hg if x in nd.famous]
    assert fam and all(isinstance(nd.famous[x], bool) for x in fam)
    assert nd.famous.get(b"nope") is None
    with pytest.raises(KeyError):
        nd.round[b"nope"]


def test_each_node_matches_oracle_replay(sim):
    for nd in sim:
        tr, sizes = node_sim.node_trace(nd)
        assert_same(node_sim.replay_oracle(tr, sizes), node_sim.node_results(nd), KEYS, "node vs oracle replay")


@pytest.mark.skipif(not rh.reference_available(), reason="reference not mounted (GPU box)")
def test_each_node_matches_reference_replay(sim):
    for nd in sim[:2]:
        tr, sizes = node_sim.node_trace(nd)
        assert_same(node_sim.replay_reference(tr, sizes), node_sim.node_results(nd), KEYS, "node vs reference replay")


@pytest.mark.skipif(not rh.reference_available(), reason="reference not mounted (GPU box)")
def test_drop_in_for_the_reference_drivers():
    """`swirld.Node = GpuNode` before `swirld.test(...)`: the reference's own driver and
    main loop run unchanged over the replacement class."""
    import contextlib
    import functools
    import io
    swirld = rh.load_reference()
    from swirld_b200 import node as gnode
    saved = swirld.Node
    try:
        swirld.Node = functools.partial(gnode.GpuNode, engine_factory=_factory)
        import random
        random.seed(20260922)           # the driver gossips with the global RNG (the key pairs stay random)
        with contextlib.redirect_stdout(io.StringIO()):
            nodes = swirld.test(4, 300)
    finally:
        swirld.Node = saved
    assert min(len(n.transactions) for n in nodes) > 5
    for nd in nodes:       # each node equals the reference's replay of its own trace + schedule
        tr, sizes = node_sim.node_trace(nd)
        assert_same(node_sim.replay_reference(tr, sizes), node_sim.node_results(nd), KEYS, "drop-in node vs reference")
