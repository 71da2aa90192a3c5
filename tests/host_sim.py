"""A small gossip host of our own for the tests that run where the reference is not available (the GPU box):
what `swirld_b200.node.bind()` needs from a host class -- `new_event`, `is_valid_event`, `sync`, `ask_sync`,
`main` -- with a simpler, fork-free pull protocol than the reference's (the asker names the head it holds of
every member, the peer answers with the chain suffixes beyond those heads in its own arrival order).
Where the reference is mounted the same tests also run over `bind(swirld.Node)` itself."""
import contextlib
import io
import pickle
import random
import time
from collections import namedtuple

import sodium

Event = namedtuple("Event", "d p t c s")      # the fields the engine boundary reads: parents, time, creator, signature


def _event_id(ev):
    return sodium.crypto_generichash(pickle.dumps(ev))


class HostNode:
    """Mixed in BEHIND swirld_b200.node.GpuConsensus (which owns hg, head, _heads, add_event and the hot path)."""

    def new_event(self, d, p):
        stamp = time.time()
        sig = sodium.crypto_sign_detached(pickle.dumps((d, p, stamp, self.pk)), self.sk)
        ev = Event(d, p, stamp, self.pk, sig)
        return _event_id(ev), ev

    def is_valid_event(self, h, ev):
        try:
            sodium.crypto_sign_verify_detached(ev.s, pickle.dumps(tuple(ev[:4])), ev.c)
        except ValueError:
            return False
        if _event_id(ev) != h:
            return False
        if not ev.p:
            return True
        if len(ev.p) != 2 or any(q not in self.hg for q in ev.p):
            return False
        return self.hg[ev.p[0]].c == ev.c and self.hg[ev.p[1]].c != ev.c

    def ask_sync(self, asker, request):
        """Answer with my head and, per member, my events beyond the head the asker holds."""
        theirs = pickle.loads(sodium.crypto_sign_open(request, asker))
        missing = []
        for member, mine in self._heads.items():
            stop, cursor = theirs.get(member), mine
            while cursor is not None and cursor != stop:
                missing.append(cursor)
                parents = self.hg[cursor].p
                cursor = parents[0] if parents else None
        missing.sort(key=self._h2i.__getitem__)          # my arrival order is a topological order
        return sodium.crypto_sign(pickle.dumps((self.head, [(h, self.hg[h]) for h in missing])), self.sk)

    def sync(self, peer, payload):
        request = sodium.crypto_sign(pickle.dumps(dict(self._heads)), self.sk)
        their_head, items = pickle.loads(sodium.crypto_sign_open(self.network[peer](self.pk, request), peer))
        fresh = []
        for h, ev in items:
            if h not in self.hg and self.is_valid_event(h, ev):
                self.add_event(h, ev)
                fresh.append(h)
        if their_head in self.hg:
            h, ev = self.new_event(payload, (self.head, their_head))
            self.add_event(h, ev)
            self.head = h
            fresh.append(h)
        return tuple(fresh)

    def main(self):
        """One gossip step per send(): sync with a random peer, then the three hot-path calls (the schedule of
        swirld.py:319-328: one (divide_rounds, decide_fame, find_order) triple per sync)."""
        fresh = ()
        while True:
            payload = yield fresh
            peer = random.choice([pk for pk in self.network if pk != self.pk])
            fresh = self.sync(peer, payload)
            self.divide_rounds(fresh)
            self.find_order(self.decide_fame())


def run_sim(n_nodes, n_turns, node_cls, seed=None, **kw):
    """n_nodes nodes sharing one `network` dict of bound ask_sync methods, n_turns random gossip steps."""
    rng = random.Random(seed)
    keys = [sodium.crypto_sign_keypair() for _ in range(n_nodes)]
    network, stake = {}, {kp[0]: 1 for kp in keys}
    with contextlib.redirect_stdout(io.StringIO()):
        nodes = [node_cls(kp, network, n_nodes, stake, **kw) for kp in keys]
        for nd in nodes:
            network[nd.pk] = nd.ask_sync
        loops = [nd.main() for nd in nodes]
        for lp in loops:
            next(lp)
        for _ in range(n_turns):
            next(loops[rng.randrange(n_nodes)])
    return nodes
