"""GpuNode -- the reference's `Node` method surface over the B200 engine.

Drop-in for `swirld.Node` (/root/reference/swirld.py:37-328): same constructor
`(kp, network, n_nodes, stake)`, same methods (`main`, `sync`, `ask_sync`,
`new_event`, `is_valid_event`, `add_event`, `divide_rounds`, `decide_fame`,
`find_order`) and the same attributes that `viz.py` and the drivers read
(`hg`, `head`, `height`, `round`, `famous`, `idx`, `transactions`, `consensus`,
`witnesses`, `can_see`, `tbd`, `n`, `pk`, `network`, `stake`):

    import swirld, swirld_b200.node
    swirld.Node = swirld_b200.node.GpuNode      # before swirld.test(...) / importing viz

Gossip, crypto and the wire format stay host Python (out of scope, SURVEY.md
section 2); the consensus state lives on the GPU.  Hashes (bytes) are mapped to
arrival indices and public keys to member ids at this boundary; the attributes are
lazy mapping views that pull from the device only what is asked for.

No CPU fallback: constructing a GpuNode without the CUDA library / a GPU raises.
(`engine_factory` exists so the CPU test-suite can exercise this host logic against
the oracle; the product never passes it.)
"""
from __future__ import annotations

from collections import deque, namedtuple
from collections.abc import Mapping
from pickle import dumps, loads
from time import time

import numpy as np

from . import crypto

C = 6                                     # coin period (swirld.py:17)
Event = namedtuple("Event", "d p t c s")  # same fields as swirld.py:30


def _toposort(nodes, parents):
    """Parents-first order of `nodes` (iterative DFS); ValueError on a cycle
    (utils.py:8-21 semantics)."""
    state, out = {}, []
    for root in nodes:
        if root in state:
            continue
        stack = [(root, iter(parents(root)))]
        state[root] = 0
        while stack:
            u, it = stack[-1]
            for v in it:
                if v not in nodes:
                    continue
                s = state.get(v)
                if s == 0:
                    raise ValueError("not a DAG")
                if s is None:
                    state[v] = 0
                    stack.append((v, iter(parents(v))))
                    break
            else:
                state[u] = 1
                out.append(u)
                stack.pop()
    return out


def _bfs(starts, succ):
    starts = tuple(starts)
    seen = set(starts)
    q = deque(starts)
    while q:
        u = q.popleft()
        yield u
        for v in succ(u):
            if v not in seen:
                seen.add(v)
                q.append(v)


class _View(Mapping):
    """Read-only dict-like view keyed by event hash."""

    def __init__(self, node, getter, present):
        self._n, self._get, self._present = node, getter, present

    def __getitem__(self, h):
        i = self._n._h2i[h]
        if not self._present(i):
            raise KeyError(h)
        return self._get(i)

    def __contains__(self, h):
        i = self._n._h2i.get(h)
        return i is not None and self._present(i)

    def __iter__(self):
        return (h for i, h in enumerate(self._n._i2h) if self._present(i))

    def __len__(self):
        return sum(1 for _ in self)


class _Witnesses(Mapping):
    """witnesses[r] -> {member pk: event hash}  (swirld.py:61)."""

    def __init__(self, node):
        self._n = node

    def __getitem__(self, r):
        n = self._n
        if r < 0 or r > n._eng.max_round:
            raise KeyError(r)
        row = n._eng.witness_table(r, 1)[0]
        return {n._m2pk[c]: n._i2h[int(w)] for c, w in enumerate(row) if w >= 0}

    def __iter__(self):
        return iter(range(self._n._eng.max_round + 1))

    def __len__(self):
        return self._n._eng.max_round + 1


class GpuNode:
    def __init__(self, kp, network, n_nodes, stake, capacity=1 << 15, device=0, engine_factory=None):
        self.pk, self.sk = kp
        self.network = network            # {pk -> Node.ask_sync}
        self.n = n_nodes
        self.stake = stake
        self.tot_stake = sum(stake.values())
        self.min_s = 2 * self.tot_stake / 3
        # member ids: the order of the stake dict (every node is built with the same one)
        self._m2pk = list(stake.keys())
        self._pk2m = {pk: m for m, pk in enumerate(self._m2pk)}
        self._stake_list = [int(stake[pk]) for pk in self._m2pk]
        self._device = device
        self._factory = engine_factory
        self._capacity = int(capacity)
        self._eng = self._make_engine(self._capacity)
        # host side of the graph (the wire format needs the Event objects)
        self.hg = {}
        self.height = {}
        self._h2i, self._i2h = {}, []
        self._pending = []                # appended to hg, not yet on the device
        self._n_on_device = 0
        self._ops = []                    # call schedule, replayed if the engine has to grow
        self.head = None
        self.transactions = []
        self.consensus = set()
        self.votes = {}                   # swirld.py:59; internal to decide_fame, kept for shape only
        self._round_cache = np.empty(0, np.int32)
        self._famous_cache = None
        self._idx_cache = None
        self.round = _View(self, self._round_of, lambda i: i < self._eng.n_divided)
        self.famous = _View(self, lambda i: bool(self._famous()[i]), lambda i: i < self._eng.n_events and self._famous()[i] >= 0)
        self.idx = _View(self, lambda i: int(self._idx()[i]), lambda i: i < self._eng.n_events and self._idx()[i] >= 0)
        self.can_see = _View(self, self._can_see_of, lambda i: i < self._eng.n_divided)
        self.witnesses = _Witnesses(self)
        # first local event (swirld.py:75-80)
        h, ev = self.new_event(None, ())
        self.add_event(h, ev)
        self.divide_rounds((h,))
        self.head = h

    # ------------------------------------------------------------------ engine plumbing
    def _make_engine(self, capacity):
        if self._factory is not None:
            return self._factory(len(self._m2pk), capacity, self._stake_list, C)
        from .engine import Engine
        return Engine(len(self._m2pk), capacity, self._stake_list, C, self._device)

    def _columns(self, idxs):
        n = len(idxs)
        p0 = np.empty(n, np.int32); p1 = np.empty(n, np.int32); cr = np.empty(n, np.int32)
        t = np.empty(n, np.float64); sig = np.empty((n, 64), np.uint8)
        for j, i in enumerate(idxs):
            ev = self.hg[self._i2h[i]]
            if ev.p == ():
                p0[j] = p1[j] = -1
            else:
                p0[j], p1[j] = self._h2i[ev.p[0]], self._h2i[ev.p[1]]
            cr[j] = self._pk2m[ev.c]
            t[j] = ev.t
            sig[j] = np.frombuffer(ev.s, np.uint8)
        return p0, p1, cr, t, sig

    def _flush(self):
        """Move events added since the last call to the device (one sw_append)."""
        if not self._pending:
            return
        need = self._n_on_device + len(self._pending)
        if need > self._capacity:
            self._grow(need)
        self._eng.append(*self._columns(self._pending))
        self._n_on_device = need
        self._pending = []

    def _grow(self, need):
        """Double the engine and replay the recorded call schedule (the final order
        depends on the schedule, so it is replayed call by call)."""
        while self._capacity < need:
            self._capacity *= 2
        old = self._eng
        self._eng = self._make_engine(self._capacity)
        if self._n_on_device:
            self._eng.append(*self._columns(list(range(self._n_on_device))))
        for op in self._ops:
            if op[0] == "d":
                self._eng.divide_rounds(op[1], op[2])
            elif op[0] == "f":
                self._eng.decide_fame()
            else:
                self._eng.find_order(list(op[1]))
        old.close()

    def _round_of(self, i):
        nd = self._eng.n_divided
        if self._round_cache.shape[0] < nd:        # rounds never change once assigned
            have = self._round_cache.shape[0]
            self._round_cache = np.concatenate([self._round_cache, self._eng.rounds(have, nd - have)])
        return int(self._round_cache[i])

    def _famous(self):
        if self._famous_cache is None:
            self._famous_cache = self._eng.famous()
        return self._famous_cache

    def _idx(self):
        if self._idx_cache is None:
            self._idx_cache = self._eng.idx()
        return self._idx_cache

    def _can_see_of(self, i):
        row = self._eng.can_see(i, 1)[0]
        return {self._m2pk[c]: self._i2h[int(k)] for c, k in enumerate(row) if k >= 0}

    @property
    def tbd(self):
        """Events whose final order is still to be determined (swirld.py:52)."""
        idx = self._idx()
        return {h for i, h in enumerate(self._i2h) if i >= idx.shape[0] or idx[i] < 0}

    # ------------------------------------------------------------------ events (host, out of scope for the GPU)
    def new_event(self, d, p):
        """Create a new event and its id (swirld.py:82-95)."""
        assert p == () or len(p) == 2
        assert p == () or self.hg[p[0]].c == self.pk
        assert p == () or self.hg[p[1]].c != self.pk
        t = time()
        s = crypto.crypto_sign_detached(dumps((d, p, t, self.pk)), self.sk)
        ev = Event(d, p, t, self.pk, s)
        return crypto.crypto_generichash(dumps(ev)), ev

    def is_valid_event(self, h, ev):
        try:
            crypto.crypto_sign_verify_detached(ev.s, dumps(ev[:-1]), ev.c)
        except ValueError:
            return False
        if crypto.crypto_generichash(dumps(ev)) != h:
            return False
        if ev.p == ():
            return True
        return (len(ev.p) == 2 and ev.p[0] in self.hg and ev.p[1] in self.hg
                and self.hg[ev.p[0]].c == ev.c and self.hg[ev.p[1]].c != ev.c)

    def add_event(self, h, ev):
        """hg[h] = ev, height (swirld.py:114-120); the device copy is batched into
        the next divide_rounds."""
        self.hg[h] = ev
        self.height[h] = 0 if ev.p == () else max(self.height[p] for p in ev.p) + 1
        self._h2i[h] = len(self._i2h)
        self._i2h.append(h)
        self._pending.append(self._h2i[h])
        self._idx_cache = None
        self._famous_cache = None

    def sync(self, pk, payload):
        """Pull-gossip with member pk; returns the new event ids in topological
        order (swirld.py:122-146)."""
        info = crypto.crypto_sign(dumps({c: self.height[h] for c, h in self.can_see[self.head].items()}), self.sk)
        msg = crypto.crypto_sign_open(self.network[pk](self.pk, info), pk)
        remote_head, remote_hg = loads(msg)
        unknown = remote_hg.keys() - self.hg.keys()
        new = tuple(_toposort(unknown, lambda u: remote_hg[u].p))
        for h in new:
            ev = remote_hg[h]
            if self.is_valid_event(h, ev):
                self.add_event(h, ev)
        if self.is_valid_event(remote_head, remote_hg[remote_head]):
            h, ev = self.new_event(payload, (self.head, remote_head))
            assert self.is_valid_event(h, ev)
            self.add_event(h, ev)
            self.head = h
        return new + (h,)

    def ask_sync(self, pk, info):
        """Answer a sync request (swirld.py:148-161)."""
        cs = loads(crypto.crypto_sign_open(info, pk))

        def unknown_parents(u):
            for p in self.hg[u].p:
                c = self.hg[p].c
                if c not in cs or self.height[p] > cs[c]:
                    yield p
        subset = {h: self.hg[h] for h in _bfs((self.head,), unknown_parents)}
        return crypto.crypto_sign(dumps((self.head, subset)), self.sk)

    # ------------------------------------------------------------------ the hot path (GPU)
    def divide_rounds(self, events):
        """can_see / round / witnesses for the topologically sorted new events
        (swirld.py:187-222) -- one sw_divide_rounds."""
        events = tuple(events)
        if not events:
            return
        self._flush()
        first = self._eng.n_divided
        for j, h in enumerate(events):
            if self._h2i[h] != first + j:        # KeyError for unknown ids, like swirld.py:194
                raise ValueError("divide_rounds: events must be the new events in arrival order")
        self._eng.divide_rounds(first, len(events))
        self._ops.append(("d", first, len(events)))

    def decide_fame(self):
        """Virtual voting (swirld.py:224-277); returns the set of new consensus rounds."""
        self._flush()
        new_c = set(self._eng.decide_fame())
        self._ops.append(("f",))
        self.consensus |= new_c
        self._famous_cache = None
        return new_c

    def find_order(self, new_c):
        """Consensus order of the events received in the new rounds (swirld.py:280-311)."""
        new_c = sorted(new_c)
        if new_c:
            added = self._eng.find_order(new_c)
            self._ops.append(("o", tuple(new_c)))
            if added:
                have = len(self.transactions)
                self.transactions += [self._i2h[int(i)] for i in self._eng.transactions(have, added)]
                self._idx_cache = None
        if self.consensus:
            print(self.consensus)                  # swirld.py:310-311

    def main(self):
        """Main working loop (swirld.py:315-328): a coroutine, one gossip step per send()."""
        new = ()
        while True:
            payload = (yield new)
            peers = tuple(self.network.keys() - {self.pk})
            new = self.sync(peers[crypto.randrange(self.n - 1)], payload)
            self.divide_rounds(new)
            new_c = self.decide_fame()
            self.find_order(new_c)


def test(n_nodes, n_turns, node_cls=GpuNode, **kw):
    """The reference's simulation driver (swirld.py:331-345) over GpuNode."""
    kps = [crypto.crypto_sign_keypair() for _ in range(n_nodes)]
    network = {}
    stake = {kp[0]: 1 for kp in kps}
    nodes = [node_cls(kp, network, n_nodes, stake, **kw) for kp in kps]
    for n in nodes:
        network[n.pk] = n.ask_sync
    mains = [n.main() for n in nodes]
    for m in mains:
        next(m)
    for i in range(n_turns):
        r = crypto.randrange(n_nodes)
        print("working node: %i, event number: %i" % (r, i))
        next(mains[r])
    return nodes
