"""Adapter that gives the ORACLE the engine's Python surface, so the CPU test-suite
can exercise GpuNode's host logic (hash<->index maps, views, batching, growth replay)
without a GPU.  Test infrastructure only: the product never sees this class."""
import numpy as np

import oracle as orc
from swirld_b200.traces import Trace


class OracleEngine:
    def __init__(self, M, capacity, stake=None, coin_period=6, device=0):
        self.M, self.capacity = M, capacity
        self._o = orc.Oracle(M, stake, coin_period)
        self._nd = 0

    def close(self):
        self._o.close()

    def append(self, p0, p1, creator, t, sig):
        if self._o.n + len(p0) > self.capacity:
            raise RuntimeError("capacity")
        self._o.append(Trace(self.M, np.asarray(p0, np.int32), np.asarray(p1, np.int32),
                             np.asarray(creator, np.int32), np.asarray(t, np.float64),
                             np.asarray(sig, np.uint8).reshape(-1, 64)))

    def divide_rounds(self, first, n):
        assert first == self._nd
        self._o.divide_rounds(first, n)
        self._nd += n

    def decide_fame(self):
        return self._o.decide_fame()

    def find_order(self, new_c):
        before = orc.lib().or_n_transactions(self._o._h)
        self._o.find_order(new_c)
        return orc.lib().or_n_transactions(self._o._h) - before

    n_events = property(lambda s: s._o.n)
    n_divided = property(lambda s: s._nd)
    max_round = property(lambda s: s._o.max_round)
    n_transactions = property(lambda s: orc.lib().or_n_transactions(s._o._h))

    def rounds(self, first=0, n=None):
        r = self._o.results()["round"][:self._nd]
        return r[first:] if n is None else r[first:first + n]

    def famous(self, first=0, n=None):
        return self._o.results()["famous"]

    def idx(self, first=0, n=None):
        out = np.empty(self._o.n, np.int32)
        orc.lib().or_get_idx(self._o._h, out)
        return out

    def can_see(self, first=0, n=None):
        return self._o.can_see(first, n)

    def witness_table(self, first_round=0, n_rounds=None):
        wt = self._o.results()["witness_table"]
        return wt[first_round:] if n_rounds is None else wt[first_round:first_round + n_rounds]

    def transactions(self, first=0, n=None):
        tx = self._o.results()["transactions"]
        return tx[first:] if n is None else tx[first:first + n]

    def results(self):
        r = self._o.results()
        r["round"] = r["round"][:self._nd]
        return r
