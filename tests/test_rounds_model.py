"""Executable model of the round-batch scheme of py-swirld_b200/csrc/swirld_rounds.cuh:
round[h] >= r+1  <=>  P_r(h), with P_r evaluated on per-member windows of pending events,
rounds ascending, against the literal oracle (CPU).  Covers several launches, one-event
launches, stale other-parents, integer stakes and 64 members."""
import numpy as np
import pytest

import oracle as orc
from swirld_b200 import traces


class RoundBatch:
    """L: pending events tested per chain and step (swirld_rounds.cuh).  frontier=True: the windows of
    swirld_wide.cuh instead -- the pending events with an index below X = (lowest pending index at rmin) + L*M,
    at most LMAX per chain.  exchange: callable(step results of MY chains) -> results of all chains (several
    ranks: every rank evaluates P only for the chains c with c % nranks == rank)."""
    LMAX = 32

    def __init__(self, tr, rows, stake=None, L=8, frontier=False, rank=0, nranks=1, exchange=None):
        self.tr, self.M, self.rows, self.L = tr, tr.M, rows, L
        self.frontier, self.rank, self.nranks, self.exchange = frontier, rank, nranks, exchange
        self.stake = [1] * tr.M if stake is None else list(stake)
        self.tot2 = 2 * sum(self.stake)
        self.round = np.full(tr.N, -1, np.int64)
        self.Wf = {}                                   # r -> first event of round >= r per member
        self.headround = [-1] * tr.M
        self.steps = 0

    def wf(self, r):
        return self.Wf.setdefault(r, [-1] * self.M)

    def P(self, h, r):
        M, W, tr = self.M, self.wf(r), self.tr
        pre = self.rows[h].copy()
        pre[tr.creator[h]] = tr.p0[h]
        hits = [0] * M
        for c in range(M):
            k = pre[c]
            if k >= 0 and W[c] >= 0 and k >= W[c]:
                rk = self.rows[k]
                for c_ in range(M):
                    if W[c_] >= 0 and rk[c_] >= W[c_]:
                        hits[c_] += self.stake[c]
        return 3 * sum(1 for c_ in range(M) if 3 * hits[c_] > self.tot2) > self.tot2

    def divide(self, first, n):
        tr, M = self.tr, self.M
        chains = [[] for _ in range(M)]
        for h in range(first, first + n):
            chains[tr.creator[h]].append(h)
        pos, cur = [0] * M, [0] * M
        for c in range(M):
            if chains[c]:
                h0 = chains[c][0]
                if tr.p0[h0] < 0:
                    cur[c] = 0
                    if self.wf(0)[c] < 0:
                        self.wf(0)[c] = h0
                else:
                    cur[c] = self.round[tr.p0[h0]]
        while True:
            act = [c for c in range(M) if pos[c] < len(chains[c])]
            if not act:
                break
            r = min(cur[c] for c in act)
            self.steps += 1
            opened = {}
            at_r = [c for c in act if cur[c] == r]
            if self.frontier:
                X = min(chains[c][pos[c]] for c in at_r) + self.L * M
                wins = {c: [h for h in chains[c][pos[c]:pos[c] + self.LMAX] if h < X] for c in at_r}
            else:
                wins = {c: chains[c][pos[c]:pos[c] + self.L] for c in at_r}
            # first hit per chain: my share of the chains, then the exchange
            mine = {c: next((i for i, h in enumerate(wins[c]) if tr.p0[h] >= 0 and self.P(h, r)), None)
                    for c in at_r if c % self.nranks == self.rank}
            hits = self.exchange(mine) if self.exchange else mine
            for c in at_r:
                win = wins[c]
                ft = hits[c]
                for h in win[:len(win) if ft is None else ft]:
                    self.round[h] = r
                pos[c] += len(win) if ft is None else ft
                if ft is not None:
                    opened[c] = win[ft]
                    cur[c] = r + 1
            for c, h in opened.items():
                self.wf(r + 1)[c] = h


@pytest.mark.parametrize("gen,M,N,chunks,stake,L", [
    ("gossip", 4, 600, [600], None, 8), ("gossip", 4, 400, [1] * 400, None, 4),
    ("gossip", 8, 1500, [100] * 15, None, 8), ("adversarial", 8, 1500, [250] * 6, None, 8),
    ("tick", 16, 2000, [700, 1300], None, 8), ("gossip", 7, 1507, [11] * 137, [1, 1, 2, 1, 1, 1, 0], 8),
    ("gossip", 64, 3000, [1200, 1800], None, 18),
    # beyond this build's 64 members: the predicate does not care (member masks become multi-word)
    ("gossip", 96, 7000, [3000, 4000], None, 24), ("adversarial", 130, 9000, [9000], None, 16)])
def test_round_batch_equals_oracle(gen, M, N, chunks, stake, L):
    tr = getattr(traces, gen)(M, N, 3)
    o = orc.Oracle(M, stake)
    o.append(tr)
    o.divide_rounds(0, N)
    rb = RoundBatch(tr, o.can_see(), stake, L)
    first = 0
    for n in chunks:
        rb.divide(first, n)
        first += n
    assert np.array_equal(rb.round, o.results()["round"])
    # the frontier windows of swirld_wide.cuh give the same rounds
    rf = RoundBatch(tr, o.can_see(), stake, max(1, L // 4), frontier=True)
    first = 0
    for n in chunks:
        rf.divide(first, n)
        first += n
    assert np.array_equal(rf.round, o.results()["round"])
