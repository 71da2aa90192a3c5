"""Executable model of find_order as the kernels split it (swirld_kernels.cuh: k_order_rounds /
k_order_cuts / k_order_list), on CPU against the literal oracle:

  A  per consensus round, independent of what earlier rounds ordered: famous witnesses, and per member
     chain the received-threshold `thr` and the reach over ALL famous witnesses (+ their seq numbers);
  B  one sequential pass over the rounds: a famous witness that is already ordered does not seed the
     search (then the reach is recomputed over the others); cut = min(reach, thr); the chain gives the
     events (lastord, cut], counted by seq numbers;
  C  the events are listed from the cut down the self-parent chain.
"""
import numpy as np
import pytest

import engine_model as em
import golden_specs as gs
import oracle as orc

NAMES = ["g1_m4_n2000_s1_k50", "g1_m4_n2000_s3_k7", "g1_m7_n3000_s4_k11_stake",
         "g2_m16_n12000_s1_k500", "g3_m16_n6000_s1_k700", "g1_m33_n6000_s7_k640"]


class SplitOrder(em.Model):
    fallbacks = 0
    calls = 0

    def find_order(self, new_c):
        SplitOrder.calls += 1
        M = self.M
        seq = self._seq()
        rounds = sorted(new_c)
        plans = []
        for r in rounds:                                               # A (parallel over rounds)
            fw = [self.W[r][m] for m in range(M) if self.W[r][m] >= 0 and self.famous[r][m] == 1]
            thr, uall = [-1] * M, [-1] * M
            for c in range(M):
                for w in fw:
                    v = self.row[w][c]
                    uall[c] = max(uall[c], v)
                    if v > thr[c] and 2 * sum(self.stake[self.cr[k]] for k in fw if self.row[k][c] >= v) > self.tot:
                        thr[c] = v
            plans.append((fw, thr, uall))
        lo = list(self.lastord)                                        # B (sequential, 64 threads)
        cuts = []
        for fw, thr, uall in plans:
            tbd = [w > lo[self.cr[w]] for w in fw]
            U = uall
            if not all(tbd):
                SplitOrder.fallbacks += 1
                U = [max([self.row[w][c] for w, ok in zip(fw, tbd) if ok] + [-1]) for c in range(M)]
            per_chain = []
            for c in range(M):
                cut = min(U[c], thr[c])
                cnt = seq[cut] - (seq[lo[c]] if lo[c] >= 0 else -1) if cut > lo[c] else 0
                per_chain.append((cut, cnt))
            for c, (cut, cnt) in enumerate(per_chain):
                if cnt > 0:
                    lo[c] = cut
            cuts.append(per_chain)
        for (fw, thr, uall), per_chain in zip(plans, cuts):            # C + the unchanged timing / sort
            batch = []
            for c, (cut, cnt) in enumerate(per_chain):
                x = cut
                for _ in range(cnt):
                    batch.append(x)
                    x = self.p0[x]
            self._order_batch(fw, batch)
        self.lastord = lo

    def _seq(self):
        cnt, out = [0] * self.M, []
        for c in self.cr:
            out.append(cnt[c])
            cnt[c] += 1
        return out

    def _order_batch(self, fw, batch):
        white = bytes(64)
        for w in fw:
            white = bytes(a ^ b for a, b in zip(white, self.sig[w]))
        keys = []
        for x in batch:
            c = self.cr[x]
            times = []
            for w in fw:
                if self.row[w][c] >= x:
                    a = w
                    while self.row[a][c] >= x and self.p0[a] >= 0:
                        a = self.p0[a]
                    times.append(self.t[a])
            times.sort()
            n = len(times)
            ts = .5 * (times[n // 2] + times[(n + 1) // 2])
            keys.append((ts, bytes(a ^ b for a, b in zip(white, self.sig[x])), x))
        keys.sort()
        self.transactions += [k[2] for k in keys]


@pytest.mark.parametrize("name", NAMES)
def test_split_find_order_matches_oracle(name, monkeypatch):
    tr, K, stake = gs.make_trace(name)
    o = orc.run_oracle(tr, K, stake)
    monkeypatch.setattr(em, "Model", SplitOrder)
    before = SplitOrder.calls
    m = em.run_model(tr, K, stake)
    assert SplitOrder.calls > before
    assert np.array_equal(o["transactions"], m["transactions"]), name


def test_ordered_witness_fallback_is_exercised():
    """An adversarial trace with long-undecided rounds: some consensus round has a famous witness that an
    earlier round already ordered, so the reach must be recomputed without it."""
    from swirld_b200 import traces
    SplitOrder.fallbacks = 0
    hit = False
    for seed in range(1, 6):
        tr = traces.adversarial(8, 4000, seed, 0.02, 0.5)
        o = orc.run_oracle(tr, 37)
        saved = em.Model
        em.Model = SplitOrder
        try:
            m = em.run_model(tr, 37)
        finally:
            em.Model = saved
        assert np.array_equal(o["transactions"], m["transactions"])
        hit = hit or SplitOrder.fallbacks > 0
    print("fallbacks:", SplitOrder.fallbacks)
