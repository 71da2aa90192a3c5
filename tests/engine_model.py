"""Executable model of the ENGINE's reformulation of the hot path (DESIGN.md
"Kernel math"), in plain Python ints-as-bitmasks.  Test infrastructure: it is
how the index-space / bit-matrix restatement the CUDA kernels implement was
proven against the literal oracle before any kernel was written, and it stays
as a CPU test so a change to the math is caught without a GPU.

Nothing here is on the product path.

Notation (fork-free graphs, event id == arrival index):
  row(h)[c]   latest event of member c that h can see (can_see, swirld.py:72)
  W[r][c]     member c's round-r witness (witnesses[r][c], swirld.py:61), -1 none
  SM(h)       M-bit mask {c_ : W[round h][c_] >= 0 and row(h)[c_] >= W[round h][c_]}
  T(h)[c_]    M-bit mask over members c: "h sees a round-(round h) event of c that
              sees W[round h][c_]"  -- the transposed strongly-sees matrix; it
              obeys  T(h) = T(p0) | T(p1) | own-term  inside one round.
"""
from __future__ import annotations

import struct


def popc(x):
    return bin(x).count("1")


class Model:
    def __init__(self, M, stake=None, C=6):
        self.M = M
        self.stake = [1] * M if stake is None else list(stake)
        self.T2 = 2 * sum(self.stake)      # compare 3*x > 2*tot  (== x > 2*tot/3)
        self.tot = sum(self.stake)
        self.C = C
        self.unit = all(s == 1 for s in self.stake)
        self.p0, self.p1, self.cr, self.t, self.sig = [], [], [], [], []
        self.row, self.T, self.SM, self.round, self.wit = [], [], [], [], []
        self.W = []            # W[r] = list of M event ids
        self.S = []            # S[r][m] = strongly-seen mask of witness W[r][m] (decide_fame's s)
        self.famous = []       # famous[r][m] in {-1, 0, 1}
        self.consensus = set()
        self.lastord = [-1] * M   # per chain: latest ordered event
        self.transactions = []
        self.n_done = 0

    # ---- helpers
    def wsum(self, mask):
        if self.unit:
            return popc(mask)
        s, c = 0, 0
        while mask:
            if mask & 1:
                s += self.stake[c]
            mask >>= 1
            c += 1
        return s

    def _round_tables(self, r):
        while len(self.W) <= r:
            self.W.append([-1] * self.M)
            self.S.append([0] * self.M)
            self.famous.append([-1] * self.M)

    def append(self, tr):
        self.p0 += tr.p0.tolist(); self.p1 += tr.p1.tolist(); self.cr += tr.creator.tolist()
        self.t += tr.t.tolist(); self.sig += [bytes(tr.sig[i]) for i in range(tr.N)]

    # ---- K1 + K2: can_see rows, rounds, witnesses (swirld.py:187-222)
    def divide_rounds(self, first, n):
        M = self.M
        for h in range(first, first + n):
            pa, pb, cr = self.p0[h], self.p1[h], self.cr[h]
            if pa < 0:
                row = [-1] * M; row[cr] = h
                rh, wit, t, promoted = 0, True, [0] * M, True
            else:
                ra, rb = self.round[pa], self.round[pb]
                r = max(ra, rb)
                row = [max(a, b) for a, b in zip(self.row[pa], self.row[pb])]
                Ta = self.T[pa] if ra == r else [0] * M
                Tb = self.T[pb] if rb == r else [0] * M
                t = [a | b for a, b in zip(Ta, Tb)]
                cnt = sum(1 for c_ in range(M) if 3 * self.wsum(t[c_]) > self.T2)
                promoted = 3 * cnt > self.T2
                rh = r + (1 if promoted else 0)
                wit = rh > ra
                row[cr] = h
            self._round_tables(rh)
            if wit:
                self.W[rh][cr] = h
            Wr = self.W[rh]
            sm = 0
            for c_ in range(M):
                if Wr[c_] >= 0 and row[c_] >= Wr[c_]:
                    sm |= 1 << c_
            base = [0] * M if promoted else t
            Th = [base[c_] | (((sm >> c_) & 1) << cr) for c_ in range(M)]
            self.row.append(row); self.T.append(Th); self.SM.append(sm)
            self.round.append(rh); self.wit.append(wit)
            if wit and rh >= 1:
                self.S[rh][cr] = self._strongly_seen(h, rh - 1)
        self.n_done = first + n

    def _strongly_seen(self, y, r):
        """decide_fame's s(y) (swirld.py:245-254) as a member mask: live columns
        are those whose latest seen event has round EXACTLY r (quirk Q15)."""
        M = self.M
        hits = [0] * M
        for c in range(M):
            k = self.row[y][c]
            if k >= 0 and self.round[k] == r:
                m = self.SM[k]
                for c_ in range(M):
                    if (m >> c_) & 1:
                        hits[c_] += self.stake[c]
        s = 0
        for c_ in range(M):
            if 3 * hits[c_] > self.T2:
                s |= 1 << c_
        return s

    # ---- K3: decide_fame (swirld.py:224-277)
    def decide_fame(self):
        M = self.M
        max_r = len(self.W) - 1
        while max_r >= 0 and all(w < 0 for w in self.W[max_r]):
            max_r -= 1
        max_c = 0
        while max_c in self.consensus:
            max_c += 1
        done = set()
        V = {}          # (r, m) -> mask over voter members of the previous voter round
        for r_ in range(max_c + 1, max_r + 1):
            voters = [m for m in range(M) if self.W[r_][m] >= 0]
            Vn = {}
            for r in range(max_c, r_):
                if r in self.consensus:
                    continue
                for mx in range(M):
                    if self.W[r][mx] < 0 or self.famous[r][mx] >= 0:
                        continue
                    d = r_ - r
                    mask, decided = 0, None
                    for m in voters:
                        s = self.S[r_][m]
                        if d == 1:
                            vote = (s >> mx) & 1
                        else:
                            prev = V[(r, mx)]
                            yes = self.wsum(s & prev)
                            no = self.wsum(s) - yes
                            v = 0 if no > yes else 1
                            tt = max(yes, no)
                            if d % self.C != 0:
                                if 3 * tt > self.T2:
                                    if decided is None:
                                        decided = v
                                    continue
                                vote = v
                            else:
                                vote = v if 3 * tt > self.T2 else (self.sig[self.W[r_][m]][0] >> 7)
                        mask |= vote << m
                    if decided is not None:
                        self.famous[r][mx] = decided
                        done.add(r)
                    else:
                        Vn[(r, mx)] = mask
            V = Vn
        new_c = set()
        for r in done:
            if all(self.famous[r][m] >= 0 for m in range(M) if self.W[r][m] >= 0):
                new_c.add(r)
        self.consensus |= new_c
        return sorted(new_c)

    # ---- K4: find_order (swirld.py:280-311)
    def find_order(self, new_c):
        M = self.M
        for r in sorted(new_c):
            fw = [self.W[r][m] for m in range(M) if self.W[r][m] >= 0 and self.famous[r][m] == 1]
            white = bytes(64)
            for w in fw:
                white = bytes(a ^ b for a, b in zip(white, self.sig[w]))
            batch = []
            for c in range(M):
                # reach: candidates the reference's BFS visits on chain c
                U = max([self.row[w][c] for w in fw if w > self.lastord[self.cr[w]]] + [-1])
                # received: > half the stake of fw sees it
                vals = sorted(((self.row[w][c], self.stake[self.cr[w]]) for w in fw), reverse=True)
                acc, thr = 0, -1
                for v, s in vals:
                    acc += s
                    if 2 * acc > self.tot:
                        thr = v
                        break
                cut = min(U, thr)
                x = cut
                chain = []
                while x > self.lastord[c] and x >= 0:
                    chain.append(x)
                    x = self.p0[x]
                if chain:
                    self.lastord[c] = chain[0]
                batch += chain
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
                if (n + 1) // 2 >= n:
                    raise IndexError("list index out of range")
                ts = .5 * (times[n // 2] + times[(n + 1) // 2])
                keys.append((ts, bytes(a ^ b for a, b in zip(white, self.sig[x])), x))
            keys.sort()
            self.transactions += [k[2] for k in keys]


def run_model(tr, K, stake=None, C=6):
    import numpy as np
    from swirld_b200.traces import chunks
    m = Model(tr.M, stake, C)
    m.append(tr)
    ncs = []
    for first, cnt in chunks(tr.N, K):
        m.divide_rounds(first, cnt)
        nc = m.decide_fame()
        m.find_order(nc)
        ncs.append(nc)
    R = len(m.W)
    while R > 0 and all(w < 0 for w in m.W[R - 1]):
        R -= 1
    fam = np.full(tr.N, -1, dtype=np.int8)
    wit = np.zeros(tr.N, dtype=np.uint8)
    for r in range(R):
        for c in range(tr.M):
            w = m.W[r][c]
            if w >= 0:
                wit[w] = 1
                fam[w] = m.famous[r][c]
    return {
        "round": np.array(m.round, dtype=np.int32),
        "witness": wit,
        "witness_table": np.array(m.W[:R], dtype=np.int32).reshape(R, tr.M),
        "famous": fam,
        "consensus": np.array(sorted(m.consensus), dtype=np.int32),
        "transactions": np.array(m.transactions, dtype=np.int32),
        "can_see": np.array(m.row, dtype=np.int32),
        "new_c_per_call": ncs,
    }
