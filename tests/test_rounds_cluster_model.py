"""Executable model of the cluster round kernel (py-swirld_b200/csrc/swirld_rcluster.cuh): the round-batch scheme of
swirld_rounds.cuh with everything a step needs held inside ONE thread-block cluster --

  * rows in SEQ space (rs[h][c] = chain position of the event of member c that h sees), a sliding window of WN rows
    per chain in the shared memory of the CTA that owns the chain, loaded one step ahead;
  * per step the masks S_r of the members' ranges [Wls_r[m], mend[m]) (at most MR per member), pushed to every CTA;
  * the first pending event of a chain that passes P_r, or that the masks do not cover: both are monotone along a
    chain, so the kernel may test every position of the window (unit stake, bit-sliced) or search it (WPC+1)-ary
    (integer stakes) -- the model searches, and asserts that the search closes;
  * anything the windows cannot decide (no progress for STALL steps, a round outside the mirror, an event beyond the
    ring) hands the rest of the chunk to the grid-wide kernel (modelled by RoundBatch with a start state).

The model makes the same decisions as the kernel (same window arithmetic) and is checked against the literal oracle."""
import numpy as np
import pytest

import oracle as orc
from swirld_b200 import traces


class ClusterRounds:
    def __init__(self, tr, rows, stake=None, CPC=4, WPC=4, LW=32, WN=128, MR=64, PF=32, WR=32, RING=None, STALL=3):
        self.tr, self.M = tr, tr.M
        RING = WN // 2 if RING is None else RING         # rows before the chunk come from the ring, at launch only
        self.CPC, self.WPC, self.LW, self.WN, self.MR, self.PF, self.WR, self.RING, self.STALL = CPC, WPC, LW, WN, MR, PF, WR, RING, STALL
        self.passes = 1
        while (WPC + 1) ** self.passes < LW + 1:
            self.passes += 1
        M, N = tr.M, tr.N
        self.stake = np.array([1] * M if stake is None else list(stake), np.int64)
        self.tot2 = 2 * int(self.stake.sum())
        self.seq = np.zeros(N, np.int64)
        self.chain = [[] for _ in range(M)]
        for h in range(N):
            c = tr.creator[h]
            self.seq[h] = len(self.chain[c])
            self.chain[c].append(h)
        rows = np.asarray(rows)
        self.rs = np.where(rows >= 0, self.seq[np.maximum(rows, 0)], -1)          # rows in seq space
        self.round = np.full(N, -1, np.int64)
        self.Wf = {}
        self.max_round = 0
        self.steps = self.bails = self.tests = self.unknowns = 0

    def wf(self, r):
        return self.Wf.setdefault(r, [-1] * self.M)

    # ---- v(j): 0 = P false, 1 = P true, 2 = an event the prepared masks do not cover
    def v(self, c, sq, lo, mend):
        self.tests += 1
        M = self.M
        h = self.chain[c][sq]
        pre = self.rs[h].copy()
        pre[c] = sq - 1
        live = (lo >= 0) & (pre >= lo)
        lv = int(self.stake[live].sum())
        if 3 * lv <= self.tot2:
            return 0
        if np.any(live & (pre >= mend)):
            self.unknowns += 1
            return 2
        hits = np.zeros(M, np.int64)
        for m in np.nonzero(live)[0]:
            k = self.chain[m][pre[m]]
            hits += self.stake[m] * ((lo >= 0) & (self.rs[k] >= lo))
        return 1 if 3 * int((3 * hits > self.tot2).sum()) > self.tot2 else 0

    def divide(self, first, n):
        """Returns None when the cluster kernel finished the chunk, else (pos, cur) where it gave up."""
        tr, M, seq = self.tr, self.M, self.seq
        LW, WN, MR, PF, WR = self.LW, self.WN, self.MR, self.PF, self.WR
        ev = [[h for h in range(first, first + n) if tr.creator[h] == c] for c in range(M)]
        ln = [len(e) for e in ev]
        cmin = [int(seq[e[0]]) if e else 0 for e in ev]
        ctot = [sum(1 for h in self.chain[c] if h < first + n) for c in range(M)]
        INF = 1 << 30
        pos, cur = [0] * M, [INF] * M
        rtop = self.max_round
        Wls = {}
        for r in range(max(0, rtop - WR + 1), rtop + 1):
            Wls[r & (WR - 1)] = [int(seq[w]) if w >= 0 else -1 for w in self.wf(r)]
        for slot in range(WR):
            Wls.setdefault(slot, [-1] * M)
        for c in range(M):
            if ln[c]:
                h0 = ev[c][0]
                cur[c] = 0 if tr.p0[h0] < 0 else int(self.round[tr.p0[h0]])
                if tr.p0[h0] < 0:
                    self.wf(0)[c] = h0
                    if rtop < WR:
                        Wls[0][c] = 0
        is_open = lambda c: pos[c] < ln[c]
        seqpos = lambda c: cmin[c] + pos[c] if ln[c] else ctot[c]
        before = lambda c: cmin[c] if ln[c] else ctot[c]

        def give_up():
            self.bails += 1
            self.max_round = rtop
            return list(pos), list(cur)

        open_c = [c for c in range(M) if is_open(c)]
        if not open_c:
            return None
        rmin = min(cur[c] for c in open_c)
        if rmin <= rtop - WR:
            return give_up()
        # ---- the windows: [wlo, wrd) is in shared memory, [wrd, wld) on its way
        wlo, wld, wrd = [0] * M, [0] * M, [0] * M
        for c in range(M):
            lo = Wls[rmin & (WR - 1)][c]
            wlo[c] = min(lo, seqpos(c)) if lo >= 0 else seqpos(c)
            if before(c) - wlo[c] > self.RING:
                return give_up()
            wld[c] = wrd[c] = min(ctot[c], wlo[c] + WN, seqpos(c) + LW + PF)
            assert wld[c] >= before(c) or wld[c] == ctot[c], "rows before the chunk are loaded at launch only"
        stall = 0
        while True:
            open_c = [c for c in range(M) if is_open(c)]
            if not open_c:
                break
            rmin = min(cur[c] for c in open_c)
            if rmin <= rtop - WR:
                return give_up()
            self.steps += 1
            lo = np.array(Wls[rmin & (WR - 1)], np.int64)
            mend = np.full(M, -1, np.int64)
            for m in range(M):
                if lo[m] >= 0:
                    assert lo[m] >= wlo[m], "the window dropped rows the masks of this round need"
                    mend[m] = min(wrd[m], seqpos(m) + LW if is_open(m) else ctot[m], lo[m] + MR)
            progress = False
            opened = []
            for c in range(M):
                if not (is_open(c) and cur[c] == rmin):
                    continue
                win = min(LW, ln[c] - pos[c], wrd[c] - seqpos(c))
                a, b, vb = -1, max(win, 0), 0
                for _ in range(self.passes):
                    nun = b - a - 1
                    if nun <= 0:
                        break
                    ts = [a + 1 + i for i in range(nun)] if nun < self.WPC else \
                         [a + ((i + 1) * (nun + 1)) // (self.WPC + 1) for i in range(self.WPC)]
                    assert len(set(ts)) == len(ts) and all(a < t < b for t in ts)
                    for t in ts:
                        r = self.v(c, seqpos(c) + t, lo, mend)
                        if r == 0:
                            a = max(a, t)
                        elif t < b:
                            b, vb = t, r
                assert b - a == 1 or win <= 0, "the search did not close"
                f, vf = b, (vb if b < win else 0)
                for j in range(f):
                    self.round[ev[c][pos[c] + j]] = rmin
                if vf == 1:
                    opened.append((c, cmin[c] + pos[c] + f, ev[c][pos[c] + f]))
                    cur[c] = rmin + 1
                progress |= f > 0 or vf == 1
                pos[c] += f
            if opened and rmin + 1 > rtop:
                rtop = rmin + 1
                Wls[rtop & (WR - 1)] = [-1] * M
            for c, sq, h in opened:
                Wls[(rmin + 1) & (WR - 1)][c] = sq
                self.wf(rmin + 1)[c] = h
            # ---- slide the windows: what was issued a step ago is there now; issue the next rows
            open_c = [c for c in range(M) if is_open(c)]
            rnext = min((cur[c] for c in open_c), default=rmin)
            moved = False
            for c in range(M):
                moved |= wrd[c] != wld[c]
                wrd[c] = wld[c]
                if rnext > rtop - WR:
                    l2 = Wls[rnext & (WR - 1)][c]
                    keep = min(l2, seqpos(c)) if l2 >= 0 else seqpos(c)
                    wlo[c] = max(wlo[c], keep)
                hi = min(ctot[c], wlo[c] + WN, seqpos(c) + LW + PF)
                if hi > wld[c]:
                    wld[c] = hi
                    moved = True
            stall = 0 if (progress or moved) else stall + 1
            if stall >= self.STALL:
                return give_up()
        self.max_round = rtop
        return None


class Continuation:
    """The grid-wide kernel taking over from (pos, cur): windows of L events per chain, every position tested."""

    def __init__(self, cl, L=33):
        self.cl, self.L = cl, L

    def P(self, h, r):
        cl = self.cl
        W = np.array(cl.wf(r), np.int64)
        lo = np.where(W >= 0, cl.seq[np.maximum(W, 0)], -1)
        c = cl.tr.creator[h]
        big = np.full(cl.M, 1 << 40, np.int64)
        return cl.v(c, int(cl.seq[h]), lo, big) == 1

    def divide(self, first, n, pos, cur):
        cl = self.cl
        tr, M = cl.tr, cl.M
        ev = [[h for h in range(first, first + n) if tr.creator[h] == c] for c in range(M)]
        while True:
            act = [c for c in range(M) if pos[c] < len(ev[c])]
            if not act:
                break
            r = min(cur[c] for c in act)
            opened = []
            for c in act:
                if cur[c] != r:
                    continue
                win = ev[c][pos[c]:pos[c] + self.L]
                ft = next((i for i, h in enumerate(win) if tr.p0[h] >= 0 and self.P(h, r)), None)
                for h in win[:len(win) if ft is None else ft]:
                    cl.round[h] = r
                pos[c] += len(win) if ft is None else ft
                if ft is not None:
                    opened.append((c, win[ft]))
                    cur[c] = r + 1
            for c, h in opened:
                cl.wf(r + 1)[c] = h
                cl.max_round = max(cl.max_round, r + 1)


def run(tr, stake, chunks, **kw):
    o = orc.Oracle(tr.M, stake)
    o.append(tr)
    o.divide_rounds(0, tr.N)
    cl = ClusterRounds(tr, o.can_see(), stake, **kw)
    first = 0
    for n in chunks:
        state = cl.divide(first, n)
        if state is not None:
            Continuation(cl).divide(first, n, *state)
        first += n
    assert np.array_equal(cl.round, o.results()["round"])
    return cl


@pytest.mark.parametrize("gen,M,N,chunks,stake", [
    ("gossip", 4, 600, [600], None), ("gossip", 4, 300, [1] * 300, None),
    ("gossip", 8, 1500, [100] * 15, None), ("adversarial", 8, 1500, [250] * 6, None),
    ("tick", 16, 2000, [700, 1300], None), ("gossip", 7, 1507, [11] * 137, [1, 1, 2, 1, 1, 1, 0]),
    ("gossip", 33, 4000, [1500, 2500], None), ("gossip", 64, 6000, [2500, 3500], None),
    ("adversarial", 64, 6000, [6000], None)])
def test_cluster_rounds_equal_oracle(gen, M, N, chunks, stake):
    tr = getattr(traces, gen)(M, N, 3)
    cl = run(tr, stake, chunks)
    assert cl.steps > 0


@pytest.mark.parametrize("kw", [dict(WN=16, LW=8, PF=4, MR=8), dict(WN=8, LW=4, PF=2, MR=4, WPC=2), dict(WR=4), dict(RING=8),
                                dict(LW=32, WPC=2), dict(LW=5, WPC=4, WN=32, PF=3)])
def test_cluster_rounds_small_windows_fall_back(kw):
    """Windows too small for the graph: events are left undecided, the chunk is handed over -- same rounds."""
    for gen, M, N, chunks in [("gossip", 8, 1200, [300] * 4), ("adversarial", 16, 2500, [2500]), ("tick", 16, 1500, [500, 1000])]:
        tr = getattr(traces, gen)(M, N, 5)
        run(tr, None, chunks, **kw)
