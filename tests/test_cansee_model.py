"""Executable model of the blocked can_see scan (py-swirld_b200/csrc/swirld_cansee.cuh:
partial rows per block with out-of-block parents as leaves, block-start heads as the
representatives of what lies below the block for the rows later blocks depend on, then the
exact in-block walk with the finished out-of-block rows) against the literal oracle, on CPU.  Covers stale other-parents, several launches and one-event launches."""
import numpy as np
import pytest

import oracle as orc
from swirld_b200 import traces


def scan_launch(tr, first, n, B, row, carry, stats):
    M = tr.M
    p0, p1, cr = tr.p0, tr.p1, tr.creator
    exported = np.zeros(tr.N, bool)
    nb = (n + B - 1) // B
    last = np.full((nb, M), -1, np.int64)
    ar = np.arange(M)
    for blk in range(nb):                                   # A: k_cs_local
        s = first + blk * B
        for h in range(s, min(s + B, first + n)):
            pa, pb, c_ = p0[h], p1[h], cr[h]
            v = np.full(M, -1, np.int64)
            if pa >= 0:
                a = row[pa].copy() if pa >= s else np.where(ar == cr[pa], pa, -1)
                b = row[pb].copy() if pb >= s else np.where(ar == cr[pb], pb, -1)
                v = np.maximum(a, b)
                for p in (pa, pb):
                    if first <= p < s:
                        exported[p] = True
            v[c_] = h
            row[h] = v
            last[blk, c_] = h
    Q = np.full((nb + 1, M), -1, np.int64)
    Q[0] = carry
    CM = np.full((nb, M), -1, np.int64)

    def complete(x, lim, blk, fast_ok):
        pr = row[x].copy()
        inb = pr >= lim
        if inb.all():
            stats["final"] += 1
            return
        q = Q[blk]
        if fast_ok and np.all(inb | (pr == q)):
            stats["fast"] += 1      # every out-of-block column shows its member's head: already final
            assert np.array_equal(CM[blk], q)        # (the head of c is the largest value column c takes below the block)
            assert np.array_equal(np.where(inb, pr, np.maximum(pr, CM[blk])), pr)
            return
        stats["slow"] += 1
        acc = pr.copy()
        for m in range(M):
            e = q[m] if pr[m] >= lim else pr[m]
            if e >= 0:
                acc = np.maximum(acc, row[e])
        row[x] = np.where(inb, pr, acc)

    for blk in range(nb):                                   # B: k_cs_boundary
        lim = first + blk * B
        for m in range(M):
            if Q[blk, m] >= 0:
                CM[blk] = np.maximum(CM[blk], row[Q[blk, m]])
        todo = {x for x in range(lim, min(lim + B, first + n)) if exported[x]} | {int(v) for v in last[blk] if v >= 0}
        for x in sorted(todo):
            complete(x, lim, blk, True)
            exported[x] = True
        Q[blk + 1] = np.where(last[blk] >= 0, last[blk], Q[blk])
    for blk in range(nb):                                   # C: k_cs_local<2>, the exact in-block walk
        s = first + blk * B
        # per-member cache (event, its row): the block-start heads with their final rows
        tv = {m: (int(Q[blk, m]), row[Q[blk, m]].copy()) for m in range(M) if Q[blk, m] >= 0}
        for h in range(s, min(s + B, first + n)):
            pa, pb, c_ = p0[h], p1[h], cr[h]
            v = np.full(M, -1, np.int64)
            if pa >= 0:
                def parent_row(p):
                    hit = tv.get(int(cr[p]))
                    if hit is not None and hit[0] == p:
                        stats["cached"] += 1
                        return hit[1]
                    stats["table"] += 1
                    # in-block: written earlier in this pass; out-of-block: finished by B or an earlier launch
                    assert p < s and (p < first or exported[p]) or p >= s
                    return row[p]
                v = np.maximum(parent_row(pa), parent_row(pb))
            v[c_] = h
            row[h] = v
            tv[int(c_)] = (h, v.copy())
    return Q[nb].copy()


@pytest.mark.parametrize("M,N,B,chunks", [(8, 600, 64, [600]), (8, 600, 64, [100, 200, 300]),
                                          (16, 2000, 128, [700, 1300]), (16, 2500, 512, [2500]), (5, 200, 16, [1] * 200), (80, 4000, 512, [1500, 2500])])
@pytest.mark.parametrize("gen", ["gossip", "adversarial", "tick"])
def test_blocked_scan_equals_oracle(M, N, B, chunks, gen):
    tr = getattr(traces, gen)(M, N, 3)
    o = orc.Oracle(M)
    o.append(tr)
    o.divide_rounds(0, N)
    ref = o.can_see()
    row = np.full((N, M), -1, np.int64)
    carry = np.full(M, -1, np.int64)
    stats = {"final": 0, "fast": 0, "slow": 0, "cached": 0, "table": 0}
    first = 0
    for n in chunks:
        carry = scan_launch(tr, first, n, B, row, carry, stats)
        first += n
    assert np.array_equal(row, ref)
