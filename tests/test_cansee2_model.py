"""Executable model of the column-tiled can_see scan of py-swirld_b200/csrc/swirld_cansee.cuh
(round 2): prep (stale flags, referenced rows, per-block last events), pass 1 with a per-member
value cache that writes only the rows something else will read, block-start heads by a scan over
the blocks, a PARALLEL finality check of the listed rows (rows that fail it are finished block by
block), and the exact pass 2 -- against the literal oracle, on CPU.  Every column is independent
in the kernels (one thread per column, column tiles as separate CTAs); the model keeps whole rows."""
import numpy as np
import pytest

import oracle as orc
from swirld_b200 import traces


def stale_flags(tr):
    """stale[h]: the other-parent is not its member's latest event below h (host side of sw_append)."""
    head = np.full(tr.M, -1, np.int64)
    st = np.zeros(tr.N, bool)
    for h in range(tr.N):
        b = tr.p1[h]
        if b >= 0:
            st[h] = head[tr.creator[b]] != b
        head[tr.creator[h]] = h
    return st


def scan_launch(tr, stale, first, n, B, row, carry, stats):
    M, N = tr.M, tr.N
    p0, p1, cr = tr.p0, tr.p1, tr.creator
    ar = np.arange(M)
    # blocks start at multiples of 4 in absolute index space (groups of four events are the pipelining unit)
    starts = [first]
    nxt = (first // 4) * 4 + B
    while nxt < first + n:
        starts.append(nxt)
        nxt += B
    if len(starts) > 1 and first + n - starts[-1] < 3 * B // 4:
        starts.pop()                                  # a short tail joins the block before it
    nb = len(starts)
    ends = starts[1:] + [first + n]
    blk_of = np.zeros(N, np.int64)
    for j in range(nb):
        blk_of[starts[j]:ends[j]] = j
    # ---- prep: rows somebody reads (wr), rows a LATER block reads (xb), last event of a member per block
    wr = np.zeros(N, bool)
    xb = np.zeros(N, bool)
    last = np.full((nb, M), -1, np.int64)
    for h in range(first, first + n):
        for p, is_other in ((p0[h], False), (p1[h], True)):
            if p >= first:
                if blk_of[p] != blk_of[h]:
                    xb[p] = wr[p] = True
                elif is_other and stale[h]:
                    wr[p] = True
        last[blk_of[h], cr[h]] = h
    # ---- pass 1: out-of-block parents are leaves; only wr rows and the last rows are written
    partial = {}
    for j in range(nb):
        s = starts[j]
        val = np.full((M, M), -1, np.int64)          # val[m] = cached row of member m's latest in-block event
        for h in range(s, ends[j]):
            a, b, c = p0[h], p1[h], cr[h]
            v = np.full(M, -1, np.int64)
            if a >= 0:
                x = val[c] if a >= s else np.full(M, -1, np.int64)      # (the leaf value sits in column c, overwritten below)
                if b >= s:
                    y = partial[b] if stale[h] else val[cr[b]]
                else:
                    y = np.where(ar == cr[b], b, -1)
                v = np.maximum(x, y)
            v[c] = h
            val[c] = v
            if wr[h]:
                partial[h] = v.copy()
                stats["p1_rows"] += 1
        for m in range(M):
            if last[j, m] >= 0:
                partial[int(last[j, m])] = val[m].copy()
                stats["p1_rows"] += 1
    for h, v in partial.items():
        row[h] = v
    # ---- heads at the start of every block
    Q = np.full((nb + 1, M), -1, np.int64)
    Q[0] = carry
    for j in range(nb):
        Q[j + 1] = np.where(last[j] >= 0, last[j], Q[j])
    # ---- finality check of the listed rows, all blocks at once; the rest block by block
    listed = [[] for _ in range(nb)]
    for h in range(first, first + n):
        if xb[h] or last[blk_of[h], cr[h]] == h:
            listed[blk_of[h]].append(h)
    slow = [[] for _ in range(nb)]
    for j in range(nb):
        for x in listed[j]:
            pr = row[x]
            inb = pr >= starts[j]
            if np.all(inb | (pr == Q[j])):
                stats["fast"] += 1
            else:
                slow[j].append(x)
    # in dependency waves, column by column: an open column (neither in-block nor the head itself) of a pending row is the
    # max over the row's entry events of THEIR column -- final unless the entry is itself a row in flux whose column is open
    pending = {x: j for j in range(nb) for x in slow[j]}

    def open_cols(x, j):
        pr = row[x]
        return (pr < starts[j]) & (pr != Q[j])
    while pending:
        flux = dict(pending)                         # pending, or finished in this very wave
        finished = []
        progress = False
        for x, j in list(pending.items()):
            pr = row[x].copy()
            inb = pr >= starts[j]
            ent = [Q[j, m] if (inb[m] or pr[m] == Q[j, m]) else pr[m] for m in range(M)]
            complete = True
            for cb in np.nonzero(open_cols(x, j))[0]:
                wait, acc = False, -1
                for e in ent:
                    if e < 0:
                        continue
                    assert e < starts[j]
                    acc = max(acc, row[e][cb])
                    if e in flux and open_cols(e, flux[e])[cb]:
                        wait = True
                if wait:
                    complete = False
                    continue
                if acc > row[x][cb]:
                    row[x][cb] = acc
                    progress = True
            if complete:
                finished.append(x)
        assert finished, "a wave must finish at least the lowest block's rows"
        for x in finished:
            stats["slow"] += 1
            del pending[x]
        stats["waves"] = stats.get("waves", 0) + 1
    # ---- pass 2: the exact rows; out-of-block parents contribute their final rows
    for j in range(nb):
        s = starts[j]
        val = np.full((M, M), -1, np.int64)
        for m in range(M):
            if Q[j, m] >= 0:
                val[m] = row[Q[j, m]]
        for h in range(s, ends[j]):
            a, b, c = p0[h], p1[h], cr[h]
            v = np.full(M, -1, np.int64)
            if a >= 0:
                assert a >= s or a == Q[j, c]
                x = val[c]
                if stale[h]:
                    assert b >= s or b < first or xb[b]
                    y = row[b]
                    stats["table"] += 1
                else:
                    assert b >= s or b == Q[j, cr[b]]
                    y = val[cr[b]]
                v = np.maximum(x, y)
            v[c] = h
            val[c] = v
            row[h] = v
    return Q[nb].copy()


@pytest.mark.parametrize("M,N,B,chunks", [(8, 600, 64, [600]), (8, 601, 64, [101, 199, 301]), (4, 300, 16, [7, 293]),
                                          (16, 2000, 128, [700, 1300]), (16, 2500, 512, [2500]), (5, 200, 16, [1] * 200),
                                          (80, 4000, 512, [1500, 2500]), (130, 5000, 1024, [5000])])
@pytest.mark.parametrize("gen", ["gossip", "adversarial", "tick"])
def test_tiled_scan_equals_oracle(M, N, B, chunks, gen):
    tr = getattr(traces, gen)(M, N, 3)
    o = orc.Oracle(M)
    o.append(tr)
    o.divide_rounds(0, N)
    ref = o.can_see()
    st = stale_flags(tr)
    row = np.full((N, M), -1, np.int64)
    carry = np.full(M, -1, np.int64)
    stats = {"fast": 0, "slow": 0, "table": 0, "p1_rows": 0}
    first = 0
    for n in chunks:
        carry = scan_launch(tr, st, first, n, B, row, carry, stats)
        first += n
    assert np.array_equal(row, ref)
    if gen == "gossip":
        assert stats["table"] == 0           # G1 has no stale other-parents: pass 2 never reads the table
