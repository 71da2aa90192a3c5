"""Shared comparison helpers for the parity tests."""
from __future__ import annotations

import numpy as np

import golden_specs as gs

KEYS = ["round", "witness_table", "famous", "consensus", "transactions"]


def load_golden(name):
    z = np.load(gs.path(name))
    g = {k: z[k] for k in z.files}
    offs = g["new_c_offs"]
    g["new_c_per_call"] = [g["new_c_flat"][offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]
    return g


def witness_flags_from_table(wt, n):
    wit = np.zeros(n, np.uint8)
    w = wt[wt >= 0]
    wit[w] = 1
    return wit


def describe_mismatch(key, a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return "%s: shape %s vs %s" % (key, a.shape, b.shape)
    d = np.argwhere(a != b)
    first = d[:5].tolist()
    vals = [(a[tuple(i)].item(), b[tuple(i)].item()) for i in d[:5]]
    return "%s: %d mismatches, first at %s (expected, got) = %s" % (key, len(d), first, vals)


def assert_same(expected, got, keys=KEYS, what=""):
    """Bit-exact comparison (integer / index work: no tolerance)."""
    errs = []
    for k in keys:
        if not np.array_equal(np.asarray(expected[k]), np.asarray(got[k])):
            errs.append(describe_mismatch(k, expected[k], got[k]))
    if "new_c_per_call" in expected and "new_c_per_call" in got:
        e, g = expected["new_c_per_call"], got["new_c_per_call"]
        if [list(x) for x in e] != [list(x) for x in g]:
            bad = [i for i, (x, y) in enumerate(zip(e, g)) if list(x) != list(y)][:3]
            errs.append("new_c_per_call differs at calls %s: %s vs %s" % (
                bad, [e[i] for i in bad], [g[i] for i in bad]))
    assert not errs, what + " :: " + " | ".join(errs)
