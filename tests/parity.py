"""Parity helpers shared by the tests: canonical comparison of two pass outputs.

Bit-exact means (BASELINE.json north_star): same eliminated-variable set, same
model-reconstruction stack, same clause multiset after canonical sorting.  We check more: the AoS
arena and refs vector word for word (same clause order, flags, signatures), state/value arrays,
the root trail and the scalar counters.
"""
from __future__ import annotations

import numpy as np

from seqfrost_b200 import boundary as B


def canonical(b: B.Boundary):
    """Order-free digest: (eliminated vars, sorted clause multiset, model stack, unit set)."""
    w0, sig, sz, off, lits = b.soa()
    n = sz.size
    if n:
        # clause multiset as a sorted array of (size, status|added, lits...) rows, ragged -> use tuples via lexsort keys
        order = np.lexsort((np.arange(n),))  # placeholder to keep numpy happy for n == 0
    rows = []
    lits_l = lits.tolist()
    offl = off.tolist()
    w0l = w0.tolist()
    for i in range(n):
        rows.append((tuple(lits_l[offl[i]:offl[i + 1]]), w0l[i] & 0x3, (w0l[i] >> 3) & 1))
    rows.sort()
    return (b.eliminated().tolist(), rows, b.model.tolist(), sorted(b.trail.tolist()))


def diff(a: B.Boundary, r: B.Boundary, strict=True):
    """Returns a list of human-readable mismatches between `a` (ours) and `r` (reference)."""
    out = []
    if a.status != r.status:
        return [f"status {a.status} != {r.status}"]
    if r.status != 0:
        return out
    for k in ("nvars", "n_clauses", "n_literals", "unassigned", "max_frozen", "max_melted", "propagated"):
        if getattr(a, k) != getattr(r, k):
            out.append(f"{k}: {getattr(a, k)} != {getattr(r, k)}")
    if not np.array_equal(a.eliminated(), r.eliminated()):
        out.append(f"eliminated set differs: {a.eliminated().size} vs {r.eliminated().size}")
    if not np.array_equal(a.model, r.model):
        out.append(f"model stack differs: {a.model.size} vs {r.model.size} words")
    if not np.array_equal(a.state & 7, r.state & 7):
        out.append("state[] differs")
    if not np.array_equal(a.value, r.value):
        out.append("value[] differs")
    if not np.array_equal(np.sort(a.trail), np.sort(r.trail)):
        out.append("trail set differs")
    elif strict and not np.array_equal(a.trail, r.trail):
        out.append("trail order differs")
    same_arena = a.arena.size == r.arena.size and np.array_equal(a.arena, r.arena) and np.array_equal(a.refs, r.refs)
    if not same_arena:
        ca, cr = canonical(a), canonical(r)
        if ca[1] != cr[1]:
            sa, sr = set(ca[1]), set(cr[1])
            out.append(f"clause multiset differs: {len(ca[1])} vs {len(cr[1])}; ours-only {list(sa - sr)[:3]} ref-only {list(sr - sa)[:3]}")
        elif strict:
            out.append("arena/refs differ (clause order, flags or signatures) although the clause multiset matches")
    return out
