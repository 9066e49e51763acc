"""Seeded synthetic CNF generators for the SIGmA pass workloads (BASELINE.json configs).

Every generator returns ``(nvars, offsets, lits)``: CSR clause arrays with DIMACS-signed int32
literals, clause ``i`` = ``lits[offsets[i]:offsets[i+1]]``.  Clauses never hold a repeated
variable (no duplicate literal, no tautology) and never are units, so the reference's parser
(dimacs.cpp:121) keeps them verbatim and root-level BCP has nothing to do.

There is no network and no benchmark suite in the reference: these generators ARE the data.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "random_ksat", "planted_solution_ksat", "planted_subsumption_ksat", "gate_circuit", "xor_php_mix",
    "mixed_instance", "write_dimacs", "concat",
]


def _distinct_rows(rng, nrows, k, nvars):
    """nrows x k variable ids in [1, nvars], distinct inside each row."""
    v = rng.integers(1, nvars + 1, size=(nrows, k), dtype=np.int64)
    while True:
        s = np.sort(v, axis=1)
        bad = np.nonzero((s[:, 1:] == s[:, :-1]).any(axis=1))[0]
        if bad.size == 0:
            return v
        v[bad] = rng.integers(1, nvars + 1, size=(bad.size, k), dtype=np.int64)


def _signed(rng, v):
    sign = rng.integers(0, 2, size=v.shape, dtype=np.int64) * 2 - 1
    return (v * sign).astype(np.int32)


def _csr_from_fixed(rows):
    n, k = rows.shape
    return np.arange(0, (n + 1) * k, k, dtype=np.int64), rows.reshape(-1).astype(np.int32)


def concat(parts):
    """Concatenate [(offsets, lits), ...] CSR pieces."""
    offs = [np.zeros(1, dtype=np.int64)]
    lits = []
    base = 0
    for o, l in parts:
        offs.append(o[1:] + base)
        lits.append(l)
        base += int(o[-1])
    return np.concatenate(offs), (np.concatenate(lits) if lits else np.zeros(0, np.int32))


def _ragged(rows_list):
    """rows_list: list of 2-D int arrays (each fixed width) -> CSR, keeping list order."""
    return concat([_csr_from_fixed(r) for r in rows_list if r.shape[0]])


def _shuffle_clauses(rng, offsets, lits):
    n = offsets.size - 1
    perm = rng.permutation(n)
    sizes = np.diff(offsets)
    new_sizes = sizes[perm]
    new_off = np.concatenate([[0], np.cumsum(new_sizes)]).astype(np.int64)
    # gather literals
    starts = offsets[:-1][perm]
    idx = np.repeat(starts - new_off[:-1], new_sizes) + np.arange(int(new_off[-1]), dtype=np.int64)
    return new_off, lits[idx]


def random_ksat(nvars, nclauses, k=3, seed=1):
    """Uniform random k-SAT (configs C1: k=3 ratio 4.26; C2 base: k=5)."""
    rng = np.random.default_rng(seed)
    rows = _signed(rng, _distinct_rows(rng, nclauses, k, nvars))
    o, l = _csr_from_fixed(rows)
    return nvars, o, l


def planted_solution_ksat(nvars, nclauses, k=3, seed=1):
    """Random k-SAT with a hidden satisfying assignment (every clause holds at least one literal the assignment makes
    true): the solvable stand-in for C1's "sigmify + solve" at full size - uniform 3-SAT at ratio 4.26 with 100k
    variables is beyond both binaries, and the reference's -boundsearch option is dead code (options.cpp never copies
    it into opts.boundsearch_en), so the search cannot be bounded deterministically (SURVEY 8d C1 caveat)."""
    rng = np.random.default_rng(seed)
    hidden = rng.integers(0, 2, size=nvars + 1, dtype=np.int64) * 2 - 1          # +1 / -1 per variable
    rows = _signed(rng, _distinct_rows(rng, nclauses, k, nvars)).astype(np.int64)
    while True:
        sat = (np.sign(rows) == hidden[np.abs(rows)]).any(axis=1)
        bad = np.nonzero(~sat)[0]
        if bad.size == 0:
            break
        rows[bad] = _signed(rng, _distinct_rows(rng, bad.size, k, nvars))
    o, l = _csr_from_fixed(rows.astype(np.int32))
    return nvars, o, l


def planted_subsumption_ksat(nvars, nclauses, k=5, seed=7, sup_frac=0.05, self_frac=0.02):
    """C2: random k-SAT plus planted SUB work: for ~5% of the base clauses a
    superset with 1-3 extra literals, for ~2% a self-subsuming partner (same literals, one
    flipped, optional extras). Total clause count == nclauses."""
    rng = np.random.default_rng(seed)
    n_base = int(round(nclauses / (1.0 + sup_frac + self_frac)))
    n_sup = int(round(n_base * sup_frac))
    n_self = nclauses - n_base - n_sup
    base = _signed(rng, _distinct_rows(rng, n_base, k, nvars))
    pieces = [base]
    # supersets: copy of a base clause + e extra distinct variables
    src = rng.integers(0, n_base, size=n_sup)
    extra_n = rng.integers(1, 4, size=n_sup)
    for e in (1, 2, 3):
        sel = src[extra_n == e]
        if sel.size == 0:
            continue
        rows = np.concatenate([base[sel], _signed(rng, _distinct_rows(rng, sel.size, e, nvars))], axis=1)
        ok = _rows_distinct_vars(rows)
        pieces.append(rows[ok])
    # self-subsuming partners: flip one literal, optionally add one extra literal
    src = rng.integers(0, n_base, size=n_self)
    part = base[src].copy()
    flip = rng.integers(0, k, size=n_self)
    part[np.arange(n_self), flip] *= -1
    with_extra = rng.integers(0, 2, size=n_self).astype(bool)
    pieces.append(part[~with_extra])
    rows = np.concatenate([part[with_extra], _signed(rng, _distinct_rows(rng, int(with_extra.sum()), 1, nvars))], axis=1)
    pieces.append(rows[_rows_distinct_vars(rows)])
    o, l = _ragged(pieces)
    # top up to the exact clause count with plain random clauses (rows dropped above)
    missing = nclauses - (o.size - 1)
    if missing > 0:
        o, l = concat([(o, l), _csr_from_fixed(_signed(rng, _distinct_rows(rng, missing, k, nvars)))])
    o, l = _shuffle_clauses(rng, o, l)
    return nvars, o, l


def _rows_distinct_vars(rows):
    s = np.sort(np.abs(rows), axis=1)
    return ~(s[:, 1:] == s[:, :-1]).any(axis=1)


def gate_circuit(nvars, seed=11, window=2000, n_inputs=None, frac=(0.40, 0.25, 0.25, 0.10),
                 out_constraints=0.02):
    """C3: BMC-style unrolled circuit. Variables ``n_inputs+1..nvars`` are gate outputs defined by
    Tseitin clauses over random earlier signals within a sliding window:
      AND  x = a & b        : (-x a)(-x b)(x -a -b)
      XOR  x = a ^ b        : (-x a b)(-x -a -b)(x -a b)(x a -b)
      ITE  x = c ? t : e    : (-x -c t)(-x c e)(x -c -t)(x c -e)
      EQ   x = a            : (-x a)(x -a)
    with random input polarities, plus random ternary output constraints."""
    rng = np.random.default_rng(seed)
    if n_inputs is None:
        n_inputs = max(8, min(window, nvars // 50))
    n_inputs = min(n_inputs, nvars - 1)
    gates = np.arange(n_inputs + 1, nvars + 1, dtype=np.int64)
    ng = gates.size
    kind = rng.choice(4, size=ng, p=np.array(frac) / sum(frac))

    def pick(n_in):
        lo = np.maximum(1, gates - window)
        span = gates - lo  # candidates lo .. gates-1
        ins = lo[:, None] + (rng.random((ng, n_in)) * span[:, None]).astype(np.int64)
        # make inputs distinct where the span allows it
        for _ in range(8):
            s = np.sort(ins, axis=1)
            bad = np.nonzero((s[:, 1:] == s[:, :-1]).any(axis=1) & (span >= n_in))[0]
            if bad.size == 0:
                break
            ins[bad] = lo[bad, None] + (rng.random((bad.size, n_in)) * span[bad, None]).astype(np.int64)
        pol = rng.integers(0, 2, size=(ng, n_in), dtype=np.int64) * 2 - 1
        return ins * pol, ins

    (sig3, raw3) = pick(3)
    s = np.sort(raw3, axis=1)
    dup3 = (s[:, 1:] == s[:, :-1]).any(axis=1)
    dup2 = raw3[:, 0] == raw3[:, 1]
    # degrade gates that could not get distinct inputs (tiny spans at the very start)
    kind = np.where((kind == 2) & dup3, 0, kind)
    kind = np.where(((kind == 0) | (kind == 1)) & dup2, 3, kind)
    x = gates
    a, b, c = sig3[:, 0], sig3[:, 1], sig3[:, 2]
    parts = []
    # keep gate order: build per-gate clause blocks, then interleave by gate index
    blocks = []  # (gate_index array, rows)
    m = kind == 0
    blocks.append((np.nonzero(m)[0], [np.stack([-x[m], a[m]], 1), np.stack([-x[m], b[m]], 1)],
                   [np.stack([x[m], -a[m], -b[m]], 1)]))
    m = kind == 1
    blocks.append((np.nonzero(m)[0], [], [np.stack([-x[m], a[m], b[m]], 1), np.stack([-x[m], -a[m], -b[m]], 1),
                                          np.stack([x[m], -a[m], b[m]], 1), np.stack([x[m], a[m], -b[m]], 1)]))
    m = kind == 2
    blocks.append((np.nonzero(m)[0], [], [np.stack([-x[m], -c[m], a[m]], 1), np.stack([-x[m], c[m], b[m]], 1),
                                          np.stack([x[m], -c[m], -a[m]], 1), np.stack([x[m], c[m], -b[m]], 1)]))
    m = kind == 3
    blocks.append((np.nonzero(m)[0], [np.stack([-x[m], a[m]], 1), np.stack([x[m], -a[m]], 1)], []))
    # assemble (gate idx, slot) keyed ragged clause list
    keys, sizes, flat = [], [], []
    for gi, bins, terns in blocks:
        slot = 0
        for rows in bins:
            keys.append(gi * 8 + slot); sizes.append(np.full(gi.size, 2)); flat.append(rows); slot += 1
        for rows in terns:
            keys.append(gi * 8 + slot); sizes.append(np.full(gi.size, 3)); flat.append(rows); slot += 1
    n_out = int(ng * out_constraints)
    if n_out:
        lo = max(1, nvars - max(window, nvars // 4))
        if nvars - lo + 1 >= 3:
            ov = lo - 1 + _distinct_rows(rng, n_out, 3, nvars - lo + 1)
            keys.append(np.full(n_out, ng * 8) + np.arange(n_out)); sizes.append(np.full(n_out, 3))
            flat.append(_signed(rng, ov).astype(np.int64))
    keys = np.concatenate(keys); sizes = np.concatenate(sizes)
    order = np.argsort(keys, kind="stable")
    # flatten literals of all rows in `flat` order, then permute clause-wise
    offs_in = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    lits_in = np.concatenate([r.reshape(-1) for r in flat]).astype(np.int32)
    new_sizes = sizes[order]
    new_off = np.concatenate([[0], np.cumsum(new_sizes)]).astype(np.int64)
    idx = np.repeat(offs_in[:-1][order] - new_off[:-1], new_sizes) + np.arange(int(new_off[-1]), dtype=np.int64)
    return nvars, new_off, lits_in[idx]


def unit_storm(nseeds=100_000, depth=8, seed=5, noise=2.0, planted=True):
    """Unit-heavy instance for prop() (propelim.cpp:46-81).  Units that are in the file are propagated by the solver before
    the simplifier starts (rootify(), simplify.cpp:160), so the units here are hidden: every seed literal a comes as
    (a b)(a -b), which self-subsumption strengthens to the unit clause a during SUB.  The next prop() then starts with
    about `nseeds` units on the trail and runs down `depth` layers of implications u & w -> y between consecutive
    layers.  All literals follow a hidden assignment, so the instance is satisfiable; random ternary clauses over the
    layer variables (`noise` per variable, at least one literal true under the hidden assignment) give the visits
    clauses to shrink and clauses to delete, and keep most layer variables out of the first election."""
    rng = np.random.default_rng(seed)
    n = int(nseeds)
    nl = n * depth
    nv = 2 * n + nl
    hidden = rng.integers(0, 2, size=nv + 1, dtype=np.int64) * 2 - 1              # +1 / -1 per variable
    a = np.arange(1, n + 1, dtype=np.int64) * hidden[1:n + 1]                        # literals true under `hidden`
    b = np.arange(n + 1, 2 * n + 1, dtype=np.int64)
    rows2 = np.concatenate([np.stack([a, b], 1), np.stack([a, -b], 1)])
    prev = a
    horn = []
    for k in range(depth):
        yv = 2 * n + k * n + np.arange(1, n + 1, dtype=np.int64)
        y = yv * hidden[yv]
        u = prev[rng.integers(0, n, size=n)]
        w = prev[rng.integers(0, n, size=n)]
        keep = u != w
        horn.append(np.stack([-u[keep], -w[keep], y[keep]], 1))
        prev = y
    nn = int(noise * nl)
    rows = _signed(rng, 2 * n + _distinct_rows(rng, nn, 3, nl)).astype(np.int64)
    while planted:
        bad = np.nonzero(~(np.sign(rows) == hidden[np.abs(rows)]).any(axis=1))[0]
        if bad.size == 0:
            break
        rows[bad] = _signed(rng, 2 * n + _distinct_rows(rng, bad.size, 3, nl))
    rows3 = np.concatenate(horn + [rows])
    o, l = _ragged([rows2.astype(np.int32), rows3.astype(np.int32)])
    o, l = _shuffle_clauses(rng, o, l)
    return nv, o, l


def _xor_clauses(vars_, rhs):
    """CNF expansion of XOR(vars_) = rhs: all sign patterns with the wrong parity are forbidden."""
    k = len(vars_)
    rows = []
    for mask in range(1 << k):
        # clause with literal i negated iff bit i of mask set forbids assignment (v_i = bit i)
        if (bin(mask).count("1") & 1) != rhs:
            rows.append([(-v if (mask >> i) & 1 else v) for i, v in enumerate(vars_)])
    return rows


def xor_php_mix(nclauses=1_000_000, seed=3, php_holes=7, php_share=0.25):
    """C4: XOR chains (arity 3-6, CNF-expanded) mixed with pigeonhole PHP(n+1,n) blocks.
    Exercises the pOrgs+nOrgs and ve_clause_max bounds of BVE."""
    rng = np.random.default_rng(seed)
    groups = {}  # size -> list of rows
    nv = 0
    made = 0
    php_target = int(nclauses * php_share)
    n = php_holes
    while made < php_target:
        base = nv
        p = lambda i, j: base + i * n + j + 1  # pigeon i in hole j
        for i in range(n + 1):
            groups.setdefault(n, []).append([p(i, j) for j in range(n)])
        pairs = []
        for j in range(n):
            for i1 in range(n + 1):
                for i2 in range(i1 + 1, n + 1):
                    pairs.append([-p(i1, j), -p(i2, j)])
        groups.setdefault(2, []).extend(pairs)
        made += (n + 1) + len(pairs)
        nv += (n + 1) * n
    # XOR chains: x_i ^ x_{i+1} ^ .. = b, consecutive windows sharing arity-1 variables
    patt = {k: [np.array(_xor_clauses(list(range(1, k + 1)), r), dtype=np.int64) for r in (0, 1)] for k in (3, 4, 5, 6)}
    xor_rows = {k: [] for k in (3, 4, 5, 6)}
    while made < nclauses:
        k = int(rng.integers(3, 7))
        length = int(rng.integers(50, 400))
        start = nv
        nv += length + k - 1
        pos = start + np.arange(length, dtype=np.int64)
        rhs = rng.integers(0, 2, size=length)
        for r in (0, 1):
            sel = pos[rhs == r]
            if sel.size == 0:
                continue
            pt = patt[k][r]  # (2^(k-1), k) over local ids 1..k
            rows = (np.sign(pt)[None, :, :] * (np.abs(pt)[None, :, :] + sel[:, None, None])).reshape(-1, k)
            xor_rows[k].append(rows)
            made += rows.shape[0]
    pieces = []
    for sz, rows in groups.items():
        pieces.append(np.array(rows, dtype=np.int64).reshape(-1, sz))
    for k, lst in xor_rows.items():
        if lst:
            pieces.append(np.concatenate(lst))
    o, l = _ragged(pieces)
    o, l = _shuffle_clauses(rng, o, l)
    if o.size - 1 > nclauses:
        o = o[: nclauses + 1]
        l = l[: int(o[-1])]
    return nv, o, l.astype(np.int32)


def mixed_instance(seed, nclauses=None):
    """C5: one instance of the 64-instance batch, drawn from the C2/C3/C4 generators with
    0.5-5M clauses (seeds 100..163)."""
    rng = np.random.default_rng(seed)
    if nclauses is None:
        nclauses = int(rng.integers(500_000, 5_000_001))
    which = seed % 3
    if which == 0:
        return planted_subsumption_ksat(nclauses // 5, nclauses, k=5, seed=seed)
    if which == 1:
        return gate_circuit(int(nclauses / 3.4), seed=seed)
    return xor_php_mix(nclauses, seed=seed)


def write_dimacs(path, nvars, offsets, lits, chunk=1 << 20):
    """Write a DIMACS file (fast path: one vectorised savetxt per clause-size group, kept in order
    chunk by chunk so the file order == clause order)."""
    n = offsets.size - 1
    sizes = np.diff(offsets)
    with open(path, "w") as f:
        f.write(f"p cnf {nvars} {n}\n")
        for c0 in range(0, n, chunk):
            c1 = min(n, c0 + chunk)
            sz = sizes[c0:c1]
            if sz.size and (sz == sz[0]).all():
                k = int(sz[0])
                rows = lits[offsets[c0]:offsets[c1]].reshape(-1, k)
                np.savetxt(f, np.concatenate([rows, np.zeros((rows.shape[0], 1), rows.dtype)], 1), fmt="%d")
            else:
                seg = lits[offsets[c0]:offsets[c1]]
                out = np.zeros(seg.size + (c1 - c0), dtype=np.int64)
                pos = np.arange(seg.size) + np.repeat(np.arange(c1 - c0), sz)
                out[pos] = seg
                # zeros mark clause ends; print with newline after each zero
                txt = " ".join(map(str, out.tolist())).replace(" 0 ", " 0\n")
                f.write(txt + "\n")


def fuzz_mix(seed):
    """Small mixed instance (8-120 variables): random 2..5-literal clauses, a few Tseitin gates,
    duplicated and subsumed clauses.  Roughly one seed in twelve makes the pass derive units, one
    in twelve makes it derive the empty clause: the edge cases of prop()/strengthen()/addResolvent
    (reference propelim.cpp:46-81, subsume.cpp:45-76, bounded.cpp:279-307)."""
    import random
    rng = random.Random(seed)
    nv = rng.randint(8, 120)
    clauses = []
    kind = rng.random()
    nc = int(nv * rng.uniform(1.5, 4.5))
    for _ in range(nc):
        k = rng.choice([2, 2, 3, 3, 3, 4, 5]) if kind < 0.7 else rng.choice([2, 3])
        k = min(k, nv)
        vs = rng.sample(range(1, nv + 1), k)
        clauses.append([v if rng.random() < 0.5 else -v for v in vs])
    for _ in range(rng.randint(0, nv // 3)):
        if nv < 4:
            break
        x, a, b, c = rng.sample(range(1, nv + 1), 4)
        t = rng.random()
        if t < 0.25:
            clauses += [[-x, a], [-x, b], [x, -a, -b]]
        elif t < 0.5:
            clauses += [[-x, a, b], [-x, -a, -b], [x, -a, b], [x, a, -b]]
        elif t < 0.75:
            clauses += [[-x, -a, b], [-x, a, c], [x, -a, -b], [x, a, -c]]
        else:
            clauses += [[-x, a], [x, -a]]
    for _ in range(rng.randint(0, 5)):
        c = rng.choice(clauses)
        clauses.append(list(c))
        c = rng.choice(clauses)
        extra = [v for v in range(1, nv + 1) if v not in [abs(l) for l in c]]
        if extra:
            clauses.append(c + [rng.choice(extra) * rng.choice([1, -1])])
    rng.shuffle(clauses)
    sizes = np.array([len(c) for c in clauses], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    lits = np.array([l for c in clauses for l in c], dtype=np.int32)
    return nv, off, lits
