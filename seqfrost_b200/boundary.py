"""The drop-in boundary of the SIGmA pass as plain numpy arrays.

The reference has no plugin/FFI interface; the boundary this repo replaces is the cut through
``Solver::simplifying()`` right after ``awaken()`` (reference simplify.cpp:171) and right after
``shrinkSimp()`` (simplify.cpp:210).  At both cuts the formula is an ``SCNF``: a bump arena of
uint32 buckets holding ``SCLAUSE`` records (sclause.hpp:45-49: word0 bitfield
``st:2 f:1 a:1 u:2 lbd:26``, word1 ``sig``, word2 ``size``, then the sorted literals) plus the
``refs`` vector of bucket indices in allocation order (simptypes.hpp:48-95).

``Boundary`` carries exactly that plus the solver state the pass reads/writes (state bytes,
literal values, trail, vorg, model stack).  ``read``/``write`` use the sectioned file format that
``oracle/ref_harness.cpp`` emits, so inputs/outputs of the real reference, of the C oracle port
and of the CUDA path are interchangeable.
"""
from __future__ import annotations

import dataclasses
import struct

import numpy as np

MAGIC = b"SGMA0001"

# clause status / flag bits in SCLAUSE word0 (constants.hpp:69-71, sclause.hpp:45)
ST_MASK, ST_LEARNT, ST_DELETED = 0x3, 0x1, 0x2
F_MOLTEN, F_ADDED = 0x4, 0x8
# per-variable state bits (constants.hpp:56-58)
MELTED_M, FROZEN_M, SUBSTITUTED_M = 1, 2, 4

OPT_NAMES = [
    "phases", "phase_lits_min", "mu_pos", "mu_neg", "lcve_max_occurs", "lcve_clause_max",
    "lcve_min_vars", "sub_max_occurs", "sub_clause_max", "ve_clause_max", "xor_max_arity",
    "bce_max_occurs", "ere_max_occurs", "ere_clause_max", "collect_freq", "lbd_tier1",
    "sub_en", "ve_en", "ve_plus_en", "ve_fun_en", "ve_lbound_en", "bce_en", "ere_en",
    "ere_extend", "all_en", "proof_en",
]
# reference defaults (options.cpp:24-51, :146) with ERE off: the parity configuration
# (`-no-redundancy`) until ERE is built.
DEFAULT_OPTS = dict(
    phases=5, phase_lits_min=500, mu_pos=32, mu_neg=32, lcve_max_occurs=3000,
    lcve_clause_max=30000, lcve_min_vars=1, sub_max_occurs=1000, sub_clause_max=100,
    ve_clause_max=100, xor_max_arity=10, bce_max_occurs=1000, ere_max_occurs=1000,
    ere_clause_max=200, collect_freq=2, lbd_tier1=2, sub_en=1, ve_en=1, ve_plus_en=1,
    ve_fun_en=1, ve_lbound_en=0, bce_en=0, ere_en=0, ere_extend=1, all_en=0, proof_en=0,
)


@dataclasses.dataclass
class Boundary:
    status: int = 0                      # 0 OK, 1 UNSAT (reference: killSolver inside the pass)
    nvars: int = 0                       # inf.maxVar
    n_clauses: int = 0                   # inf.nClauses
    n_literals: int = 0                  # inf.nLiterals
    unassigned: int = 0
    max_frozen: int = 0
    max_melted: int = 0
    propagated: int = 0
    simplify_calls: int = 1
    pass_us: int = 0
    arena: np.ndarray = None             # uint32 buckets
    refs: np.ndarray = None              # uint64 bucket index per clause
    state: np.ndarray = None             # uint8[nvars+1]
    value: np.ndarray = None             # int8[2*nvars+2]
    trail: np.ndarray = None             # uint32
    vorg: np.ndarray = None              # uint32[nvars+1]
    model: np.ndarray = None             # uint32 (MODEL::resolved)
    ifrozen: np.ndarray = None           # uint8[nvars+1] or None: incremental freeze mask (solve.hpp:789,816; lcve.cpp:67)
    opts: dict = None
    stats: np.ndarray = None

    # ---- canonical views used by the parity tests ------------------------------------
    def clauses(self):
        """List of (status, added, tuple(lits)) in refs order."""
        a = self.arena
        out = []
        for r in self.refs.tolist():
            w0, sz = int(a[r]), int(a[r + 2])
            out.append((w0 & ST_MASK, (w0 >> 3) & 1, tuple(a[r + 3:r + 3 + sz].tolist())))
        return out

    def clause_multiset(self):
        """Canonically sorted clause multiset: sorted literals + status + added flag."""
        return sorted((tuple(sorted(l)), st, ad) for st, ad, l in self.clauses())

    def eliminated(self):
        return np.nonzero(self.state & MELTED_M)[0]

    def soa(self):
        """(word0, sig, size, off, lits) structure-of-arrays copy of the store, refs order."""
        a, r = self.arena, self.refs.astype(np.int64)
        w0, sig, sz = a[r], a[r + 1], a[r + 2].astype(np.int64)
        off = np.concatenate([[0], np.cumsum(sz)]).astype(np.int64)
        idx = np.repeat(r + 3 - off[:-1], sz) + np.arange(int(off[-1]), dtype=np.int64)
        return w0, sig, sz, off, a[idx]


def read(path) -> Boundary:
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:8] != MAGIC:
        raise ValueError(f"{path}: not a sigma boundary file")
    pos, sec = 8, {}
    while pos < len(blob):
        name = blob[pos:pos + 8].rstrip(b"\0").decode()
        (n,) = struct.unpack_from("<Q", blob, pos + 8)
        sec[name] = blob[pos + 16:pos + 16 + n]
        pos += 16 + n
    meta = np.frombuffer(sec["meta"], dtype=np.uint64)
    b = Boundary(status=int(meta[0]))
    if b.status != 0 and "arena" not in sec:
        return b
    b.nvars, b.n_clauses, b.n_literals = int(meta[1]), int(meta[2]), int(meta[3])
    b.unassigned, b.max_frozen, b.max_melted = int(meta[4]), int(meta[5]), int(meta[6])
    b.propagated, b.simplify_calls, b.pass_us = int(meta[8]), int(meta[9]), int(meta[10])
    b.arena = np.frombuffer(sec["arena"], dtype=np.uint32).copy()
    b.refs = np.frombuffer(sec["refs"], dtype=np.uint64).copy()
    b.state = np.frombuffer(sec["state"], dtype=np.uint8).copy()
    b.value = np.frombuffer(sec["value"], dtype=np.int8).copy()
    b.trail = np.frombuffer(sec["trail"], dtype=np.uint32).copy()
    b.vorg = np.frombuffer(sec["vorg"], dtype=np.uint32).copy()
    b.model = np.frombuffer(sec["model"], dtype=np.uint32).copy()
    if "opts" in sec:
        o = np.frombuffer(sec["opts"], dtype=np.int32)
        b.opts = {k: int(o[i]) for i, k in enumerate(OPT_NAMES)}
    if "stats" in sec:
        b.stats = np.frombuffer(sec["stats"], dtype=np.int64).copy()
    if "ifrozen" in sec:
        b.ifrozen = np.frombuffer(sec["ifrozen"], dtype=np.uint8).copy()
    return b


@dataclasses.dataclass
class ClauseDb:
    """The solver's clause database right before ``awaken()`` (input of ``extract()``, reference
    simplify.cpp:118-153): CMM arena of CLAUSE records (clause.hpp:47-59), the ``orgs`` / ``learnts`` ref
    lists and the deleted stencil (solvetypes.hpp:57-81)."""
    cm: np.ndarray = None                # uint32 buckets
    orgs: np.ndarray = None              # uint64 C_REF
    learnts: np.ndarray = None           # uint64 C_REF
    stencil: np.ndarray = None           # uint8, indexed by C_REF
    # only in dumps taken after newBeginning() (ref_harness --cmout)
    lits_original: int = 0
    lits_learnt: int = 0
    subsume_before: np.ndarray = None    # sp->state[v].subsume before / after newBeginning()
    subsume_after: np.ndarray = None


def read_clause_db(path) -> ClauseDb:
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:8] != MAGIC:
        raise ValueError(f"{path}: not a sigma boundary file")
    pos, sec = 8, {}
    while pos < len(blob):
        name = blob[pos:pos + 8].rstrip(b"\0").decode()
        (n,) = struct.unpack_from("<Q", blob, pos + 8)
        sec[name] = blob[pos + 16:pos + 16 + n]
        pos += 16 + n
    db = ClauseDb(cm=np.frombuffer(sec["cm"], dtype=np.uint32).copy(), orgs=np.frombuffer(sec["orgs"], dtype=np.uint64).copy(),
                  learnts=np.frombuffer(sec["learnts"], dtype=np.uint64).copy())
    if "stencil" in sec:
        db.stencil = np.frombuffer(sec["stencil"], dtype=np.uint8).copy()
    if "cmmeta" in sec:
        m = np.frombuffer(sec["cmmeta"], dtype=np.uint64)
        db.lits_original, db.lits_learnt = int(m[0]), int(m[1])
        db.subsume_before = np.frombuffer(sec["subsume0"], dtype=np.uint8).copy()
        db.subsume_after = np.frombuffer(sec["subsume1"], dtype=np.uint8).copy()
    return db


def clause_db_from_boundary(b: "Boundary", seed=0, shuffle_lits=True, dead_frac=0.0) -> ClauseDb:
    """Synthesises a clause database whose extract() is exactly `b`'s store: every clause of `b` becomes a CLAUSE
    record (literals shuffled, since the solver keeps watches first, not sorted order), originals listed before
    learnts, plus `dead_frac` extra records flagged deleted in the stencil.  Test / bench helper."""
    rng = np.random.default_rng(seed)
    w0, sig, sz, off, lits = b.soa()
    n = int(sz.size)
    learnt = (w0 & ST_MASK) == ST_LEARNT
    assert not np.any(learnt[:-1] & ~learnt[1:]), "extract() lists originals before learnts"
    n_dead = int(n * dead_frac)
    total = n + n_dead
    is_dead = np.zeros(total, dtype=bool)                   # dead records at random places of the arena
    if n_dead:
        is_dead[rng.choice(total, n_dead, replace=False)] = True
    src = np.empty(total, dtype=np.int64)                   # which clause of `b` a record copies
    src[~is_dead] = np.arange(n)
    src[is_dead] = rng.integers(0, max(n, 1), n_dead)
    rsz = sz[src]
    ref = np.concatenate([[0], np.cumsum(rsz + 4)]).astype(np.int64)
    cm = np.zeros(int(ref[-1]), dtype=np.uint32)
    ref = ref[:-1]
    lrn = learnt[src]
    flags = np.uint32(1 << 4) | np.where(rsz == 2, np.uint32(1 << 5), np.uint32(0))          # _k = 1, _b = (size == 2)
    cm[ref] = flags | (lrn.astype(np.uint32) << 8) | ((((w0[src] >> 4) & 3) * lrn).astype(np.uint32) << 24)
    cm[ref + 1] = rsz.astype(np.uint32)
    cm[ref + 2] = 2                                          # _pos
    cm[ref + 3] = ((w0[src] >> 6) * lrn).astype(np.uint32)  # _lbd
    start = np.concatenate([[0], np.cumsum(rsz)])[:-1]
    k = np.arange(int(rsz.sum()), dtype=np.int64)
    gl = lits[np.repeat(off[:-1][src] - start, rsz) + k]
    if shuffle_lits:                                         # random order inside every clause
        perm = np.lexsort((rng.random(gl.size), np.repeat(np.arange(total), rsz)))
        gl = gl[perm]
    cm[np.repeat(ref + 4 - start, rsz) + k] = gl
    stencil = np.zeros(int(ref[-1]) + 1 if total else 0, dtype=np.uint8)
    stencil[ref[is_dead]] = 1
    # dead records also sit in the lists (extract() skips them through cm.deleted())
    in_learnts = np.where(is_dead, rng.random(total) < 0.5, lrn)
    return ClauseDb(cm=cm, orgs=ref[~in_learnts].astype(np.uint64), learnts=ref[in_learnts].astype(np.uint64), stencil=stencil)


def read_trace(path):
    with open(path, "rb") as f:
        blob = f.read()
    pos, sec = 8, {}
    while pos < len(blob):
        name = blob[pos:pos + 8].rstrip(b"\0").decode()
        (n,) = struct.unpack_from("<Q", blob, pos + 8)
        raw = blob[pos + 16:pos + 16 + n]
        sec[name] = np.frombuffer(raw, dtype=np.uint32 if name.startswith("elect") else np.uint64).copy()
        pos += 16 + n
    return sec


def write(path, b: Boundary):
    def section(f, name, arr):
        raw = np.ascontiguousarray(arr).tobytes()
        f.write(name.encode().ljust(8, b"\0"))
        f.write(struct.pack("<Q", len(raw)))
        f.write(raw)
    meta = np.zeros(16, dtype=np.uint64)
    meta[:11] = [b.status, b.nvars, b.n_clauses, b.n_literals, b.unassigned, b.max_frozen,
                 b.max_melted, 0 if b.trail is None else b.trail.size, b.propagated,
                 b.simplify_calls, b.pass_us]
    with open(path, "wb") as f:
        f.write(MAGIC)
        section(f, "meta", meta)
        if b.arena is None:
            return
        for name in ("state", "value", "trail", "vorg", "model", "arena", "refs"):
            section(f, name, getattr(b, name))
        if b.ifrozen is not None:
            section(f, "ifrozen", b.ifrozen)
        if b.opts:
            section(f, "opts", np.array([b.opts.get(k, 0) for k in OPT_NAMES] + [0] * (32 - len(OPT_NAMES)), dtype=np.int32))


def from_cnf(nvars, offsets, lits) -> Boundary:
    """Host mirror of ``awaken()``/``extract()`` (reference simplify.cpp:118-153) for a first call
    on a formula with no units, repeated literals or tautologies: each clause becomes an ORIGINAL
    SCLAUSE with ascending literals (``lit = 2*var + sign``, constants.hpp:77-82) and
    ``sig = OR 1<<(lit&31)`` (sclause.hpp:158-162), allocated back to back in input order."""
    offsets = np.asarray(offsets, dtype=np.int64)
    lits = np.asarray(lits)
    n = offsets.size - 1
    sizes = np.diff(offsets)
    enc = (2 * np.abs(lits).astype(np.uint32) + (lits < 0)).astype(np.uint32)
    refs = (offsets[:-1] + 3 * np.arange(n, dtype=np.int64)).astype(np.uint64)
    arena = np.zeros(int(offsets[-1]) + 3 * n, dtype=np.uint32)
    r = refs.astype(np.int64)
    arena[r + 2] = sizes.astype(np.uint32)
    # one vectorised row sort per distinct clause size (ascending literals inside every clause)
    for k in np.unique(sizes).tolist():
        idx = np.nonzero(sizes == k)[0]
        rows = enc[offsets[idx][:, None] + np.arange(k, dtype=np.int64)[None, :]]
        rows.sort(axis=1)
        arena[r[idx] + 1] = np.bitwise_or.reduce(np.uint32(1) << (rows & np.uint32(31)), axis=1).astype(np.uint32)
        arena[(r[idx] + 3)[:, None] + np.arange(k, dtype=np.int64)[None, :]] = rows
    return Boundary(
        status=0, nvars=int(nvars), n_clauses=n, n_literals=int(offsets[-1]), unassigned=int(nvars),
        max_frozen=0, max_melted=0, propagated=0, simplify_calls=1, arena=arena, refs=refs,
        state=np.zeros(nvars + 1, dtype=np.uint8), value=np.full(2 * nvars + 2, -1, dtype=np.int8),
        trail=np.zeros(0, dtype=np.uint32), vorg=np.arange(nvars + 1, dtype=np.uint32),
        model=np.zeros(0, dtype=np.uint32), opts=dict(DEFAULT_OPTS),
    )
