"""ctypes wrapper of oracle/libsigma_oracle.so (the plain-C restatement).  TEST USE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from seqfrost_b200 import boundary as B
from seqfrost_b200.sigma import SigmaFormula, SigmaOpts, SigmaReport, Sigma, _ptr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_LIB = os.path.join(ROOT, "oracle", "libsigma_oracle.so")


def run(b: B.Boundary, opts=None):
    """Returns (Boundary out, report dict); status 1 = UNSAT."""
    lib = C.CDLL(PORT_LIB)
    ctx = C.c_void_p()
    lib.sgo_create(C.byref(ctx))
    o = Sigma.make_opts(opts if opts is not None else b.opts)
    f = Sigma.make_formula(b)
    lib.sgo_upload.argtypes = [C.c_void_p, C.POINTER(SigmaOpts), C.POINTER(SigmaFormula)]
    lib.sgo_execute.argtypes = [C.c_void_p]
    lib.sgo_result_sizes.argtypes = [C.c_void_p, C.POINTER(SigmaFormula)]
    lib.sgo_download.argtypes = [C.c_void_p, C.POINTER(SigmaFormula)]
    lib.sgo_get_report.argtypes = [C.c_void_p, C.POINTER(SigmaReport)]
    lib.sgo_destroy.argtypes = [C.c_void_p]
    rc = lib.sgo_upload(ctx, C.byref(o), C.byref(f))
    assert rc == 0, rc
    rc = lib.sgo_execute(ctx)
    rep = SigmaReport()
    lib.sgo_get_report(ctx, C.byref(rep))
    if rc == 1:
        lib.sgo_destroy(ctx)
        return B.Boundary(status=1), rep.as_dict()
    assert rc == 0, rc
    sz = SigmaFormula()
    lib.sgo_result_sizes(ctx, C.byref(sz))
    out = B.Boundary(status=0, nvars=sz.nvars, simplify_calls=sz.simplify_calls, vorg=b.vorg, opts=b.opts)
    out.arena = np.empty(sz.n_buckets, np.uint32)
    out.refs = np.empty(sz.n_clauses, np.uint64)
    out.state = np.empty(sz.nvars + 1, np.uint8)
    out.value = np.empty(2 * sz.nvars + 2, np.int8)
    out.trail = np.empty(sz.trail_size, np.uint32)
    out.model = np.empty(sz.model_size, np.uint32)
    g = SigmaFormula()
    g.arena, g.refs, g.state, g.value, g.trail, g.model = (_ptr(out.arena), _ptr(out.refs), _ptr(out.state),
                                                            _ptr(out.value), _ptr(out.trail), _ptr(out.model))
    g.arena_cap, g.refs_cap, g.trail_cap, g.model_cap = sz.n_buckets, sz.n_clauses, sz.trail_size, sz.model_size
    rc = lib.sgo_download(ctx, C.byref(g))
    assert rc == 0, rc
    out.n_clauses, out.n_literals = g.n_clauses, g.n_literals
    out.unassigned, out.max_frozen, out.max_melted, out.propagated = g.unassigned, g.max_frozen, g.max_melted, g.propagated
    lib.sgo_destroy(ctx)
    return out, rep.as_dict()


def extract(db: B.ClauseDb):
    """oracle restatement of extract(orgs); extract(learnts): returns (arena, refs)."""
    from seqfrost_b200.sigma import SigmaCnf
    lib = C.CDLL(PORT_LIB)
    cnf = SigmaCnf()
    cnf.cm, cnf.cm_buckets = _ptr(db.cm), int(db.cm.size)
    cnf.orgs, cnf.n_orgs = _ptr(db.orgs), int(db.orgs.size)
    cnf.learnts, cnf.n_learnts = _ptr(db.learnts), int(db.learnts.size)
    cnf.stencil, cnf.stencil_size = _ptr(db.stencil), int(db.stencil.size)
    arena = np.zeros(int(db.cm.size) + 16, np.uint32)
    refs = np.zeros(int(db.orgs.size + db.learnts.size) + 1, np.uint64)
    nb = C.c_uint64()
    lib.sgo_extract.restype = C.c_uint64
    lib.sgo_extract.argtypes = [C.POINTER(SigmaCnf), C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    n = lib.sgo_extract(C.byref(cnf), _ptr(arena), _ptr(refs), C.byref(nb))
    return arena[:nb.value].copy(), refs[:n].copy()


def new_beginning(b: B.Boundary, lbd_tier1=2):
    """oracle restatement of newBeginning() without mapping: returns a ClauseDb (subsume_after = mark_ssubsume bytes)."""
    from seqfrost_b200.sigma import SigmaCnfOut
    lib = C.CDLL(PORT_LIB)
    n = int(b.refs.size)
    db = B.ClauseDb(cm=np.zeros(int(b.arena.size) + n + 16, np.uint32), orgs=np.zeros(n + 1, np.uint64), learnts=np.zeros(n + 1, np.uint64))
    db.subsume_after = np.zeros(b.nvars + 1, np.uint8)
    co = SigmaCnfOut()
    co.cm, co.orgs, co.learnts, co.subsume = _ptr(db.cm), _ptr(db.orgs), _ptr(db.learnts), _ptr(db.subsume_after)
    lib.sgo_new_beginning.restype = None
    lib.sgo_new_beginning.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.POINTER(SigmaCnfOut)]
    lib.sgo_new_beginning(_ptr(b.arena), _ptr(b.refs), n, lbd_tier1, int(b.simplify_calls > 1), C.byref(co))
    db.cm, db.orgs, db.learnts = db.cm[:co.cm_buckets].copy(), db.orgs[:co.n_orgs].copy(), db.learnts[:co.n_learnts].copy()
    db.lits_original, db.lits_learnt = int(co.lits_original), int(co.lits_learnt)
    return db
