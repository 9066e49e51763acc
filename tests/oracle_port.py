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
