"""ctypes binding of libsigma_b200.so (include/sigma_b200.h) and the host-side mirror of the
reference's pass entry point.

``Sigma.simplify(boundary)`` is the drop-in for the replaced region of
``Solver::simplifying()`` (reference simplify.cpp:180-210): it takes the formula + solver state
exactly as they are right after ``awaken()`` and returns them exactly as they are right after
``shrinkSimp()``, or raises ``Unsat`` where the reference would ``learnEmpty(); killSolver()``.

There is no CPU fallback: if the CUDA library is missing or no device is usable this module
raises, it never computes anything itself.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import boundary as B

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "csrc", "libsigma_b200.so")

SIGMA_OK, SIGMA_UNSAT, SIGMA_INTERRUPTED, SIGMA_E_NOMEM, SIGMA_E_CUDA, SIGMA_E_NOSPACE, SIGMA_E_ARG, SIGMA_E_UNSUPPORTED = range(8)
STAGES = ["h2d", "ingest", "ot_hist", "ot_scan", "ot_scatter", "ot_order", "prop", "vo", "elect", "sot", "sub", "ve",
          "count", "gc", "export", "d2h", "total_device", "ere", "ot_update", "nbh"]
T_N = 20
# stages whose work is HBM streaming (graded against the roofline) and, of those, the ones that are ONE kernel
# per launch (stage time / stage launches = that kernel's launch duration)
STREAMING_STAGES = ("ingest", "ot_hist", "ot_scan", "ot_scatter", "ot_order", "ot_update", "vo", "count", "gc", "export")
SINGLE_KERNEL_STAGES = {"ingest": "k_ingest_fused", "ot_hist": "k_otb_count", "ot_scatter": "k_otb_scatter", "ot_order": "k_otb_lists",
                        "count": "k_count_live", "export": "k_export_staged"}

EXPORTS = [
    "sigma_create", "sigma_destroy", "sigma_default_opts", "sigma_strerror", "sigma_last_error", "sigma_run",
    "sigma_upload", "sigma_execute", "sigma_result_sizes", "sigma_download", "sigma_get_report", "sigma_host_alloc",
    "sigma_host_free", "sigma_host_register", "sigma_host_unregister", "sigma_set_interrupt_flag", "sigma_set_stop", "sigma_snapshot", "sigma_upload_cnf", "sigma_extracted_sizes", "sigma_cnf_out_sizes", "sigma_download_cnf",
    "sigma_parse_dimacs", "sigma_parse_download", "sigma_upload_parsed", "sigma_watches_sizes", "sigma_download_watches",
    "sigma_set_var_map", "sigma_download_state",
]


class SigmaOpts(C.Structure):
    _fields_ = [(n, C.c_int32) for n in B.OPT_NAMES] + [("reserved", C.c_int32 * 6)]


class SigmaFormula(C.Structure):
    _fields_ = [
        ("nvars", C.c_uint32), ("n_clauses", C.c_uint32), ("n_buckets", C.c_uint64), ("n_literals", C.c_uint64),
        ("arena", C.c_void_p), ("refs", C.c_void_p), ("state", C.c_void_p), ("value", C.c_void_p),
        ("trail", C.c_void_p), ("trail_size", C.c_uint32), ("propagated", C.c_uint32), ("vorg", C.c_void_p),
        ("ifrozen", C.c_void_p), ("model", C.c_void_p), ("model_size", C.c_uint64),
        ("unassigned", C.c_uint32), ("max_frozen", C.c_uint32), ("max_melted", C.c_uint32), ("simplify_calls", C.c_uint32),
        ("arena_cap", C.c_uint64), ("refs_cap", C.c_uint32), ("trail_cap", C.c_uint32), ("model_cap", C.c_uint64),
    ]


class SigmaCnf(C.Structure):
    """sigma_cnf (include/sigma_b200.h)."""
    _fields_ = [("cm", C.c_void_p), ("cm_buckets", C.c_uint64), ("orgs", C.c_void_p), ("n_orgs", C.c_uint64),
                ("learnts", C.c_void_p), ("n_learnts", C.c_uint64), ("stencil", C.c_void_p), ("stencil_size", C.c_uint64)]


class SigmaCnfOut(C.Structure):
    """sigma_cnf_out (include/sigma_b200.h)."""
    _fields_ = [("cm", C.c_void_p), ("cm_cap", C.c_uint64), ("orgs", C.c_void_p), ("orgs_cap", C.c_uint64),
                ("learnts", C.c_void_p), ("learnts_cap", C.c_uint64), ("subsume", C.c_void_p),
                ("cm_buckets", C.c_uint64), ("n_orgs", C.c_uint64), ("n_learnts", C.c_uint64),
                ("lits_original", C.c_uint64), ("lits_learnt", C.c_uint64), ("last_ref", C.c_uint64)]


class SigmaDimacs(C.Structure):
    """sigma_dimacs (include/sigma_b200.h)."""
    _fields_ = [("cm", C.c_void_p), ("cm_cap", C.c_uint64), ("orgs", C.c_void_p), ("orgs_cap", C.c_uint64),
                ("trail", C.c_void_p), ("trail_cap", C.c_uint64), ("value", C.c_void_p), ("state", C.c_void_p),
                ("nvars", C.c_uint32), ("header_clauses", C.c_uint32), ("cm_buckets", C.c_uint64), ("n_orgs", C.c_uint64),
                ("n_units", C.c_uint64), ("binaries", C.c_uint64), ("ternaries", C.c_uint64), ("large", C.c_uint64),
                ("literals", C.c_uint64), ("max_clause_size", C.c_uint32), ("unit_rounds", C.c_uint32),
                ("body_offset", C.c_uint64), ("h2d_ms", C.c_float), ("device_ms", C.c_float)]


class SigmaWatches(C.Structure):
    """sigma_watches (include/sigma_b200.h)."""
    _fields_ = [("offs", C.c_void_p), ("offs_cap", C.c_uint64), ("records", C.c_void_p), ("records_cap", C.c_uint64),
                ("n_records", C.c_uint64), ("nlit", C.c_uint32)]


class SigmaReport(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("phases_run", C.c_int32), ("n_clauses", C.c_uint32), ("max_melted_delta", C.c_uint32),
        ("n_literals", C.c_uint64), ("literals_in", C.c_uint64),
        ("pures", C.c_int64), ("inverters", C.c_int64), ("andors", C.c_int64), ("ites", C.c_int64), ("xors", C.c_int64),
        ("funs", C.c_int64), ("resolutions", C.c_int64), ("subsumed", C.c_int64), ("strengthened", C.c_int64),
        ("units", C.c_int64), ("sort_ties_gt24", C.c_int64),
        ("stage_ms", C.c_float * T_N), ("stage_launches", C.c_uint32 * T_N), ("stage_bytes", C.c_uint64 * T_N),
        ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("elected_per_phase", C.c_uint32 * 8), ("election_rounds", C.c_uint32 * 8), ("reserved", C.c_uint64 * 8),
        ("ot_build_mode", C.c_uint32), ("reserved32", C.c_uint32 * 15),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k in ("status", "phases_run", "n_clauses", "max_melted_delta", "n_literals", "literals_in",
                                            "pures", "inverters", "andors", "ites", "xors", "funs", "resolutions", "subsumed",
                                            "strengthened", "units", "sort_ties_gt24", "kernel_launches", "h2d_bytes", "d2h_bytes")}
        d["stage_ms"] = {s: float(self.stage_ms[i]) for i, s in enumerate(STAGES)}
        d["stage_launches"] = {s: int(self.stage_launches[i]) for i, s in enumerate(STAGES)}
        d["stage_bytes"] = {s: int(self.stage_bytes[i]) for i, s in enumerate(STAGES)}
        d["elected_per_phase"] = list(self.elected_per_phase)
        d["election_rounds"] = list(self.election_rounds)
        d["ere"] = dict(zip(("events", "dirty_vars", "full_pairs", "event_pairs"), list(self.reserved)[:4]))
        d["host_round_trips"] = int(self.reserved[4])
        d["extract_us"], d["extract_bytes"] = int(self.reserved[5]), int(self.reserved[6])
        d["ot_build_mode"] = int(self.ot_build_mode)
        return d


class Unsat(Exception):
    """The pass derived the empty clause (reference: learnEmpty(); killSolver(), control.cpp:221)."""


class SigmaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sigma error {code}: {msg}")
        self.code = code


def load(path=None):
    path = path or os.environ.get("SIGMA_B200_LIB") or DEFAULT_LIB
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not found: build it with `python __graft_entry__.py build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.sigma_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.sigma_destroy.argtypes = [C.c_void_p]
    lib.sigma_default_opts.argtypes = [C.POINTER(SigmaOpts)]
    lib.sigma_strerror.restype = C.c_char_p
    lib.sigma_strerror.argtypes = [C.c_int]
    lib.sigma_last_error.restype = C.c_char_p
    lib.sigma_last_error.argtypes = [C.c_void_p]
    lib.sigma_run.argtypes = [C.c_void_p, C.POINTER(SigmaOpts), C.POINTER(SigmaFormula), C.POINTER(SigmaFormula), C.POINTER(SigmaReport)]
    lib.sigma_upload.argtypes = [C.c_void_p, C.POINTER(SigmaOpts), C.POINTER(SigmaFormula)]
    lib.sigma_upload_cnf.argtypes = [C.c_void_p, C.POINTER(SigmaOpts), C.POINTER(SigmaCnf), C.POINTER(SigmaFormula)]
    lib.sigma_extracted_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.sigma_cnf_out_sizes.argtypes = [C.c_void_p, C.POINTER(SigmaCnfOut)]
    lib.sigma_download_cnf.argtypes = [C.c_void_p, C.POINTER(SigmaCnfOut), C.POINTER(SigmaFormula)]
    lib.sigma_parse_dimacs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(SigmaDimacs)]
    lib.sigma_parse_download.argtypes = [C.c_void_p, C.POINTER(SigmaDimacs)]
    lib.sigma_upload_parsed.argtypes = [C.c_void_p, C.POINTER(SigmaOpts)]
    lib.sigma_set_var_map.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.sigma_download_state.argtypes = [C.c_void_p, C.POINTER(SigmaFormula)]
    lib.sigma_watches_sizes.argtypes = [C.c_void_p, C.c_int, C.POINTER(SigmaWatches)]
    lib.sigma_download_watches.argtypes = [C.c_void_p, C.POINTER(SigmaWatches)]
    lib.sigma_execute.argtypes = [C.c_void_p]
    lib.sigma_result_sizes.argtypes = [C.c_void_p, C.POINTER(SigmaFormula)]
    lib.sigma_download.argtypes = [C.c_void_p, C.POINTER(SigmaFormula)]
    lib.sigma_get_report.argtypes = [C.c_void_p, C.POINTER(SigmaReport)]
    lib.sigma_host_alloc.restype = C.c_void_p
    lib.sigma_host_alloc.argtypes = [C.c_uint64]
    lib.sigma_host_free.argtypes = [C.c_void_p]
    lib.sigma_set_stop.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.sigma_set_interrupt_flag.argtypes = [C.c_void_p, C.c_void_p]
    lib.sigma_set_interrupt_flag.restype = None
    lib.sigma_snapshot.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class PinnedArray:
    """numpy view over pinned host memory from sigma_host_alloc (freed with the object)."""

    def __init__(self, lib, dtype, n):
        self.lib, self.nbytes = lib, max(1, int(n) * np.dtype(dtype).itemsize)
        self.ptr = lib.sigma_host_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError("sigma_host_alloc failed")
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(n))

    def __del__(self):
        try:
            if self.ptr:
                self.array = None
                self.lib.sigma_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Sigma:
    """One context per GPU (not thread-safe; different contexts are independent)."""

    def __init__(self, device=0, lib_path=None):
        self.lib = load(lib_path)
        self.ctx = C.c_void_p()
        rc = self.lib.sigma_create(device, C.byref(self.ctx))
        if rc != SIGMA_OK:
            msg = self.lib.sigma_last_error(self.ctx).decode() if self.ctx else "no context"
            if self.ctx:
                self.lib.sigma_destroy(self.ctx)
                self.ctx = None
            raise SigmaError(rc, msg)
        self._keep = None

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.sigma_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc == SIGMA_UNSAT:
            raise Unsat()
        if rc != SIGMA_OK:
            raise SigmaError(rc, self.lib.sigma_last_error(self.ctx).decode())

    @staticmethod
    def make_opts(opts=None):
        o = SigmaOpts()
        d = dict(B.DEFAULT_OPTS)
        d.update(opts or {})
        for k in B.OPT_NAMES:
            setattr(o, k, int(d[k]))
        return o

    @staticmethod
    def make_formula(b: B.Boundary):
        f = SigmaFormula()
        f.nvars, f.n_clauses = b.nvars, int(b.refs.size)
        f.n_buckets = int(b.arena.size)
        f.n_literals = int(b.n_literals) if b.n_literals else int(b.arena.size) - 3 * int(b.refs.size)
        f.arena, f.refs, f.state, f.value = _ptr(b.arena), _ptr(b.refs), _ptr(b.state), _ptr(b.value)
        f.trail, f.trail_size, f.propagated = _ptr(b.trail), int(b.trail.size), b.propagated
        f.vorg, f.ifrozen = _ptr(b.vorg), _ptr(getattr(b, "ifrozen", None))     # incremental freeze mask (lcve.cpp:67)
        f.model, f.model_size = _ptr(b.model), int(b.model.size)
        f.unassigned, f.max_frozen, f.max_melted, f.simplify_calls = b.unassigned, b.max_frozen, b.max_melted, b.simplify_calls
        return f

    # ---- split form -------------------------------------------------------------------
    def upload(self, b: B.Boundary, opts=None):
        self._keep = (b, self.make_opts(opts if opts is not None else b.opts), self.make_formula(b))
        self._check(self.lib.sigma_upload(self.ctx, C.byref(self._keep[1]), C.byref(self._keep[2])))

    def upload_cnf(self, db: B.ClauseDb, b: B.Boundary, opts=None):
        """extract() on the device (reference simplify.cpp:118-153): `db` is the solver's clause database, `b` supplies
        everything else (state, value, trail, vorg, model, counters); b.arena / b.refs are not read."""
        f = self.make_formula(b)
        f.arena, f.refs, f.n_clauses, f.n_buckets, f.n_literals = None, None, 0, 0, 0
        cnf = SigmaCnf()
        cnf.cm, cnf.cm_buckets = _ptr(db.cm), 0 if db.cm is None else int(db.cm.size)
        cnf.orgs, cnf.n_orgs = _ptr(db.orgs), int(db.orgs.size)
        cnf.learnts, cnf.n_learnts = _ptr(db.learnts), int(db.learnts.size)
        cnf.stencil, cnf.stencil_size = _ptr(db.stencil), int(db.stencil.size)
        self._keep = (b, self.make_opts(opts if opts is not None else b.opts), f, cnf, db)
        self._check(self.lib.sigma_upload_cnf(self.ctx, C.byref(self._keep[1]), C.byref(cnf), C.byref(f)))

    def extracted_sizes(self):
        n, nb, nl = C.c_uint32(), C.c_uint64(), C.c_uint64()
        self._check(self.lib.sigma_extracted_sizes(self.ctx, C.byref(n), C.byref(nb), C.byref(nl)))
        return n.value, nb.value, nl.value

    def execute(self):
        self._check(self.lib.sigma_execute(self.ctx))

    def report(self):
        r = SigmaReport()
        self.lib.sigma_get_report(self.ctx, C.byref(r))
        return r.as_dict()

    def download(self, pinned=False) -> B.Boundary:
        sz = SigmaFormula()
        self._check(self.lib.sigma_result_sizes(self.ctx, C.byref(sz)))
        src = self._keep[0]
        out = B.Boundary(status=0, nvars=sz.nvars, simplify_calls=sz.simplify_calls, vorg=src.vorg, opts=src.opts)
        keep = []

        def alloc(dtype, n):
            if pinned:
                p = PinnedArray(self.lib, dtype, n)
                keep.append(p)
                return p.array
            return np.empty(int(n), dtype=dtype)
        out.arena = alloc(np.uint32, sz.n_buckets)
        out.refs = alloc(np.uint64, sz.n_clauses)
        out.state = alloc(np.uint8, sz.nvars + 1)
        out.value = alloc(np.int8, 2 * sz.nvars + 2)
        out.trail = alloc(np.uint32, sz.trail_size)
        out.model = alloc(np.uint32, sz.model_size)
        f = SigmaFormula()
        f.arena, f.refs, f.state, f.value, f.trail, f.model = (_ptr(out.arena), _ptr(out.refs), _ptr(out.state),
                                                                _ptr(out.value), _ptr(out.trail), _ptr(out.model))
        f.arena_cap, f.refs_cap, f.trail_cap, f.model_cap = sz.n_buckets, sz.n_clauses, sz.trail_size, sz.model_size
        self._check(self.lib.sigma_download(self.ctx, C.byref(f)))
        out.n_clauses, out.n_literals = f.n_clauses, f.n_literals
        out.unassigned, out.max_frozen, out.max_melted, out.propagated = f.unassigned, f.max_frozen, f.max_melted, f.propagated
        out._pinned = keep
        return out

    def download_cnf(self):
        """newBeginning() on the device (reference simplify.cpp:249-268): returns (ClauseDb, Boundary without arena / refs)."""
        sz, co = SigmaFormula(), SigmaCnfOut()
        self._check(self.lib.sigma_result_sizes(self.ctx, C.byref(sz)))
        self._check(self.lib.sigma_cnf_out_sizes(self.ctx, C.byref(co)))
        src = self._keep[0]
        out = B.Boundary(status=0, nvars=sz.nvars, simplify_calls=sz.simplify_calls, vorg=src.vorg, opts=src.opts)
        db = B.ClauseDb(cm=np.empty(int(co.cm_buckets), np.uint32), orgs=np.empty(int(co.n_orgs), np.uint64),
                        learnts=np.empty(int(co.n_learnts), np.uint64))
        db.subsume_after = np.empty(sz.nvars + 1, np.uint8)
        out.state, out.value = np.empty(sz.nvars + 1, np.uint8), np.empty(2 * sz.nvars + 2, np.int8)
        out.trail, out.model = np.empty(int(sz.trail_size), np.uint32), np.empty(int(sz.model_size), np.uint32)
        co.cm, co.orgs, co.learnts, co.subsume = _ptr(db.cm), _ptr(db.orgs), _ptr(db.learnts), _ptr(db.subsume_after)
        co.cm_cap, co.orgs_cap, co.learnts_cap = db.cm.size, db.orgs.size, db.learnts.size
        f = SigmaFormula()
        f.state, f.value, f.trail, f.model = _ptr(out.state), _ptr(out.value), _ptr(out.trail), _ptr(out.model)
        f.trail_cap, f.model_cap = sz.trail_size, sz.model_size
        self._check(self.lib.sigma_download_cnf(self.ctx, C.byref(co), C.byref(f)))
        db.lits_original, db.lits_learnt = int(co.lits_original), int(co.lits_learnt)
        out.n_clauses, out.n_literals = f.n_clauses, f.n_literals
        out.unassigned, out.max_frozen, out.max_melted, out.propagated = f.unassigned, f.max_frozen, f.max_melted, f.propagated
        return db, out

    # ---- reusable pinned hand-off buffers (what an integrated solver would own) -----------
    def alloc_pinned_like(self, b: B.Boundary, slack=1.25):
        """Pinned copies of the input arrays plus pinned output buffers with capacity
        `slack` x the input sizes; returns (pinned input Boundary, output buffer dict)."""
        keep = []

        def pin(a, n=None):
            p = PinnedArray(self.lib, a.dtype, a.size if n is None else n)
            keep.append(p)
            if n is None:
                p.array[:] = a
            return p.array
        bi = B.Boundary(**{k: getattr(b, k) for k in ("status", "nvars", "n_clauses", "n_literals", "unassigned", "max_frozen",
                                                       "max_melted", "propagated", "simplify_calls", "opts")})
        bi.arena, bi.refs, bi.state, bi.value = pin(b.arena), pin(b.refs), pin(b.state), pin(b.value)
        bi.trail, bi.vorg, bi.model = pin(b.trail), pin(b.vorg), pin(b.model)
        bi.ifrozen = None if b.ifrozen is None else pin(b.ifrozen)
        bufs = dict(arena=pin(b.arena, int(b.arena.size * slack) + 1024), refs=pin(b.refs, int(b.refs.size * slack) + 1024),
                    state=pin(b.state, b.nvars + 1), value=pin(b.value, 2 * b.nvars + 2),
                    trail=pin(np.zeros(0, np.uint32), b.nvars + 16),
                    model=pin(np.zeros(0, np.uint32), int(b.arena.size) + b.model.size + 1024))
        bi._pinned = keep
        return bi, bufs

    def download_into(self, bufs) -> SigmaFormula:
        """D2H into caller-owned buffers (raises SigmaError(E_NOSPACE) when too small)."""
        f = SigmaFormula()
        f.arena, f.refs, f.state, f.value, f.trail, f.model = (_ptr(bufs["arena"]), _ptr(bufs["refs"]), _ptr(bufs["state"]),
                                                                _ptr(bufs["value"]), _ptr(bufs["trail"]), _ptr(bufs["model"]))
        f.arena_cap, f.refs_cap = bufs["arena"].size, bufs["refs"].size
        f.trail_cap, f.model_cap = bufs["trail"].size, bufs["model"].size
        self._check(self.lib.sigma_download(self.ctx, C.byref(f)))
        return f

    # ---- one-shot: the drop-in for simplify.cpp:180-210 ---------------------------------
    def simplify(self, b: B.Boundary, opts=None):
        """Returns (Boundary after the pass, report dict). Raises Unsat like the reference exits."""
        self.upload(b, opts)
        self.execute()
        out = self.download()
        return out, self.report()

    # ---- the DIMACS parser on the device (reference dimacs.cpp:25-177) ----------------------
    def parse_dimacs(self, text, download=True):
        """`text`: the file image (bytes / numpy uint8 / PinnedArray.array).  Returns a dict with the parsed clause database
        (cm, orgs, trail, value, state as numpy arrays when download=True) and the counters.  Raises Unsat for an empty
        clause, SigmaError(E_ARG) for anything the reference's parser stops on."""
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else text
        d = SigmaDimacs()
        self._check(self.lib.sigma_parse_dimacs(self.ctx, _ptr(buf), C.c_uint64(buf.size), C.byref(d)))
        out = {k: getattr(d, k) for k in ("nvars", "header_clauses", "cm_buckets", "n_orgs", "n_units", "binaries", "ternaries", "large",
                                          "literals", "max_clause_size", "unit_rounds", "body_offset", "h2d_ms", "device_ms")}
        if download:
            cm, orgs, trail = np.empty(d.cm_buckets, np.uint32), np.empty(d.n_orgs, np.uint64), np.empty(d.n_units, np.uint32)
            value, state = np.empty(2 * d.nvars + 2, np.int8), np.empty(d.nvars + 1, np.uint8)
            d.cm, d.cm_cap, d.orgs, d.orgs_cap, d.trail, d.trail_cap = _ptr(cm), cm.size, _ptr(orgs), orgs.size, _ptr(trail), trail.size
            d.value, d.state = _ptr(value), _ptr(state)
            self._check(self.lib.sigma_parse_download(self.ctx, C.byref(d)))
            out.update(cm=cm, orgs=orgs, trail=trail, value=value, state=state)
        return out

    def upload_parsed(self, opts=None, nvars=None):
        """The parsed file (parse_dimacs) becomes the pass input without leaving the device."""
        o = self.make_opts(opts)
        self._check(self.lib.sigma_upload_parsed(self.ctx, C.byref(o)))
        src = B.Boundary(nvars=nvars or 0, vorg=None if nvars is None else np.arange(nvars + 1, dtype=np.uint32), opts=dict(B.DEFAULT_OPTS, **(opts or {})))
        self._keep = (src, o, None)

    def set_var_map(self, vmap, new_nvars):
        """map(true): clause database and watch table in the numbering vmap[old var] -> new var (None switches back)."""
        if vmap is None:
            self._check(self.lib.sigma_set_var_map(self.ctx, None, 0))
        else:
            m = np.ascontiguousarray(vmap, dtype=np.uint32)
            self._check(self.lib.sigma_set_var_map(self.ctx, _ptr(m), int(new_nvars)))

    def download_state(self):
        """state / value / trail / model after execute (no clause arrays)."""
        sz = SigmaFormula()
        self._check(self.lib.sigma_result_sizes(self.ctx, C.byref(sz)))
        out = B.Boundary(status=0, nvars=sz.nvars)
        out.state, out.value = np.empty(sz.nvars + 1, np.uint8), np.empty(2 * sz.nvars + 2, np.int8)
        out.trail, out.model = np.empty(int(sz.trail_size), np.uint32), np.empty(int(sz.model_size), np.uint32)
        f = SigmaFormula()
        f.state, f.value, f.trail, f.model = _ptr(out.state), _ptr(out.value), _ptr(out.trail), _ptr(out.model)
        f.trail_cap, f.model_cap = sz.trail_size, sz.model_size
        self._check(self.lib.sigma_download_state(self.ctx, C.byref(f)))
        return out

    def watches(self, priorbins=1):
        """rebuildWT(priorbins) on the device; call before download_cnf.  Returns (offs uint32[nlit + 1], records uint32[n, 4] =
        {ref lo, ref hi, imp, size})."""
        w = SigmaWatches()
        self._check(self.lib.sigma_watches_sizes(self.ctx, int(priorbins), C.byref(w)))
        offs, rec = np.empty(w.nlit + 1, np.uint32), np.empty((int(w.n_records), 4), np.uint32)
        w.offs, w.offs_cap, w.records, w.records_cap = _ptr(offs), offs.size, _ptr(rec), int(w.n_records)
        self._check(self.lib.sigma_download_watches(self.ctx, C.byref(w)))
        return offs, rec

    def set_interrupt_flag(self, flag):
        """`flag`: a one-byte numpy array the caller may set to non-zero from another thread / a signal handler (the
        reference's `bool interrupted`, solve.hpp:126); polled where the reference polls INTERRUPTED.  None detaches."""
        self._flag = flag
        self.lib.sigma_set_interrupt_flag(self.ctx, _ptr(flag))

    def set_stop(self, phase, stage):
        self._check(self.lib.sigma_set_stop(self.ctx, phase, STAGES.index(stage) if isinstance(stage, str) else stage))

    def snapshot(self, name, dtype):
        n = C.c_uint64()
        self.lib.sigma_snapshot(self.ctx, name.encode(), None, 0, C.byref(n))
        buf = np.empty(n.value // np.dtype(dtype).itemsize, dtype=dtype)
        self._check(self.lib.sigma_snapshot(self.ctx, name.encode(), _ptr(buf), n.value, C.byref(n)))
        return buf
