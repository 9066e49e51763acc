"""Edge cases the reference's own code paths make special: occurrence lists longer than 24 with tied sort keys
(std::sort branch of sortOT, sort.hpp:64-72), clauses beyond --subsumeclausemax / --boundedclausemax, hub
variables beyond the election caps, formulas the pass empties completely, and the C-ABI error paths.
CPU: serial stand-in of the launch layer; GPU (-m gpu): the CUDA library.  Checker: oracle/_ref/ref_sigma."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

import parity
from conftest import GOLDEN, REF_BIN
from seqfrost_b200 import boundary as B, sigma

needs_ref = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ref_sigma not built")


def _cases():
    rng = random.Random(1)
    out = {}
    cl = [[1, 2 + (r % 3), 50] for r in range(28)] + [[-1, 60 + r, 70 + r] for r in range(3)]
    for _ in range(900):
        vs = rng.sample(range(2, 401), 3)
        cl.append([v if rng.random() < .5 else -v for v in vs])
    out["ties_gt24"] = (400, cl, [])
    cl = []
    for _ in range(40):
        vs = rng.sample(range(1, 601), rng.choice([120, 150, 30, 26]))
        cl.append([v if rng.random() < .5 else -v for v in vs])
    for _ in range(1500):
        vs = rng.sample(range(1, 601), 3)
        cl.append([v if rng.random() < .5 else -v for v in vs])
    out["big_clauses"] = (600, cl, [])
    cl = [[1 if k % 2 else -1, 2 + k, 1300 + k] for k in range(1200)]
    for _ in range(6000):
        vs = rng.sample(range(2, 3001), 3)
        cl.append([v if rng.random() < .5 else -v for v in vs])
    out["hub"] = (3000, cl, [])
    out["hub_wide_mu"] = (3000, cl, ["--mupos=2000", "--muneg=2000", "--subsumemaxoccurs=100"])
    # a literal in 3000 clauses: the order fix of its occurrence list takes the heap-sort route (> 1024 entries)
    cl = [[1, 2 + k, 3100 + (k % 700)] for k in range(3000)] + [[-1, 2 + k, 3 + k] for k in range(50)]
    for _ in range(5000):
        vs = rng.sample(range(2, 3801), 3)
        cl.append([v if rng.random() < .5 else -v for v in vs])
    out["hub_huge"] = (3800, cl, [])
    out["emptied_4"] = (3, [[1, 2], [-1, 3], [2, 3], [-2, -3, 1]], [])
    out["emptied_2"] = (2, [[1, 2], [-1, -2]], [])
    return out


CASES = _cases()


def _reference(tmp_path, name):
    nv, cl, opts = CASES[name]
    cnf = str(tmp_path / (name + ".cnf"))
    with open(cnf, "w") as f:
        f.write(f"p cnf {nv} {len(cl)}\n")
        for c in cl:
            f.write(" ".join(map(str, c)) + " 0\n")
    fin, fout = str(tmp_path / (name + ".in")), str(tmp_path / (name + ".out"))
    subprocess.run([REF_BIN, cnf, "-quiet", f"--sgin={fin}", f"--sgout={fout}", *opts], capture_output=True, timeout=120)
    return B.read(fin), B.read(fout)


def _check(s, tmp_path, name):
    bi, bo = _reference(tmp_path, name)
    try:
        out, rep = s.simplify(bi)
    except sigma.Unsat:
        out, rep = B.Boundary(status=1), {}
    assert parity.diff(out, bo) == []
    if name == "ties_gt24":
        assert rep["sort_ties_gt24"] > 0          # the std::sort branch with equal keys really ran
    if name.startswith("emptied"):
        assert out.refs.size == 0


@needs_ref
@pytest.mark.parametrize("name", sorted(CASES))
def test_edge_case_cpu_standin(emul_lib, tmp_path, name):
    s = sigma.Sigma(0, emul_lib)
    _check(s, tmp_path, name)
    s.close()


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_edge_case_gpu(gpu_sigma, tmp_path, name):
    _check(gpu_sigma, tmp_path, name)


def _abi_errors(lib_path):
    s = sigma.Sigma(0, lib_path)
    lib = s.lib
    # execute / download before upload
    assert lib.sigma_execute(s.ctx) == sigma.SIGMA_E_ARG
    f = sigma.SigmaFormula()
    assert lib.sigma_download(s.ctx, C.byref(f)) == sigma.SIGMA_E_ARG
    assert lib.sigma_upload(s.ctx, None, None) == sigma.SIGMA_E_ARG
    # null arrays
    assert lib.sigma_upload(s.ctx, None, C.byref(f)) == sigma.SIGMA_E_ARG
    # output buffers too small: required sizes come back, the result stays downloadable
    bi = B.read(os.path.join(GOLDEN, "fuzz_ok_a.in"))
    s.upload(bi)
    s.execute()
    small = sigma.SigmaFormula()
    assert lib.sigma_download(s.ctx, C.byref(small)) == sigma.SIGMA_E_NOSPACE
    assert small.n_buckets > 0 and small.n_clauses > 0
    out = s.download()
    bo = B.read(os.path.join(GOLDEN, "fuzz_ok_a.out"))
    assert parity.diff(out, bo) == []
    assert lib.sigma_strerror(sigma.SIGMA_E_NOSPACE).decode() != ""
    # malformed input is answered with SIGMA_E_ARG, never followed (the ABI's promise): scalars first ...
    def code_of(mutate, run=False):
        b2 = B.read(os.path.join(GOLDEN, "fuzz_ok_a.in"))
        f2 = sigma.Sigma.make_formula(b2)
        keep = mutate(b2, f2)                       # noqa: F841  (keeps replaced arrays alive)
        rc = lib.sigma_upload(s.ctx, C.byref(sigma.Sigma.make_opts(b2.opts)), C.byref(f2))
        if rc == sigma.SIGMA_OK and run:
            rc = lib.sigma_execute(s.ctx)
        return rc

    def set_(f2, **kw):
        for k, v in kw.items():
            setattr(f2, k, v)
    assert code_of(lambda b2, f2: set_(f2, trail_size=b2.nvars + 1)) == sigma.SIGMA_E_ARG
    assert code_of(lambda b2, f2: set_(f2, propagated=f2.trail_size + 1)) == sigma.SIGMA_E_ARG
    assert code_of(lambda b2, f2: set_(f2, n_literals=f2.n_buckets)) == sigma.SIGMA_E_ARG
    assert code_of(lambda b2, f2: set_(f2, n_buckets=3 * f2.n_clauses - 1)) == sigma.SIGMA_E_ARG
    # the device store uses 32-bit bucket offsets (the reference's S_REF is 64-bit): a larger arena is refused, not truncated
    assert code_of(lambda b2, f2: set_(f2, n_buckets=1 << 32)) == sigma.SIGMA_E_UNSUPPORTED
    assert code_of(lambda b2, f2: set_(f2, nvars=1 << 29)) == sigma.SIGMA_E_UNSUPPORTED
    # ... then refs / size words that leave the arena: caught on the device by the ingest kernels

    def bad_ref(b2, f2):
        r = b2.refs.copy(); r[r.size // 2] = b2.arena.size + 100
        f2.refs = r.ctypes.data_as(C.c_void_p)
        return r

    def bad_size(b2, f2):
        a = b2.arena.copy(); a[int(b2.refs[b2.refs.size // 2]) + 2] = 0x7FFFFFF0
        f2.arena = a.ctypes.data_as(C.c_void_p)
        return a

    def overlap(b2, f2):                            # two refs to the same record: more literals than the arena holds
        r = b2.refs.copy(); r[1:] = r[0]; a = b2.arena.copy(); a[int(r[0]) + 2] = 40
        f2.refs, f2.arena = r.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p)
        return r, a
    assert code_of(bad_ref, run=True) == sigma.SIGMA_E_ARG
    assert code_of(bad_size, run=True) == sigma.SIGMA_E_ARG
    assert code_of(overlap, run=True) == sigma.SIGMA_E_ARG
    # a run cut short by sigma_set_stop leaves nothing to download
    s.upload(bi)
    s.set_stop(0, "elect")
    s.execute()
    assert lib.sigma_download(s.ctx, C.byref(sigma.SigmaFormula())) == sigma.SIGMA_E_ARG
    s.set_stop(-1, -1)
    # collect_freq = 0 would divide by zero in the reference (solve.hpp:628): rejected
    with pytest.raises(sigma.SigmaError) as e:
        s.simplify(bi, dict(bi.opts, collect_freq=0))
    assert e.value.code == sigma.SIGMA_E_ARG
    s.close()


def test_abi_error_paths_cpu_standin(emul_lib):
    _abi_errors(emul_lib)


@pytest.mark.gpu
def test_abi_error_paths_gpu(cuda_lib):
    _abi_errors(cuda_lib)


def test_no_device_fails_loudly(cuda_lib):
    """In the GPU-less container the CUDA library must refuse to create a context (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(sigma.SigmaError) as e:
        sigma.Sigma(0, cuda_lib)
    assert e.value.code == sigma.SIGMA_E_CUDA and "no CPU fallback" in str(e.value)


def test_header_symbols_are_exported(cuda_lib):
    """Every entry point include/sigma_b200.h declares is exported by libsigma_b200.so."""
    import re
    hdr = open(os.path.join(os.path.dirname(GOLDEN), "..", "include", "sigma_b200.h")).read()
    names = set(re.findall(r"\b(sigma_[a-z_]+)\s*\(", hdr))
    lib = C.CDLL(cuda_lib)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(sigma.EXPORTS) <= names


def test_ctypes_mirror_matches_the_header(tmp_path):
    """sizeof of the five ABI structs as gcc sees them == the ctypes mirrors in seqfrost_b200/sigma.py."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "sigma_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(sigma_opts), '
                   'sizeof(sigma_formula), sizeof(sigma_report), sizeof(sigma_cnf), sizeof(sigma_cnf_out), sizeof(sigma_dimacs), '
                   'sizeof(sigma_watches));return 0;}\n')
    exe = tmp_path / "sz"
    inc = os.path.join(os.path.dirname(GOLDEN), "..", "include")
    subprocess.run(["gcc", "-I", inc, str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(sigma.SigmaOpts), C.sizeof(sigma.SigmaFormula), C.sizeof(sigma.SigmaReport),
                   C.sizeof(sigma.SigmaCnf), C.sizeof(sigma.SigmaCnfOut), C.sizeof(sigma.SigmaDimacs), C.sizeof(sigma.SigmaWatches)]


def _gapped(b: B.Boundary, pad=2):
    """Same formula in an arena with `pad` unused buckets after every clause (the general ingest path:
    refs are then not ref[i] = ref[i-1] + 3 + size)."""
    g = B.Boundary(**{k: getattr(b, k) for k in ("status", "nvars", "n_clauses", "n_literals", "unassigned", "max_frozen", "max_melted",
                                                  "propagated", "simplify_calls", "state", "value", "trail", "vorg", "model", "opts")})
    sizes = b.arena[b.refs.astype(np.int64) + 2].astype(np.int64)
    newrefs = np.concatenate([[0], np.cumsum(sizes + 3 + pad)[:-1]]).astype(np.uint64)
    arena = np.full(int((sizes + 3 + pad).sum()), 0xDEADBEEF, dtype=np.uint32)
    for r_old, r_new, sz in zip(b.refs.astype(np.int64).tolist(), newrefs.astype(np.int64).tolist(), sizes.tolist()):
        arena[r_new:r_new + 3 + sz] = b.arena[r_old:r_old + 3 + sz]
    g.arena, g.refs = arena, newrefs
    return g


def _gapped_check(lib_path):
    for name in ("fuzz_ok_a", "gates_w200"):
        bi, bo = B.read(os.path.join(GOLDEN, name + ".in")), B.read(os.path.join(GOLDEN, name + ".out"))
        s = sigma.Sigma(0, lib_path)
        f = sigma.Sigma.make_formula(_gapped(bi))
        out, _ = s.simplify(_gapped(bi))
        assert parity.diff(out, bo) == []
        s.close()


def test_arena_with_gaps_cpu_standin(emul_lib):
    _gapped_check(emul_lib)


@pytest.mark.gpu
def test_arena_with_gaps_gpu(cuda_lib):
    _gapped_check(cuda_lib)


def _incremental_and_interrupt(lib_path):
    """(1) The incremental freeze mask (`ifrozen`, solve.hpp:789,816): LCVE must skip frozen / assumed variables
    (lcve.cpp:67); the fixture comes from the reference's incremental solver (ref_harness --incr).  Dropping the mask must
    change the result, otherwise the fixture would not test anything.  (2) The interrupt flag: set before the pass =>
    SIGMA_INTERRUPTED and no result; cleared => the same uploaded input runs to the reference's result."""
    s = sigma.Sigma(0, lib_path)
    for name in ("incr_gates", "incr_ksat3"):
        bi, bo = B.read(os.path.join(GOLDEN, name + ".in")), B.read(os.path.join(GOLDEN, name + ".out"))
        assert bi.ifrozen is not None and bi.ifrozen.sum() > 0
        out, _ = s.simplify(bi)
        assert parity.diff(out, bo) == []
        assert not (bi.ifrozen[out.eliminated()]).any()
        frozen, bi.ifrozen = bi.ifrozen, None
        out2, _ = s.simplify(bi)
        assert frozen[out2.eliminated()].any()          # without the mask some frozen variable is eliminated
        bi.ifrozen = frozen
    flag = np.ones(1, dtype=np.uint8)
    s.set_interrupt_flag(flag)
    s.upload(bi)
    assert s.lib.sigma_execute(s.ctx) == sigma.SIGMA_INTERRUPTED
    assert s.lib.sigma_download(s.ctx, C.byref(sigma.SigmaFormula())) == sigma.SIGMA_E_ARG
    flag[0] = 0
    s.execute()
    assert parity.diff(s.download(), bo) == []
    s.set_interrupt_flag(None)
    s.close()


def test_incremental_mask_and_interrupt_cpu_standin(emul_lib):
    _incremental_and_interrupt(emul_lib)


@pytest.mark.gpu
def test_incremental_mask_and_interrupt_gpu(cuda_lib):
    _incremental_and_interrupt(cuda_lib)
