"""extract() on the device (reference simplify.cpp:118-153): the solver's clause database goes in, the
SCNF the reference's awaken() builds must come out bit for bit - and the pass on top of it must give the reference's
result.  Fixtures: tests/golden/<name>.cm (clause database right before awaken(), dumped from the unmodified
reference by oracle/ref_harness --cmin) next to <name>.in (the SCNF right after awaken()) and <name>.out."""
import os

import numpy as np
import pytest

import parity
from conftest import GOLDEN
from seqfrost_b200 import boundary as B
from seqfrost_b200 import cnfgen, sigma

CM_CASES = ["gates_small", "ksat5_planted", "xorphp_small", "fuzz_units4", "ere_fuzz", "later_ksat3", "later_gates"]


def _load(name):
    return (B.read_clause_db(os.path.join(GOLDEN, name + ".cm")), B.read(os.path.join(GOLDEN, name + ".in")),
            B.read(os.path.join(GOLDEN, name + ".out")))


def _check_extract(s, db, bi):
    s.upload_cnf(db, bi)
    n, nb, nl = s.extracted_sizes()
    assert (n, nb, nl) == (bi.refs.size, bi.arena.size, bi.n_literals)            # inf.nClauses / inf.nLiterals after extract()
    assert np.array_equal(s.snapshot("arena_in", np.uint32), bi.arena)            # status, usage, lbd, sig, size, sorted literals
    assert np.array_equal(s.snapshot("refs_in", np.uint64), bi.refs)


def _run_cases(lib, names):
    s = sigma.Sigma(0, lib)
    for name in names:
        db, bi, bo = _load(name)
        _check_extract(s, db, bi)
        s.execute()
        assert parity.diff(s.download(), bo) == [], name
    s.close()


def test_fixture_clause_databases_are_what_the_reference_extracts_from():
    """Sanity of the fixtures themselves (numpy restatement of extract()): every live record of the .cm, originals first,
    is the corresponding clause of the .in with its literals in some other order."""
    for name in CM_CASES:
        db, bi, _ = _load(name)
        refs = np.concatenate([db.orgs, db.learnts]).astype(np.int64)
        live = refs[db.stencil[refs] == 0]
        assert live.size == bi.refs.size
        w0, sig, sz, off, lits = bi.soa()
        assert np.array_equal(db.cm[live + 1], sz.astype(np.uint32))
        assert np.array_equal((db.cm[live] >> 8) & 0xFF, w0 & 3)                   # _l  <->  LEARNT status
        for k in (0, live.size // 2, live.size - 1):
            r = int(live[k])
            assert sorted(db.cm[r + 4:r + 4 + int(sz[k])].tolist()) == lits[off[k]:off[k + 1]].tolist()
        if name.startswith("later"):
            assert db.learnts.size and (w0 & 3).max() == 1                          # learnt clauses, later simplify() call


def test_emul_extract_matches_reference(emul_lib):
    _run_cases(emul_lib, CM_CASES)


def test_emul_extract_synthetic_database(emul_lib):
    """Deleted records in both lists, long clauses (all three sort paths: <= 8, <= 32, heap sort), shuffled literals."""
    rng = np.random.default_rng(5)
    nv, off, lits = cnfgen.random_ksat(300, 900, k=3, seed=9)
    # add clauses of 9..200 distinct literals
    rows = [rng.choice(np.arange(1, nv + 1), size=int(k), replace=False) * rng.choice([-1, 1], int(k))     # DIMACS-signed
            for k in rng.integers(9, 200, 40)]
    off = np.concatenate([off, off[-1] + np.cumsum([r.size for r in rows])])
    lits = np.concatenate([lits] + [r.astype(lits.dtype) for r in rows])
    bi = B.from_cnf(nv, off, lits)
    db = B.clause_db_from_boundary(bi, seed=3, dead_frac=0.2)
    assert db.stencil.sum() > 0
    s = sigma.Sigma(0, emul_lib)
    _check_extract(s, db, bi)
    s.execute()
    out_a = s.download()
    s.upload(bi); s.execute()
    assert parity.diff(out_a, s.download()) == []
    s.close()


def _check_new_beginning(s, name):
    """After execute: the device-built clause database equals the one the reference's newBeginning() builds."""
    exp = B.read_clause_db(os.path.join(GOLDEN, name + ".cmo"))
    db, rest = s.download_cnf()
    mask = np.full(exp.cm.size, 0xFFFFFFFF, dtype=np.uint32)
    mask[np.concatenate([exp.orgs, exp.learnts]).astype(np.int64)] = 0xFFFFFF3F      # 2 padding bits of the CLAUSE bitfield byte
    assert db.cm.size == exp.cm.size
    assert np.array_equal(db.cm & mask, exp.cm & mask), name                          # _k _b _l _used _sz _pos _lbd + literals
    assert np.array_equal(db.orgs, exp.orgs) and np.array_equal(db.learnts, exp.learnts)
    assert (db.lits_original, db.lits_learnt) == (exp.lits_original, exp.lits_learnt)
    assert np.array_equal(exp.subsume_before | db.subsume_after, exp.subsume_after)   # mark_ssubsume, sclause.hpp:209
    bo = B.read(os.path.join(GOLDEN, name + ".out"))
    for k in ("state", "value", "trail", "model"):
        assert np.array_equal(getattr(rest, k), getattr(bo, k)), k
    return db


def test_emul_new_beginning_matches_reference(emul_lib):
    s = sigma.Sigma(0, emul_lib)
    marks = 0
    for name in CM_CASES:
        db, bi, _ = _load(name)
        s.upload_cnf(db, bi)
        s.execute()
        marks += int(_check_new_beginning(s, name).subsume_after.sum())
        s.upload(bi); s.execute()                             # same through the SCNF upload
        _check_new_beginning(s, name)
    assert marks > 0                                          # the later_* cases mark the variables of their resolvents
    s.close()


def test_download_cnf_reports_required_sizes(emul_lib):
    _, bi, _ = _load("gates_small")
    s = sigma.Sigma(0, emul_lib)
    s.upload(bi); s.execute()
    co, f = sigma.SigmaCnfOut(), sigma.SigmaFormula()
    rc = s.lib.sigma_download_cnf(s.ctx, co, f)
    assert rc == sigma.SIGMA_E_NOSPACE and co.cm_buckets > 0 and co.n_orgs > 0
    s.close()


def test_emul_extract_of_an_empty_or_fully_deleted_database(emul_lib):
    """extract() of nothing: no clauses, the pass has nothing to do, newBeginning() hands back an empty database."""
    _, bi, _ = _load("fuzz_units4")
    db, _, _ = _load("fuzz_units4")
    dead = B.ClauseDb(cm=db.cm, orgs=db.orgs, learnts=db.learnts, stencil=np.ones(db.cm.size, np.uint8))
    none = B.ClauseDb(cm=db.cm, orgs=np.zeros(0, np.uint64), learnts=np.zeros(0, np.uint64), stencil=np.zeros(0, np.uint8))
    s = sigma.Sigma(0, emul_lib)
    for d in (dead, none):
        s.upload_cnf(d, bi)
        assert s.extracted_sizes() == (0, 0, 0)
        s.execute()
        out = s.download()
        assert out.refs.size == 0 and out.arena.size == 0
        cdb, _ = s.download_cnf()
        assert cdb.cm.size == 0 and cdb.orgs.size == 0 and cdb.learnts.size == 0 and cdb.lits_original == 0
    s.close()


def test_upload_cnf_argument_errors(emul_lib):
    db, bi, _ = _load("fuzz_units4")
    s = sigma.Sigma(0, emul_lib)
    bad = B.ClauseDb(cm=db.cm, orgs=db.orgs, learnts=db.learnts, stencil=db.stencil)
    bad.cm = None
    with pytest.raises(sigma.SigmaError):
        s.upload_cnf(bad, bi)
    with pytest.raises(sigma.SigmaError):
        s.extracted_sizes()                                  # nothing uploaded
    s.close()


@pytest.mark.gpu
def test_gpu_extract_matches_reference(cuda_lib):
    _run_cases(cuda_lib, CM_CASES)


@pytest.mark.gpu
def test_gpu_new_beginning_matches_reference(cuda_lib):
    s = sigma.Sigma(0, cuda_lib)
    for name in CM_CASES:
        db, bi, _ = _load(name)
        s.upload_cnf(db, bi)
        s.execute()
        _check_new_beginning(s, name)
    s.close()


@pytest.mark.gpu
def test_gpu_extract_at_scale(cuda_lib):
    """1M-clause 5-SAT with 10 % deleted records and shuffled literals: the device-extracted store equals the host one,
    and the pass on top of it gives the same bits as the pass on the host-extracted store."""
    bi = B.from_cnf(*cnfgen.planted_subsumption_ksat(200_000, 1_000_000, k=5, seed=7))
    db = B.clause_db_from_boundary(bi, seed=1, dead_frac=0.1)
    s = sigma.Sigma(0, cuda_lib)
    _check_extract(s, db, bi)
    s.execute()
    a = s.download()
    rep = s.report()
    assert rep["extract_us"] > 0 and rep["extract_bytes"] > 0
    s.upload(bi); s.execute()
    b2 = s.download()
    assert parity.diff(a, b2) == []
    # newBeginning() on the device against its numpy restatement on the downloaded SCNF
    out, _ = s.download_cnf()
    w0, sig, sz, off, lits = b2.soa()
    ref = b2.refs.astype(np.int64) + np.arange(b2.refs.size)
    assert np.array_equal(out.orgs, ref[(w0 & 3) == 0].astype(np.uint64)) and out.learnts.size == 0
    assert np.array_equal(out.cm[ref + 1], sz.astype(np.uint32)) and out.lits_original == int(sz.sum())
    idx = np.repeat(ref + 4 - off[:-1], sz) + np.arange(int(off[-1]), dtype=np.int64)
    assert np.array_equal(out.cm[idx], lits)
    s.close()


# ---- rebuildWT on the device (reference watch.cpp:135-157): the .cmo fixtures also carry the reference's watch table --------
def _read_sections(path):
    import struct
    with open(path, "rb") as f:
        blob = f.read()
    pos, sec = 8, {}
    while pos < len(blob):
        name = blob[pos:pos + 8].rstrip(b"\0").decode()
        (n,) = struct.unpack_from("<Q", blob, pos + 8)
        sec[name] = blob[pos + 16:pos + 16 + n]
        pos += 16 + n
    return sec


def _check_watches(s, name):
    sec = _read_sections(os.path.join(GOLDEN, name + ".cmo"))
    code = int(np.frombuffer(sec["wtcode"], np.uint64)[0])
    woff, wrec = np.frombuffer(sec["wtoff"], np.uint32), np.frombuffer(sec["wtrec"], np.uint32).reshape(-1, 4)
    offs, rec = s.watches(code)
    assert np.array_equal(offs, woff), name
    assert np.array_equal(rec, wrec), name                                   # {ref, imp, size} of every list, in push order
    exp = B.read_clause_db(os.path.join(GOLDEN, name + ".cmo"))
    cm2 = np.frombuffer(sec["cm2"], np.uint32)
    db, _ = s.download_cnf()                                                 # after the watch build: sortClause has run
    mask = np.full(cm2.size, 0xFFFFFFFF, dtype=np.uint32)
    mask[np.concatenate([exp.orgs, exp.learnts]).astype(np.int64)] = 0xFFFFFF3F
    assert np.array_equal(db.cm & mask, cm2 & mask), name
    # the other attach orders against a direct restatement over the downloaded records
    for other in (0, 1, 3):
        offs2, rec2 = s.watches(other)
        db2, _ = s.download_cnf()
        lists = [[] for _ in range(offs2.size - 1)]
        groups = {}
        for learnt, refs in ((0, db2.orgs), (1, db2.learnts)):
            for r in refs.astype(np.int64).tolist():
                size = int(db2.cm[r + 1])
                g = (learnt if not other & 1 else ((learnt ^ 1 if other & 2 else learnt) if size == 2 else 2 + learnt))
                groups.setdefault(g, []).append(r)
        for g in sorted(groups):
            for r in groups[g]:
                size, a, b = int(db2.cm[r + 1]), int(db2.cm[r + 4]), int(db2.cm[r + 5])
                lists[a ^ 1].append((r & 0xFFFFFFFF, r >> 32, b, size))
                lists[b ^ 1].append((r & 0xFFFFFFFF, r >> 32, a, size))
        flat = [w for l in lists for w in l]
        assert np.array_equal(rec2, np.array(flat, dtype=np.uint32).reshape(-1, 4)), (name, other)
        assert offs2.tolist() == np.concatenate([[0], np.cumsum([len(l) for l in lists])]).tolist()


def test_emul_watch_table_matches_reference(emul_lib):
    s = sigma.Sigma(0, emul_lib)
    for name in CM_CASES:
        db, bi, _ = _load(name)
        s.upload_cnf(db, bi)
        s.execute()
        _check_watches(s, name)
    s.close()


@pytest.mark.gpu
def test_gpu_watch_table_matches_reference(cuda_lib):
    s = sigma.Sigma(0, cuda_lib)
    for name in CM_CASES:
        db, bi, _ = _load(name)
        s.upload_cnf(db, bi)
        s.execute()
        _check_watches(s, name)
    s.close()


# ---- newBeginning() under variable mapping: map(true) (reference map.cpp:24-46) ------------------------------------------------
CMM_CASES = ["gates_small", "xorphp_small", "later_gates", "fuzz_units4"]


def _check_mapped(s, name):
    sec = _read_sections(os.path.join(GOLDEN, name + ".cmm"))
    vmap = np.frombuffer(sec["vmap"], np.uint32)
    new_nvars = int(np.frombuffer(sec["vmapn"], np.uint64)[0])
    st = s.download_state()
    # vmap.initiate (map.hpp:78-95) from the downloaded state: active variables in order, plus the first root-assigned one
    mine, n, first = np.zeros_like(vmap), 0, 0
    for v in range(1, st.nvars + 1):
        if not st.state[v]:
            n += 1; mine[v] = n
        elif not first and st.value[2 * v] >= 0:
            first = v; n += 1; mine[v] = n
    assert n == new_nvars and np.array_equal(mine, vmap), name
    s.set_var_map(vmap, new_nvars)
    code = int(np.frombuffer(sec["wtcode"], np.uint64)[0])
    offs, rec = s.watches(code)
    assert np.array_equal(offs, np.frombuffer(sec["wtoff"], np.uint32)), name
    assert np.array_equal(rec, np.frombuffer(sec["wtrec"], np.uint32).reshape(-1, 4)), name
    exp_orgs, exp_learnts = np.frombuffer(sec["orgs"], np.uint64), np.frombuffer(sec["learnts"], np.uint64)
    cm2 = np.frombuffer(sec["cm2"], np.uint32)
    db, _ = s.download_cnf()
    mask = np.full(cm2.size, 0xFFFFFFFF, dtype=np.uint32)
    mask[np.concatenate([exp_orgs, exp_learnts]).astype(np.int64)] = 0xFFFFFF3F
    assert np.array_equal(db.cm & mask, cm2 & mask), name
    assert np.array_equal(db.orgs, exp_orgs) and np.array_equal(db.learnts, exp_learnts)
    s.set_var_map(None, 0)                                   # back to the unmapped database
    _check_new_beginning(s, name) if name in CM_CASES else None


def test_emul_mapped_new_beginning_matches_reference(emul_lib):
    s = sigma.Sigma(0, emul_lib)
    for name in CMM_CASES:
        bi = B.read(os.path.join(GOLDEN, name + ".in"))
        s.upload(bi); s.execute()
        _check_mapped(s, name)
    s.close()


@pytest.mark.gpu
def test_gpu_mapped_new_beginning_matches_reference(cuda_lib):
    s = sigma.Sigma(0, cuda_lib)
    for name in CMM_CASES:
        bi = B.read(os.path.join(GOLDEN, name + ".in"))
        s.upload(bi); s.execute()
        _check_mapped(s, name)
    s.close()
