"""The DIMACS parser (reference dimacs.cpp:25-177): oracle restatement and the data-parallel version against the
UNMODIFIED reference's own parser (`ref_sigma --parseout`, dumped right after parser() returns): CMM arena, orgs, parse-time
units in order, value[], state[], header and formula counters, word for word.  CPU: oracle + the serial stand-in of the
launch layer (same per-item bodies as the kernels); -m gpu: the CUDA kernels through the C ABI."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import REF_BIN, ROOT
from seqfrost_b200 import cnfgen, sigma

needs_ref = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ref_sigma not built")
PORT_LIB = os.path.join(ROOT, "oracle", "libsigma_oracle.so")

HAND = {
    "units_chain": "c hello\np cnf 5 6\n1 -2 3 0\n-1 2 0 4\n 5 0\n-4 -5 1 0\n2 2 -3 0 c x\n3 0 1 -3 0 -3 2 0\n",
    "weird_tokens": "c a\nc b\n\n\np cnf 9 9\n1-2+3 -0\n+4\t-5\r\n6 0 007 8 0c tail comment\n9 -9 1 0\n  1 2 00 3 4 0\n%\n0\n1 2 x\n",
    "tautology_dups": "p cnf 4 8\n1 1 2 0\n1 -1 2 0\n3 3 3 0\n-3 4 0\n4 4 0\n1 2 3 4 1 2 3 4 0\n",
    "falsified": "p cnf 3 5\n1 0\n2 0\n-1 -2 0\n-1 3 0\n-3 1 0\n",
    "conflicting_units": "p cnf 2 4\n1 0\n-1 0\n1 2 0\n-1 2 0\n",
    "no_trailing_newline": "p cnf 3 2\n1 2 0\n-1 3 0",
    "comment_at_eof": "p cnf 3 2\n1 2 0\n-1 3 0\nc the end",
    "unit_cascade": "p cnf 6 7\n1 0\n-1 2 0\n-2 3 0\n-3 4 0\n-4 5 0\n-5 6 0\n-6 -1 0\n",
    "header_spread": "c x\np cnf\n 4\n\n 3\n1 2 0\n3 4 0\n-1 -3 0\n",
}
ERRORS = {
    "empty_clause": ("p cnf 2 3\n1 2 0\n0\n-1 0\n", 1),
    "bad_char": ("p cnf 2 2\n1 a 0\n", 6),
    "comment_in_clause": ("p cnf 2 2\n1\nc no\n2 0\n", 6),
    "var_too_big": ("p cnf 2 2\n1 3 0\n", 6),
    "too_many_clauses": ("p cnf 3 1\n1 2 0\n2 3 0\n", 6),
    "unterminated": ("p cnf 3 2\n1 2 0\n-1 3", 6),
    "no_header": ("c nothing\n1 2 0\n", 6),
    "bad_header": ("p  cnf 2 2\n1 2 0\n", 6),
    "sign_alone": ("p cnf 2 2\n1 - 2 0\n", 6),
}


def read_parse_dump(path):
    with open(path, "rb") as f:
        blob = f.read()
    pos, sec = 8, {}
    while pos < len(blob):
        name = blob[pos:pos + 8].rstrip(b"\0").decode()
        (n,) = struct.unpack_from("<Q", blob, pos + 8)
        sec[name] = blob[pos + 16:pos + 16 + n]
        pos += 16 + n
    m = np.frombuffer(sec["meta"], np.uint64)
    d = dict(ok=int(m[0]), nvars=int(m[1]), header_clauses=int(m[2]), n_units=int(m[4]), binaries=int(m[5]), ternaries=int(m[6]),
             large=int(m[7]), max_clause_size=int(m[8]), literals=int(m[9]), n_orgs=int(m[10]))
    if d["ok"]:
        d.update(cm=np.frombuffer(sec["cm"], np.uint32), orgs=np.frombuffer(sec["orgs"], np.uint64), trail=np.frombuffer(sec["trail"], np.uint32),
                 value=np.frombuffer(sec["value"], np.int8), state=np.frombuffer(sec["state"], np.uint8))
    return d


def reference_parse(tmp_path, name, text):
    cnf, out = str(tmp_path / (name + ".cnf")), str(tmp_path / (name + ".parse"))
    with open(cnf, "wb") as f:
        f.write(text if isinstance(text, bytes) else text.encode())
    r = subprocess.run([REF_BIN, cnf, "-quiet", f"--parseout={out}"], capture_output=True, timeout=600)
    if r.returncode != 0 or not os.path.exists(out):
        return None                                          # LOGERR: the reference stopped on the file
    return read_parse_dump(out)


def oracle_parse(text):
    lib = C.CDLL(PORT_LIB)
    buf = np.frombuffer(text if isinstance(text, bytes) else text.encode(), np.uint8)
    d = sigma.SigmaDimacs()
    cap = buf.size + 64
    cm, orgs, trail = np.zeros(cap, np.uint32), np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)
    value, state = np.zeros(2 * cap + 2, np.int8), np.zeros(cap + 1, np.uint8)
    d.cm, d.cm_cap, d.orgs, d.orgs_cap, d.trail, d.trail_cap = cm.ctypes.data, cap, orgs.ctypes.data, cap, trail.ctypes.data, cap
    d.value, d.state = value.ctypes.data, state.ctypes.data
    lib.sgo_parse_dimacs.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(sigma.SigmaDimacs)]
    rc = lib.sgo_parse_dimacs(buf.ctypes.data, buf.size, C.byref(d))
    out = {k: getattr(d, k) for k in ("nvars", "header_clauses", "cm_buckets", "n_orgs", "n_units", "binaries", "ternaries", "large", "literals", "max_clause_size")}
    out.update(rc=rc, cm=cm[:d.cm_buckets], orgs=orgs[:d.n_orgs], trail=trail[:d.n_units], value=value[:2 * d.nvars + 2], state=state[:d.nvars + 1])
    return out


def same(ours, ref):
    bad = [k for k in ("nvars", "header_clauses", "n_orgs", "n_units", "binaries", "ternaries", "large", "literals", "max_clause_size") if ours[k] != ref[k]]
    bad += [k for k in ("cm", "orgs", "trail", "value", "state") if not np.array_equal(ours[k], ref[k])]
    return bad


def generated(kind):
    if kind == "ksat3":
        nv, off, lits = cnfgen.random_ksat(3000, 12000, k=3, seed=5)
    elif kind == "gates":
        nv, off, lits = cnfgen.gate_circuit(4000, seed=6)
    elif kind == "xorphp":
        nv, off, lits = cnfgen.xor_php_mix(8000, seed=7)
    else:
        nv, off, lits = cnfgen.fuzz_mix(int(kind[4:]))
    import io
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".cnf", delete=False) as f:
        path = f.name
    cnfgen.write_dimacs(path, nv, off, lits)
    with open(path, "rb") as f:
        text = f.read()
    os.remove(path)
    return text


def unit_heavy(seed, nvars=400, nclauses=1500):
    """Random short clauses with many units, duplicates and tautologies: parse-time units act on later clauses."""
    rng = np.random.default_rng(seed)
    lines = [f"c seed {seed}", f"p cnf {nvars} {nclauses}"]
    for _ in range(nclauses):
        k = int(rng.choice([1, 1, 2, 2, 3, 4]))
        v = rng.integers(1, nvars + 1, size=k)
        s = rng.integers(0, 2, size=k)
        lines.append(" ".join(str(int(a) if b else -int(a)) for a, b in zip(v, s)) + " 0")
    return "\n".join(lines) + "\n"


CASES = list(HAND) + ["ksat3", "gates", "xorphp", "fuzz3", "fuzz11", "units7", "units8", "units9"]


def case_text(name):
    if name in HAND:
        return HAND[name]
    if name.startswith("units"):
        return unit_heavy(int(name[5:]))
    return generated(name)


@needs_ref
@pytest.mark.parametrize("name", CASES)
def test_oracle_parser_matches_reference(tmp_path, name):
    text = case_text(name)
    ref = reference_parse(tmp_path, name, text)
    assert ref is not None and ref["ok"] == 1
    o = oracle_parse(text)
    assert o["rc"] == 0
    assert same(o, ref) == []


@needs_ref
@pytest.mark.parametrize("name", sorted(ERRORS))
def test_oracle_parser_stops_where_the_reference_stops(tmp_path, name):
    text, want = ERRORS[name]
    ref = reference_parse(tmp_path, name, text)
    assert ref is None or ref["ok"] == 0                     # LOGERR (process ends) or parser() == false
    assert (ref is not None and ref["ok"] == 0) == (want == 1)
    assert oracle_parse(text)["rc"] == want


def _ours(s, text):
    try:
        return s.parse_dimacs(text if isinstance(text, bytes) else text.encode()), 0
    except sigma.Unsat:
        return None, 1
    except sigma.SigmaError as e:
        return None, e.code


@needs_ref
@pytest.mark.parametrize("name", CASES)
def test_emulated_parser_matches_reference(emul_lib, tmp_path, name):
    text = case_text(name)
    ref = reference_parse(tmp_path, name, text)
    s = sigma.Sigma(0, emul_lib)
    got, rc = _ours(s, text)
    assert rc == 0 and same(got, ref) == []
    assert got["unit_rounds"] >= 1
    s.close()


@pytest.mark.parametrize("name", sorted(ERRORS))
def test_emulated_parser_error_codes(emul_lib, name):
    text, want = ERRORS[name]
    s = sigma.Sigma(0, emul_lib)
    assert _ours(s, text)[1] == want
    s.close()


def test_emulated_parser_matches_oracle_on_unit_fuzz(emul_lib):
    s = sigma.Sigma(0, emul_lib)
    rounds = 0
    for seed in range(40):
        text = unit_heavy(100 + seed, nvars=60 + seed, nclauses=300)
        o = oracle_parse(text)
        got, rc = _ours(s, text)
        assert rc == o["rc"]
        if rc == 0:
            assert same(got, o) == [], seed
            rounds = max(rounds, got["unit_rounds"])
    assert rounds > 2                                        # units really did act on later clauses
    s.close()


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("name", CASES)
def test_cuda_parser_matches_reference(gpu_sigma, tmp_path, name):
    text = case_text(name)
    ref = reference_parse(tmp_path, name, text)
    got, rc = _ours(gpu_sigma, text)
    assert rc == 0 and same(got, ref) == []


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(ERRORS))
def test_cuda_parser_error_codes(gpu_sigma, name):
    text, want = ERRORS[name]
    assert _ours(gpu_sigma, text)[1] == want


@pytest.mark.gpu
def test_cuda_parser_matches_oracle_on_unit_fuzz(gpu_sigma):
    for seed in range(40):
        text = unit_heavy(100 + seed, nvars=60 + seed, nclauses=300)
        o = oracle_parse(text)
        got, rc = _ours(gpu_sigma, text)
        assert rc == o["rc"]
        if rc == 0:
            assert same(got, o) == [], seed


@pytest.mark.gpu
@needs_ref
def test_cuda_parser_on_a_large_file(gpu_sigma, tmp_path):
    """1M-clause 5-SAT (42 MB of text): the file the reference parses in about a second."""
    nv, off, lits = cnfgen.planted_subsumption_ksat(200_000, 1_000_000, k=5, seed=7)
    path = str(tmp_path / "big.cnf")
    cnfgen.write_dimacs(path, nv, off, lits)
    with open(path, "rb") as f:
        text = f.read()
    ref = reference_parse(tmp_path, "big", text)
    got, rc = _ours(gpu_sigma, text)
    assert rc == 0 and same(got, ref) == []
    again, rc = _ours(gpu_sigma, text)                       # buffers allocated: the kernels alone
    assert rc == 0 and same(again, ref) == []
    assert again["device_ms"] > 0 and len(text) / 1e6 / again["device_ms"] > 5.0      # > 5 GB/s of text on the device


def _file_to_pass(s, tmp_path, name, inst):
    """DIMACS text -> parse -> extract -> pass, all on the device, against the reference run on the same file."""
    import parity
    from seqfrost_b200 import boundary as B
    nv, off, lits = inst
    cnf, fout = str(tmp_path / (name + ".cnf")), str(tmp_path / (name + ".out"))
    cnfgen.write_dimacs(cnf, nv, off, lits)
    subprocess.run([REF_BIN, cnf, "-quiet", f"--sgout={fout}"], capture_output=True, timeout=600, check=True)
    with open(cnf, "rb") as f:
        text = f.read()
    d = s.parse_dimacs(text, download=False)
    assert d["n_units"] == 0
    s.upload_parsed({"ere_en": 1}, nvars=d["nvars"])
    s.execute()
    out = s.download()
    assert parity.diff(out, B.read(fout)) == []


@needs_ref
def test_emulated_file_to_pass_without_leaving_the_device(emul_lib, tmp_path):
    s = sigma.Sigma(0, emul_lib)
    _file_to_pass(s, tmp_path, "g", cnfgen.gate_circuit(1500, seed=21))
    _file_to_pass(s, tmp_path, "k", cnfgen.random_ksat(800, 3300, k=3, seed=22))
    s.close()


def test_upload_parsed_refuses_files_with_units(emul_lib):
    s = sigma.Sigma(0, emul_lib)
    s.parse_dimacs(HAND["unit_cascade"].encode(), download=False)
    with pytest.raises(sigma.SigmaError) as e:
        s.upload_parsed()
    assert e.value.code == 7
    s.close()


@pytest.mark.gpu
@needs_ref
def test_cuda_file_to_pass_without_leaving_the_device(gpu_sigma, tmp_path):
    _file_to_pass(gpu_sigma, tmp_path, "g", cnfgen.gate_circuit(60_000, seed=21))
    _file_to_pass(gpu_sigma, tmp_path, "k", cnfgen.planted_subsumption_ksat(40_000, 200_000, k=5, seed=22))
    _file_to_pass(gpu_sigma, tmp_path, "x", cnfgen.xor_php_mix(50_000, seed=23))
