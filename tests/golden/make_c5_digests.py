"""Golden digests of BASELINE.json configs[4] (the 64-instance C5 batch): for every seed the UNMODIFIED reference
(oracle/_ref/ref_sigma, default options) simplifies the instance's DIMACS file; the sha256 of its whole output boundary
(seqfrost_b200/batch.py: boundary_digest) goes into tests/golden/c5_digests.json, which `bench.py --config c5 --verify`
and tests/test_batch_queue.py compare the CUDA path against.  Also checks that the boundary bench.py feeds the GPU
(boundary.from_cnf, the host mirror of awaken()) is word for word what the reference's own awaken() produced.

    python tests/golden/make_c5_digests.py [n_instances] [workers]        (about 10 minutes on 8 cores)
"""
import json
import os
import subprocess
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from seqfrost_b200 import batch, boundary as B  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_sigma")
CACHE = os.environ.get("SIGMA_BENCH_CACHE", "/tmp/seqfrost_b200_bench")


def one(it):
    seed = it["seed"]
    batch.c5_prepare_one((it, CACHE, True))
    cnf, fin, fout = batch.c5_path(CACHE, it, "cnf"), batch.c5_path(CACHE, it, "refin"), batch.c5_path(CACHE, it, "refout")
    subprocess.run([REF_BIN, cnf, "-quiet", f"--sgin={fin}", f"--sgout={fout}"], capture_output=True, timeout=3600, check=True)
    ri, ro, mine = B.read(fin), B.read(fout), B.read(batch.c5_path(CACHE, seed))
    same_in = bool((ri.arena.size == mine.arena.size) and (ri.arena == mine.arena).all() and (ri.refs == mine.refs).all() and ri.nvars == mine.nvars)
    rec = {"seed": seed, "digest": batch.boundary_digest(ro) if ro.status == 0 else "UNSAT", "input_matches_from_cnf": same_in,
           "clauses_in": int(ri.refs.size), "literals_in": int(ri.n_literals), "clauses_out": int(ro.n_clauses) if ro.status == 0 else 0}
    for f in (fin, fout):
        os.remove(f)
    return rec


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    items = batch.c5_items(n)
    with ProcessPoolExecutor(max_workers=workers) as ex:
        recs = list(ex.map(one, items))
    bad = [r["seed"] for r in recs if not r["input_matches_from_cnf"]]
    out = {"what": "sha256 of the reference's output boundary per C5 instance (seeds 100..163, cnfgen.mixed_instance), reference default options",
           "made_by": "tests/golden/make_c5_digests.py", "input_mismatch_seeds": bad,
           "digests": {str(r["seed"]): r["digest"] for r in recs},
           "sizes": {str(r["seed"]): [r["clauses_in"], r["literals_in"], r["clauses_out"]] for r in recs}}
    with open(os.path.join(ROOT, "tests", "golden", "c5_digests.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("instances", len(recs), "input mismatches", bad)
