"""Dev aid (not a test): `seqfrost_sigma <bench instance> -no-solve` with SIGMA_BINDING_TIMES=1 - where the host side of
the drop-in spends its time."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seqfrost_b200 import cnfgen
nv, nc = int(os.environ.get("VARS", 2_000_000)), int(os.environ.get("CLAUSES", 10_000_000))
path = f"/tmp/perf_dimacs_{nv}_{nc}.cnf"
if not os.path.exists(path):
    cnfgen.write_dimacs(path, *cnfgen.planted_subsumption_ksat(nv, nc, k=5, seed=7))
binp = os.path.join(ROOT, "seqfrost_b200", "host", "_build", "seqfrost_sigma")
for env in ({}, {"SIGMA_HOST_REBUILDWT": "1"}):
    r = subprocess.run([binp, path, "-no-solve", "-report"], capture_output=True, text=True, env={**os.environ, "SIGMA_BINDING_TIMES": "1", **env})
    print("---", env)
    print("\n".join(l for l in r.stderr.splitlines() if "time:" in l))
    print("\n".join(l for l in r.stdout.splitlines() if "Simplifier time" in l or "Parsing" in l or "Read " in l))
