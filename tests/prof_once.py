"""Dev aid (not a test): ONE pass over the bench instance - device-side extract, execute, device-side newBeginning -
for ncu captures (`ncu --set full -k regex:... python tests/prof_once.py`).  VARS / CLAUSES select the size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seqfrost_b200 import sigma, boundary as B, cnfgen

nv, nc = int(os.environ.get("VARS", 2_000_000)), int(os.environ.get("CLAUSES", 10_000_000))
b = B.from_cnf(*cnfgen.planted_subsumption_ksat(nv, nc, k=5, seed=7))
b.opts["ere_en"] = 1
db = B.clause_db_from_boundary(b, seed=7, dead_frac=0.05)
s = sigma.Sigma(0)
s.upload_cnf(db, b)
s.execute()
out, rest = s.download_cnf()
r = s.report()
print({k: round(v, 3) for k, v in r["stage_ms"].items() if v}, r["extract_us"], r["host_round_trips"], out.orgs.size)
