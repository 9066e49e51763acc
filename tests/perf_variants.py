"""Dev aid (not a test): time the pass with alternative builds of the library (seqfrost_b200/csrc/variants/*.so)."""
import glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seqfrost_b200 import sigma, boundary as B, cnfgen

nv, nc = int(os.environ.get("VARS", 2_000_000)), int(os.environ.get("CLAUSES", 10_000_000))
b = B.from_cnf(*cnfgen.planted_subsumption_ksat(nv, nc, k=5, seed=7))
b.opts["ere_en"] = 1
for lib in sorted(glob.glob(os.path.join(ROOT, "seqfrost_b200", "csrc", "variants", "*.so"))):
    s = sigma.Sigma(0, lib)
    s.upload(b)
    acc, N = {}, 6
    for i in range(N + 2):
        s.execute()
        if i >= 2:
            for k, v in s.report()["stage_ms"].items():
                acc[k] = acc.get(k, 0) + v / N
    print(os.path.basename(lib), json.dumps({k: round(v, 2) for k, v in acc.items() if v and k not in ("h2d", "d2h")}))
    s.close()
