"""Dev aid (not a test): prop() on a unit storm, frontier kernels against the one-thread walk (SIGMA_PROP_WIDE=0).
Prints the `prop` stage time of every repetition, the number of units and whether the two results are identical."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity
from seqfrost_b200 import boundary as B, cnfgen, sigma
n, depth = int(os.environ.get("SEEDS", 100_000)), int(os.environ.get("DEPTH", 8))
b = B.from_cnf(*cnfgen.unit_storm(n, depth=depth, seed=5))
b.opts["ere_en"] = 1
res = {}
for wide in os.environ.get("WIDES", "64,0").split(","):
    os.environ["SIGMA_PROP_WIDE"] = wide
    s = sigma.Sigma(0)
    s.upload(b)
    for i in range(int(os.environ.get("REPS", 3))):
        s.execute()
        r = s.report()
        print("wide", wide, "rep", i, "prop_ms", round(r["stage_ms"]["prop"], 3), "pass_ms", round(r["stage_ms"]["total_device"], 2), "launches", r["kernel_launches"],
              "round_trips", r["host_round_trips"], flush=True)
    out = s.download()
    res[wide] = (out, r["stage_ms"]["prop"])
    print("wide", wide, "units", out.trail.size, flush=True)
    s.close()
ks = list(res)
if len(ks) > 1:
    print("identical", parity.diff(res[ks[0]][0], res[ks[1]][0]) == [], "speedup", round(res[ks[-1]][1] / max(res[ks[0]][1], 1e-9), 1))
