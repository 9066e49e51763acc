"""Dev aid (not a test): the device DIMACS parser on the bench instance's file (CLAUSES / VARS select the size) next to the
reference's own parser on one host core.  Prints one JSON line."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from seqfrost_b200 import sigma, cnfgen

nv, nc = int(os.environ.get("VARS", 2_000_000)), int(os.environ.get("CLAUSES", 10_000_000))
path = f"/tmp/perf_dimacs_{nv}_{nc}.cnf"
if not os.path.exists(path):
    cnfgen.write_dimacs(path, *cnfgen.planted_subsumption_ksat(nv, nc, k=5, seed=7))
size = os.path.getsize(path)
s = sigma.Sigma(0)
pin = sigma.PinnedArray(s.lib, np.dtype(np.uint8), size)
with open(path, "rb") as f:
    f.readinto(memoryview(pin.array))
res = []
for _ in range(4):
    t0 = time.perf_counter()
    d = s.parse_dimacs(pin.array, download=False)
    res.append((1e3 * (time.perf_counter() - t0), d["h2d_ms"], d["device_ms"]))
wall, h2d, dev = min(res)
out = {"file_MB": size / 1e6, "clauses": d["n_orgs"], "literals": d["literals"], "unit_rounds": d["unit_rounds"],
       "wall_ms": wall, "h2d_ms": h2d, "device_ms": dev, "device_GBps": size / 1e6 / dev, "e2e_GBps": size / 1e6 / wall}
ref = os.path.join(ROOT, "oracle", "_ref", "ref_sigma")
if os.path.exists(ref) and os.environ.get("NO_REF") != "1":
    t0 = time.perf_counter()
    subprocess.run([ref, path, "-quiet", "--parseout=/tmp/perf_dimacs.parse"], capture_output=True)
    out["reference_parse_s_incl_dump"] = time.perf_counter() - t0
print(json.dumps(out))
