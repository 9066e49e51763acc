"""Dev aid (not a test): createOT stage times of the bench instance under different bucket widths (SIGMA_OTB_SHIFT)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seqfrost_b200 import boundary as B, cnfgen, sigma
shape = os.environ.get("SHAPE", "c2")
gen = {"c3": lambda: cnfgen.gate_circuit(int(os.environ.get("VARS", 5_000_000)), seed=11), "c4": lambda: cnfgen.xor_php_mix(1_000_000, seed=3),
       "c1": lambda: cnfgen.random_ksat(100_000, 426_000, k=3, seed=1),
       "c2": lambda: cnfgen.planted_subsumption_ksat(2_000_000, 10_000_000, k=5, seed=7)}[shape]
b = B.from_cnf(*gen())
b.opts["ere_en"] = 1
s = sigma.Sigma(0)
s.upload(b)
for sh in os.environ.get("SHIFTS", ",10,11,12").split(","):
    if sh:
        os.environ["SIGMA_OTB_SHIFT"] = sh
    elif "SIGMA_OTB_SHIFT" in os.environ:
        del os.environ["SIGMA_OTB_SHIFT"]
    acc = {}
    for i in range(5):
        s.execute()
        if i >= 2:
            for k, v in s.report()["stage_ms"].items():
                acc[k] = acc.get(k, 0) + v / 3
    print("shift", sh or "auto", {k: round(acc[k], 3) for k in ("ot_hist", "ot_scan", "ot_scatter", "ot_order", "total_device")}, flush=True)
