"""Dev aid (not a test): repeat the pass on one config shape and print the stage times of every repetition (run-to-run spread)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seqfrost_b200 import boundary as B, cnfgen, sigma
shape = os.environ.get("SHAPE", "c3")
gen = {"c3": lambda: cnfgen.gate_circuit(int(os.environ.get("VARS", 5_000_000)), seed=11), "c4": lambda: cnfgen.xor_php_mix(1_000_000, seed=3),
       "c1": lambda: cnfgen.random_ksat(100_000, 426_000, k=3, seed=1), "smoke": lambda: cnfgen.random_ksat(2000, 8520, k=3, seed=1),
       "c2": lambda: cnfgen.planted_subsumption_ksat(2_000_000, 10_000_000, k=5, seed=7)}[shape]
b = B.from_cnf(*gen())
b.opts["ere_en"] = 1
s = sigma.Sigma(0)
s.upload(b)
for i in range(int(os.environ.get("REPS", 8))):
    s.execute()
    r = s.report()
    print(i, round(r["stage_ms"]["total_device"], 2), {k: round(v, 2) for k, v in r["stage_ms"].items() if k in os.environ.get("STAGES", "ere,elect,ve,sub").split(",")}, r["ere"], r["host_round_trips"], flush=True)
