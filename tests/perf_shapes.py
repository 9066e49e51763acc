"""Stage-time breakdown of the pass on the BASELINE.json config shapes (a development aid, run on the GPU box:
`python tests/perf_shapes.py`); not a pytest module."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seqfrost_b200 import boundary as B, cnfgen, sigma  # noqa: E402

SHAPES = {
    "C3 gates 5M vars": lambda: cnfgen.gate_circuit(5_000_000, seed=11),
    "C4 xor/php 1M clauses": lambda: cnfgen.xor_php_mix(1_000_000, seed=3),
    "C1 3-SAT 100k vars": lambda: cnfgen.random_ksat(100_000, 426_000, k=3, seed=1),
}

if __name__ == "__main__":
    s = sigma.Sigma(0)
    only = os.environ.get("SHAPE")
    for name, gen in SHAPES.items():
        if only and not name.startswith(only):
            continue
        t = time.time()
        b = B.from_cnf(*gen())
        b.opts["ere_en"] = 1
        gen_s = time.time() - t
        s.upload(b)
        s.execute()
        s.execute()
        r = s.report()
        print(json.dumps({"shape": name, "gen_s": round(gen_s, 1), "clauses": int(b.refs.size), "literals": int(b.n_literals),
                          "device_ms": round(r["stage_ms"]["total_device"], 2), "phases": r["phases_run"],
                          "elected": r["elected_per_phase"], "rounds": r["election_rounds"],
                          "stage_ms": {k: round(v, 2) for k, v in r["stage_ms"].items() if v > 0.005},
                          "eliminated": r["max_melted_delta"], "gates": [r[k] for k in ("inverters", "andors", "ites", "xors", "funs", "resolutions")],
                          "launches": r["kernel_launches"], "ere": r["ere"]}))
    s.close()
