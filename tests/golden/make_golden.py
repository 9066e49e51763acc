"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref/ref_sigma,
built by `make -C oracle ref` from /root/reference/src).  Run in the build container:

    python tests/golden/make_golden.py

For every case it writes <name>.in (boundary right after awaken(), simplify.cpp:171) and
<name>.out (boundary right after shrinkSimp(), simplify.cpp:210) in the sectioned format of
seqfrost_b200/boundary.py; cases without "+ere" run the reference with -no-redundancy.  For the cases in
CM_CASES it also writes <name>.cm: the solver's clause database right BEFORE awaken() (CMM arena, orgs,
learnts, deleted stencil - the input of extract(), simplify.cpp:118-153), for the device-side extract, and
<name>.cmo: the clause database right AFTER newBeginning() without variable mapping (simplify.cpp:249-268), for
the device-side newBeginning.
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from seqfrost_b200 import cnfgen  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_sigma")

# (name, generator, kwargs, extra reference options)
CASES = [
    ("gates_small", "gate_circuit", dict(nvars=3000, seed=11), []),
    ("gates_w200", "gate_circuit", dict(nvars=1500, seed=5, window=200), []),
    ("ksat3_small", "random_ksat", dict(nvars=2000, nclauses=8520, k=3, seed=1), []),
    ("ksat5_planted", "planted_subsumption_ksat", dict(nvars=1500, nclauses=7500, k=5, seed=7), []),
    ("xorphp_small", "xor_php_mix", dict(nclauses=6000, seed=3), []),
    ("gates_bvebound", "gate_circuit", dict(nvars=1200, seed=21), ["-bvebound"]),
    ("gates_phases2", "gate_circuit", dict(nvars=1200, seed=22), ["--phases=2"]),
    ("gates_nofun_cm4", "gate_circuit", dict(nvars=1200, seed=23), ["-no-function", "--boundedclausemax=4"]),
    ("fuzz_units1", "fuzz_mix", dict(seed=20), []),
    ("fuzz_units2", "fuzz_mix", dict(seed=75), []),
    ("fuzz_units4", "fuzz_mix", dict(seed=115), []),
    ("fuzz_unsat_a", "fuzz_mix", dict(seed=83), []),
    ("fuzz_unsat_b", "fuzz_mix", dict(seed=101), []),
    ("fuzz_ok_a", "fuzz_mix", dict(seed=0), []),
    ("fuzz_ok_b", "fuzz_mix", dict(seed=2), []),
    # ERE on (reference default): "+ere" = do not pass -no-redundancy
    ("ere_gates", "gate_circuit", dict(nvars=3000, seed=11), ["+ere"]),
    ("ere_gates_w200", "gate_circuit", dict(nvars=2500, seed=31, window=200), ["+ere"]),
    ("ere_ksat3", "random_ksat", dict(nvars=2000, nclauses=8520, k=3, seed=1), ["+ere"]),
    ("ere_xorphp", "xor_php_mix", dict(nclauses=6000, seed=4), ["+ere"]),
    ("ere_extend2", "gate_circuit", dict(nvars=2500, seed=32), ["+ere", "--redundancyextend=2"]),
    ("ere_fuzz", "fuzz_mix", dict(seed=5), ["+ere"]),
    # BCE (-blocked) and -all (BVE + SUB + BCE + ERE forced on, options.cpp:343-345)
    ("bce_gates", "gate_circuit", dict(nvars=2500, seed=41), ["+ere", "-blocked"]),
    ("bce_ksat3", "random_ksat", dict(nvars=1500, nclauses=5400, k=3, seed=4), ["-blocked"]),
    ("all_gates", "gate_circuit", dict(nvars=2500, seed=42), ["+ere", "-all"]),
    # LATER calls: the reference's own first simplify() + 3000 conflicts of its CDCL loop run first
    # (ref_harness --presolve-conflicts), so the store holds learnt clauses and stats.simplify.calls == 2
    ("later_ksat3", "random_ksat", dict(nvars=2000, nclauses=7800, k=3, seed=6), ["--presolve-conflicts=3000"]),
    ("later_gates", "gate_circuit", dict(nvars=8000, seed=7), ["--presolve-conflicts=500"]),
    # incremental solver (Solver() + iadd / itoClause), isolve()'s prologue with a seeded 15 % / 8 % of the variables frozen
    # through ifreeze() and a few assumptions through iassume(): the .in carries the `ifrozen` mask (lcve.cpp:67)
    ("incr_gates", "gate_circuit", dict(nvars=3000, seed=51), ["+ere", "--incr=3:150:6"]),
    ("incr_ksat3", "random_ksat", dict(nvars=2000, nclauses=8000, k=3, seed=9), ["+ere", "--incr=5:80:4"]),
    # prop() with hundreds of units per frontier (units hidden behind self-subsumption, implication layers below them);
    # the second one is not planted: prop() ends in a conflict
    ("unit_storm", "unit_storm", dict(nseeds=400, depth=6, seed=5), ["+ere"]),
    ("unit_storm_unsat", "unit_storm", dict(nseeds=300, depth=5, seed=9, planted=False), []),
]


# cases that also get the clause database + watch table after map(true) (ref_harness --cmmap, a second run of the binary)
CMM_CASES = {"gates_small", "xorphp_small", "later_gates", "fuzz_units4"}

# cases that also get the pre-awaken() clause database (ref_harness --cmin)
CM_CASES = {"gates_small", "ksat5_planted", "xorphp_small", "fuzz_units4", "ere_fuzz", "later_ksat3", "later_gates"}


def main():
    if not os.path.exists(REF):
        raise SystemExit(f"{REF} missing: run `make -C oracle ref` first")
    only = set(sys.argv[1:])                 # optional: regenerate just the named cases
    for name, gen, kw, opts in CASES:
        if only and name not in only:
            continue
        nv, off, lits = getattr(cnfgen, gen)(**kw)
        with tempfile.TemporaryDirectory() as td:
            cnf = os.path.join(td, name + ".cnf")
            cnfgen.write_dimacs(cnf, nv, off, lits)
            fin, fout = os.path.join(HERE, name + ".in"), os.path.join(HERE, name + ".out")
            for f in (fin, fout):
                if os.path.exists(f):
                    os.remove(f)
            base = [] if "+ere" in opts else ["-no-redundancy"]
            extra = []
            if name in CM_CASES:
                fcm = os.path.join(HERE, name + ".cm")
                if os.path.exists(fcm):
                    os.remove(fcm)
                fco = os.path.join(HERE, name + ".cmo")
                if os.path.exists(fco):
                    os.remove(fco)
                extra = [f"--cmin={fcm}", f"--cmout={fco}"]
            r = subprocess.run([REF, cnf, *base, "-quiet", f"--sgin={fin}", f"--sgout={fout}"] + extra + [o for o in opts if o != "+ere"],
                               capture_output=True, text=True)
            print(name, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else f"rc={r.returncode}",
                  os.path.getsize(fin), os.path.getsize(fout))
            if name in CMM_CASES:
                subprocess.run([REF, cnf, *base, "-quiet", f"--cmmap={os.path.join(HERE, name + '.cmm')}"] + [o for o in opts if o != "+ere"],
                               capture_output=True, text=True)


if __name__ == "__main__":
    main()
