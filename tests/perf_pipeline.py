"""Dev aid (not a test): where does the time go when two contexts overlap copies and kernels on one GPU?
Prints the per-stage device times of an execute that ran alone and of one that ran under the other context's copies."""
import json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seqfrost_b200 import sigma, boundary as B, cnfgen

nv, nc = int(os.environ.get("VARS", 2_000_000)), int(os.environ.get("CLAUSES", 10_000_000))
b = B.from_cnf(*cnfgen.planted_subsumption_ksat(nv, nc, k=5, seed=7))
s1, s2 = sigma.Sigma(0), sigma.Sigma(0)
bi1, out1 = s1.alloc_pinned_like(b)
bi2, out2 = s2.alloc_pinned_like(b)
for s, bi, out in ((s1, bi1, out1), (s2, bi2, out2)):
    for _ in range(3):
        s.upload(bi); s.execute(); s.download_into(out)
alone = s1.report()["stage_ms"]
mode = os.environ.get("MODE", "both")
stop = False
def copier():
    while not stop:
        if mode in ("both", "h2d"): s2.upload(bi2)
        if mode in ("both", "d2h"): s2.download_into(out2)
th = threading.Thread(target=copier); th.start()
time.sleep(0.05)
acc = {}
N = 10
for _ in range(N):
    s1.execute()
    for k, v in s1.report()["stage_ms"].items():
        acc[k] = acc.get(k, 0) + v / N
stop = True; th.join()
print(json.dumps({"mode": mode, "alone": {k: round(v, 3) for k, v in alone.items() if v}, "under_copies": {k: round(v, 3) for k, v in acc.items() if v}}))
