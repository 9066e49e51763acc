"""BASELINE.json configs[4] (C5): the batch of mixed instances as a largest-first work queue (seqfrost_b200/batch.py).
CPU: queue order, the shared counter over two gloo ranks (torch.distributed TCPStore), hand-off buffers, digests - with
the serial test stand-in of the launch layer.  GPU (-m gpu): a scaled-down batch drained by two contexts on cuda:0, every
output compared with the unmodified reference; the committed full-size digests are what `bench.py --config c5 --verify`
checks."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import GOLDEN, REF_BIN, ROOT
from seqfrost_b200 import batch, boundary as B

needs_ref = pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/ref_sigma not built")


def test_queue_order_is_largest_first_and_deterministic():
    items = batch.c5_items()
    assert len(items) == 64 and sorted(it["seed"] for it in items) == list(range(100, 164))
    assert all(500_000 <= it["clauses"] <= 5_000_000 for it in items)
    assert all(a["clauses"] >= b["clauses"] for a, b in zip(items, items[1:]))
    assert [it["index"] for it in items] == list(range(64))
    assert {it["kind"] for it in items} == {"ksat5_planted", "gate_circuit", "xor_php"}
    assert items == batch.c5_items()
    small = batch.c5_items(5, scale=0.01)
    assert all(5_000 <= it["clauses"] <= 50_000 for it in small)


def test_committed_c5_digests_cover_the_batch():
    with open(os.path.join(GOLDEN, "c5_digests.json")) as f:
        d = json.load(f)
    assert sorted(d["digests"]) == [str(s) for s in range(100, 164)]
    assert d["input_mismatch_seeds"] == []                 # boundary.from_cnf == the reference's own awaken() on every instance
    assert all(len(v) == 64 or v == "UNSAT" for v in d["digests"].values())


def test_digest_is_the_whole_output():
    bo = B.read(os.path.join(GOLDEN, "gates_w200.out"))
    d = batch.boundary_digest(bo)
    assert d == batch.boundary_digest(B.read(os.path.join(GOLDEN, "gates_w200.out")))
    bo.model = bo.model.copy()
    bo.model[-1] ^= 1
    assert d != batch.boundary_digest(bo)


def _reference_digest(cache, it):
    fout = batch.c5_path(cache, it, "refout")
    subprocess.run([REF_BIN, batch.c5_path(cache, it, "cnf"), "-quiet", f"--sgout={fout}"], capture_output=True, timeout=600, check=True)
    return batch.boundary_digest(B.read(fout))


def _queue_worker(rank, world, port, lib, cache, n, scale, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from seqfrost_b200 import sigma
    store = dist.TCPStore("127.0.0.1", port, world, rank == 0, wait_for_workers=True)
    items = batch.c5_items(n, scale=scale)
    max_bytes = {}
    for it in items:
        for name, (_, nb) in batch.boundary_sections(batch.c5_path(cache, it)).items():
            max_bytes[name] = max(max_bytes.get(name, 0), nb)
    ctxs = [sigma.Sigma(0, lib), sigma.Sigma(0, lib)]
    hos = [batch.HandOff(c, max_bytes) for c in ctxs]
    recs = batch.run_queue(ctxs, hos, items, cache, batch.StoreCounter(store, "q"), digests=True)
    out.put((rank, [(r["index"], r["seed"], r["digest"], r["literals"]) for r in recs]))
    store.add("done", 1)
    while rank == 0 and int(store.add("done", 0)) < world:      # the master keeps the store alive until everyone is through
        pass
    for c in ctxs:
        c.close()


@needs_ref
def test_two_processes_drain_one_queue(emul_lib, tmp_path):
    """world_size 2, two threads each: every instance is taken exactly once, and every output is the reference's."""
    cache, n, scale = str(tmp_path), 7, 0.004
    items = batch.c5_items(n, scale=scale)
    batch.c5_prepare(items, cache, want_cnf=True, workers=2)
    want = {it["seed"]: _reference_digest(cache, it) for it in items}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + os.getpid() % 40
    procs = [ctx.Process(target=_queue_worker, args=(r, 2, port, emul_lib, cache, n, scale, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    taken = sorted(i for recs in got.values() for i, _, _, _ in recs)
    assert taken == list(range(n))                                              # exactly once, between the two processes
    for recs in got.values():
        for _, seed, digest, lits in recs:
            assert digest == want[seed], f"seed {seed}"
            assert lits > 0


@pytest.mark.gpu
@needs_ref
def test_gpu_queue_matches_reference(cuda_lib, tmp_path):
    """A scaled-down C5 batch (25k-250k clauses, all three generators) through two contexts on cuda:0."""
    from seqfrost_b200 import sigma
    cache, n, scale = str(tmp_path), 9, 0.05
    items = batch.c5_items(n, scale=scale)
    batch.c5_prepare(items, cache, want_cnf=True, workers=4)
    max_bytes = {}
    for it in items:
        for name, (_, nb) in batch.boundary_sections(batch.c5_path(cache, it)).items():
            max_bytes[name] = max(max_bytes.get(name, 0), nb)
    ctxs = [sigma.Sigma(0, cuda_lib), sigma.Sigma(0, cuda_lib)]
    hos = [batch.HandOff(c, max_bytes) for c in ctxs]
    recs = batch.run_queue(ctxs, hos, items, cache, batch.LocalCounter(), digests=True)
    assert sorted(r["index"] for r in recs) == list(range(n))
    for r in recs:
        assert r["launches"] > 0
        assert r["digest"] == _reference_digest(cache, items[r["index"]]), f"seed {r['seed']}"
    for c in ctxs:
        c.close()


def test_queue_from_dimacs_files_gives_the_same_outputs(emul_lib, tmp_path):
    """--from-dimacs: file image -> device parser -> extract -> pass; same digests as the boundary hand-off."""
    from seqfrost_b200 import sigma
    cache, n, scale = str(tmp_path), 4, 0.003
    items = batch.c5_items(n, scale=scale)
    batch.c5_prepare(items, cache, want_cnf=True, workers=2)
    max_bytes = {}
    for it in items:
        for name, (_, nb) in batch.boundary_sections(batch.c5_path(cache, it)).items():
            max_bytes[name] = max(max_bytes.get(name, 0), nb)
    tb = max(os.path.getsize(batch.c5_path(cache, it, "cnf")) for it in items)
    ctxs = [sigma.Sigma(0, emul_lib), sigma.Sigma(0, emul_lib)]
    hos = [batch.HandOff(c, max_bytes, text_bytes=tb) for c in ctxs]
    a = batch.run_queue(ctxs, hos, items, cache, batch.LocalCounter(), digests=True, from_dimacs=True)
    b = batch.run_queue(ctxs, hos, items, cache, batch.LocalCounter(), digests=True, from_dimacs=False)
    assert {r["seed"]: r["digest"] for r in a} == {r["seed"]: r["digest"] for r in b} and len(a) == n
    for c in ctxs:
        c.close()
