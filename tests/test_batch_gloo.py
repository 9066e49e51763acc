"""N>1 host logic on CPU: two gloo ranks each own independent instances (no data-path collective), run them
through the pass (serial test stand-in of the launch layer), and agree on max-over-ranks time and summed work."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT


def _worker(rank, world, port, emul_lib, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from seqfrost_b200 import batch, boundary as B, sigma
    import parity
    names = ["gates_w200", "ksat3_small", "fuzz_units4", "xorphp_small", "ere_gates"]
    mine = batch.rank_instances(len(names), rank, world)
    s = sigma.Sigma(0, emul_lib)
    lits, ok = 0, True
    for i in mine:
        bi, bo = B.read(os.path.join(GOLDEN, names[i] + ".in")), B.read(os.path.join(GOLDEN, names[i] + ".out"))
        o, _ = s.simplify(bi)
        ok = ok and parity.diff(o, bo) == []
        lits += bi.n_literals
    s.close()
    dist.barrier()
    fake_ms = 10.0 * (rank + 1)
    mx = batch.max_over_ranks(fake_ms, dist)
    total = batch.sum_over_ranks(lits, dist)
    allok = batch.sum_over_ranks(1.0 if ok else 0.0, dist)
    if rank == 0:
        out.put((mine, mx, total, allok, batch.throughput(total, mx)))
    dist.destroy_process_group()


def test_two_ranks_partition_a_batch(emul_lib):
    from seqfrost_b200 import batch, boundary as B
    assert batch.rank_instances(5, 0, 2) == [0, 2, 4] and batch.rank_instances(5, 1, 2) == [1, 3]
    assert batch.rank_seed(7, 3) == 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, emul_lib, q)) for r in range(2)]
    for p in procs:
        p.start()
    mine0, mx, total, allok, thr = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    names = ["gates_w200", "ksat3_small", "fuzz_units4", "xorphp_small", "ere_gates"]
    expect = sum(B.read(os.path.join(GOLDEN, n + ".in")).n_literals for n in names)
    assert mine0 == [0, 2, 4] and mx == 20.0 and total == expect and allok == 2.0
    assert thr == pytest.approx(expect / 0.02)
