"""Batch mode over several GPUs: a formula does not shard, so a batch of independent CNF
instances is partitioned over the ranks - one process, one sigma context and one stream per GPU, no
collective on the data path.  torch.distributed is used only to agree on the timing (barrier + max over
ranks) and to add up the work done."""
from __future__ import annotations

import sys


def rank_instances(n_instances, rank, world):
    """Indices of the instances rank `rank` owns: round-robin, so sizes sorted descending spread evenly."""
    return list(range(rank, n_instances, world))


def rank_seed(base_seed, rank):
    """Weak-scaling bench: every rank generates its own instance of the same shape."""
    return base_seed + rank


def max_over_ranks(x, dist=None, device="cpu"):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, dist=None, device="cpu"):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def throughput(units_all_ranks, ms_max_over_ranks):
    """Whole-job rate: the units ALL ranks processed divided by the slowest rank's time."""
    return units_all_ranks / (ms_max_over_ranks / 1e3)


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: a batch of 64 mixed CNF instances (0.5-5M clauses, drawn from
# the C2 / C3 / C4 generators, seeds 100..163) as a WORK QUEUE over the GPUs of one box.  One process per
# GPU; two sigma contexts per process, each driven by its own host thread with its own pinned hand-off
# buffers, so that the copies of one instance run under the kernels of another; every thread takes the next
# instance from one shared counter, largest first.  No collective on the data path.
# ---------------------------------------------------------------------------------------------------
import hashlib
import json
import os
import struct
import threading
import time


def c5_items(n_instances=64, base_seed=100, scale=1.0):
    """The batch, largest first: [{index, seed, clauses, kind}], ties by seed.  scale < 1 shrinks every instance (tests)."""
    import numpy as np
    items = []
    for seed in range(base_seed, base_seed + n_instances):
        ncls = int(np.random.default_rng(seed).integers(500_000, 5_000_001))      # cnfgen.mixed_instance's own draw
        if scale != 1.0:
            ncls = max(1000, int(ncls * scale))
        items.append({"seed": seed, "clauses": ncls, "kind": ("ksat5_planted", "gate_circuit", "xor_php")[seed % 3]})
    items.sort(key=lambda it: (-it["clauses"], it["seed"]))
    for i, it in enumerate(items):
        it["index"] = i
    return items


def c5_path(cache, it, ext="in"):
    return os.path.join(cache, f"c5_{it['seed']}_{it['clauses']}.{ext}")


def c5_prepare_one(job):
    """Generates one instance of the batch into the cache: the boundary file (what awaken() would hand over) and, on
    request, the DIMACS file for the reference arm.  Runs in worker processes."""
    it, cache, want_cnf = job
    seed = it["seed"]
    from seqfrost_b200 import boundary as B, cnfgen
    p_in, p_cnf = c5_path(cache, it), c5_path(cache, it, "cnf")
    if os.path.exists(p_in) and (not want_cnf or os.path.exists(p_cnf)):
        return seed
    nv, off, lits = cnfgen.mixed_instance(seed, it["clauses"])
    if not os.path.exists(p_in):
        b = B.from_cnf(nv, off, lits)
        b.opts["ere_en"] = 1                                   # reference defaults
        tmp = p_in + f".{os.getpid()}.tmp"
        B.write(tmp, b)
        os.replace(tmp, p_in)
    if want_cnf and not os.path.exists(p_cnf):
        tmp = p_cnf + f".{os.getpid()}.tmp"
        cnfgen.write_dimacs(tmp, nv, off, lits)
        os.replace(tmp, p_cnf)
    return seed


def c5_prepare(items, cache, rank=0, world=1, want_cnf=False, workers=None):
    """Every rank generates its round-robin share of the missing instances (in worker processes)."""
    from concurrent.futures import ProcessPoolExecutor
    os.makedirs(cache, exist_ok=True)
    mine = [(it, cache, want_cnf) for it in items[rank::world]
            if not (os.path.exists(c5_path(cache, it)) and (not want_cnf or os.path.exists(c5_path(cache, it, "cnf"))))]
    if not mine:
        return 0
    workers = workers or max(1, min(len(mine), (len(os.sched_getaffinity(0)) // max(1, world)), 16))
    with ProcessPoolExecutor(max_workers=workers) as ex:
        list(ex.map(c5_prepare_one, mine))
    return len(mine)


_SECTION_DTYPES = {"state": "uint8", "value": "int8", "trail": "uint32", "vorg": "uint32", "model": "uint32", "arena": "uint32",
                   "refs": "uint64", "ifrozen": "uint8"}


def boundary_sections(path):
    """{section: (file offset, bytes)} of a boundary file, without reading the payload."""
    out = {}
    with open(path, "rb") as f:
        from seqfrost_b200 import boundary as B
        if f.read(8) != B.MAGIC:
            raise ValueError(f"{path}: not a sigma boundary file")
        f.seek(0, 2)
        end = f.tell()
        pos = 8
        while pos < end:
            f.seek(pos)
            name = f.read(8).rstrip(b"\0").decode()
            (n,) = struct.unpack("<Q", f.read(8))
            out[name] = (pos + 16, n)
            pos += 16 + n
    return out


class HandOff:
    """The pinned hand-off buffers of one host thread: input arrays sized for the largest instance of the batch, output
    buffers with 25 % slack (what an embedding solver owns once, sigma_host_alloc)."""

    def __init__(self, sig, max_bytes, slack=1.25, text_bytes=0):
        import numpy as np
        from seqfrost_b200 import sigma as S
        self.np = np
        self.keep = {}
        self.inp = {}
        self.text = None
        if text_bytes:                                     # --from-dimacs: the file image instead of the boundary arrays
            self.keep["text"] = S.PinnedArray(sig.lib, np.dtype(np.uint8), text_bytes + 64)
            self.text = self.keep["text"].array
        for name, dt in _SECTION_DTYPES.items():
            n = max_bytes.get(name, 0) // np.dtype(dt).itemsize + 16      # trail / model are empty on a first call
            if name == "ifrozen":
                continue
            self.keep[name] = S.PinnedArray(sig.lib, np.dtype(dt), n)
            self.inp[name] = self.keep[name].array
        a, r = self.inp["arena"].size, self.inp["refs"].size
        nv1 = self.inp["state"].size
        caps = dict(arena=("uint32", int(a * slack) + 1024), refs=("uint64", int(r * slack) + 1024), state=("uint8", nv1),
                    value=("int8", 2 * nv1), trail=("uint32", nv1 + 16), model=("uint32", a + 1024))
        self.out = {}
        for name, (dt, n) in caps.items():
            self.keep["out_" + name] = S.PinnedArray(sig.lib, np.dtype(dt), n)
            self.out[name] = self.keep["out_" + name].array

    def load_text(self, path):
        """Reads a DIMACS file into the pinned text buffer; returns the view of its bytes."""
        n = os.path.getsize(path)
        view = self.text[:n]
        with open(path, "rb") as f:
            if f.readinto(memoryview(view)) != n:
                raise IOError(f"{path}: short read")
        return view

    def load(self, path):
        """Reads a boundary file straight into the pinned input arrays; returns the Boundary (views) of that instance."""
        from seqfrost_b200 import boundary as B
        np = self.np
        sec = boundary_sections(path)
        b = B.Boundary()
        with open(path, "rb") as f:
            f.seek(sec["meta"][0])
            meta = np.frombuffer(f.read(sec["meta"][1]), dtype=np.uint64)
            b.status, b.nvars, b.n_clauses, b.n_literals = int(meta[0]), int(meta[1]), int(meta[2]), int(meta[3])
            b.unassigned, b.max_frozen, b.max_melted = int(meta[4]), int(meta[5]), int(meta[6])
            b.propagated, b.simplify_calls = int(meta[8]), int(meta[9])
            for name, dt in _SECTION_DTYPES.items():
                if name not in sec or name == "ifrozen":
                    continue
                off, nbytes = sec[name]
                view = self.inp[name][:nbytes // np.dtype(dt).itemsize]
                f.seek(off)
                got = f.readinto(memoryview(view).cast("B")) if nbytes else 0
                if got != nbytes:
                    raise IOError(f"{path}: short read of section {name}")
                setattr(b, name, view)
            if "opts" in sec:
                f.seek(sec["opts"][0])
                o = np.frombuffer(f.read(sec["opts"][1]), dtype=np.int32)
                b.opts = {k: int(o[i]) for i, k in enumerate(B.OPT_NAMES)}
        return b


def output_digest(arena, refs, state, value, trail, model, scalars):
    """sha256 over the pass output, word for word (arena, refs, state & 7, value, trail in order, model stack, counters)."""
    import numpy as np
    h = hashlib.sha256()
    h.update(np.asarray(scalars, dtype=np.uint64).tobytes())
    for a in (arena, refs, state & np.uint8(7), value, trail, model):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def boundary_digest(b):
    return output_digest(b.arena, b.refs, b.state, b.value, b.trail, b.model,
                         [b.n_clauses, b.n_literals, b.unassigned, b.max_frozen, b.max_melted, b.propagated])


class LocalCounter:
    def __init__(self):
        self._n, self._lock = 0, threading.Lock()

    def next(self):
        with self._lock:
            self._n += 1
            return self._n - 1


class StoreCounter:
    """The shared counter of a multi-process queue: an atomic add on a torch.distributed TCPStore key."""

    def __init__(self, store, key):
        self.store, self.key = store, key

    def next(self):
        return int(self.store.add(self.key, 1)) - 1


def run_queue(ctxs, handoffs, items, cache, counter, digests=None, from_dimacs=False):
    """Drains the queue with one host thread per (context, hand-off) pair.  Returns the per-instance records of THIS process:
    [{index, seed, literals, device_ms, h2d_bytes, d2h_bytes, digest?}], plus the aggregated report arrays."""
    done, errs = [], []
    lock = threading.Lock()

    def worker(sig, ho):
        try:
            while True:
                i = counter.next()
                if i >= len(items):
                    return
                it = items[i]
                parse_ms = 0.0
                if from_dimacs:
                    # file image -> parser -> extract -> pass, all on the device (sigma_parse_dimacs, sigma_upload_parsed)
                    d = sig.parse_dimacs(ho.load_text(c5_path(cache, it, "cnf")), download=False)
                    sig.upload_parsed({"ere_en": 1}, nvars=d["nvars"])
                    parse_ms = d["device_ms"]
                    b = sig._keep[0]
                else:
                    b = ho.load(c5_path(cache, it))
                    sig.upload(b)
                sig.execute()
                f = sig.download_into(ho.out)
                r = sig.report()
                rec = {"index": i, "seed": it["seed"], "literals": int(r["literals_in"]), "device_ms": r["stage_ms"]["total_device"],
                       "h2d_bytes": r["h2d_bytes"] + (os.path.getsize(c5_path(cache, it, "cnf")) if from_dimacs else 0),
                       "d2h_bytes": r["d2h_bytes"], "launches": r["kernel_launches"], "parse_ms": parse_ms,
                       "stage_ms": r["stage_ms"], "stage_bytes": r["stage_bytes"], "stage_launches": r["stage_launches"]}
                if digests is not None:
                    o = ho.out
                    rec["digest"] = output_digest(o["arena"][:f.n_buckets], o["refs"][:f.n_clauses], o["state"][:b.nvars + 1],
                                                  o["value"][:2 * b.nvars + 2], o["trail"][:f.trail_size], o["model"][:f.model_size],
                                                  [f.n_clauses, f.n_literals, f.unassigned, f.max_frozen, f.max_melted, f.propagated])
                with lock:
                    done.append(rec)
        except Exception as e:      # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=worker, args=(s, h)) for s, h in zip(ctxs, handoffs)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise errs[0]
    return done


def bench_c5(args, rank, world, local):
    """`bench.py --config c5`: instances/s and literals/s of the 64-instance queue at N GPUs (ours), or of
    min(64, host cores) reference processes draining the same queue (--impl reference)."""
    import bench as Bn                                     # helpers of the bench script (clocks, reference runner)
    cache = Bn.CACHE
    items = c5_items(args.instances)
    golden = os.path.join(Bn.ROOT, "tests", "golden", "c5_digests.json")
    if args.impl == "reference":
        if rank != 0:
            return
        return _bench_c5_reference(args, Bn, items, cache)

    import numpy as np
    import torch
    from seqfrost_b200 import sigma
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = local
    torch.cuda.set_device(dev)
    pinned_to = Bn.pin_to_gpu_numa_node(torch, dev)
    from_dimacs = bool(getattr(args, "from_dimacs", False))
    c5_prepare(items, cache, rank, world, want_cnf=from_dimacs)
    device = f"cuda:{dev}"

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
    barrier()                                               # every rank's share of the cache is written
    max_bytes = {}
    for it in items:
        for name, (_, n) in boundary_sections(c5_path(cache, it)).items():
            max_bytes[name] = max(max_bytes.get(name, 0), n)
    ctxs = [sigma.Sigma(dev), sigma.Sigma(dev)]
    text_bytes = max(os.path.getsize(c5_path(cache, it, "cnf")) for it in items) if from_dimacs else 0
    handoffs = [HandOff(ctxs[0], max_bytes, text_bytes=text_bytes), HandOff(ctxs[1], max_bytes, text_bytes=text_bytes)]
    store = None
    if world > 1:
        port = int(os.environ.get("MASTER_PORT", "29500")) + 17
        store = torch.distributed.TCPStore(os.environ.get("MASTER_ADDR", "127.0.0.1"), port, world, rank == 0, wait_for_workers=True)

    def one_pass(tag, digests=None):
        counter = LocalCounter() if store is None else StoreCounter(store, f"c5_{tag}")
        barrier()
        t0 = time.perf_counter()
        recs = run_queue(ctxs, handoffs, items, cache, counter, digests, from_dimacs)
        torch.cuda.synchronize()
        mine_ms = 1e3 * (time.perf_counter() - t0)
        barrier()
        wall_ms = max_over_ranks(1e3 * (time.perf_counter() - t0), dist, device)
        return recs, wall_ms, mine_ms

    verified = None
    if args.verify:
        with open(golden) as f:
            want = json.load(f)["digests"]
        recs, _, _ = one_pass("verify", digests=True)
        bad = [r["seed"] for r in recs if want.get(str(r["seed"])) != r["digest"]]
        nbad = sum_over_ranks(len(bad), dist, device)
        nchk = sum_over_ranks(len(recs), dist, device)
        verified = {"checked": int(nchk), "mismatches": int(nbad), "against": "tests/golden/c5_digests.json (unmodified reference, sha256 of the whole output)"}
        if bad:
            print(f"rank {rank}: digest mismatch for seeds {bad}", file=sys.stderr)
    for w in range(args.warmup):
        one_pass(f"w{w}")
    clk = Bn.clocks_start(dev) if rank == 0 else None
    walls, all_recs = [], []
    for k in range(args.steps):
        recs, wall_ms, _ = one_pass(f"s{k}")
        walls.append(wall_ms)
        all_recs += recs
    clocks = Bn.clocks_stop(clk) if rank == 0 else None
    total_ms = sum(walls)
    lits = sum_over_ranks(sum(r["literals"] for r in all_recs), dist, device)
    ninst = sum_over_ranks(len(all_recs), dist, device)
    dev_ms = max_over_ranks(sum(r["device_ms"] for r in all_recs), dist, device)          # the busiest GPU's kernel time
    launches = sum_over_ranks(sum(r["launches"] for r in all_recs), dist, device)
    h2d = sum_over_ranks(sum(r["h2d_bytes"] for r in all_recs), dist, device)
    d2h = sum_over_ranks(sum(r["d2h_bytes"] for r in all_recs), dist, device)
    per_rank = [0] * world
    per_rank[rank] = len(all_recs) // max(1, args.steps)
    if dist is not None:
        t = torch.tensor(per_rank, dtype=torch.int64, device=device)
        dist.all_reduce(t)
        per_rank = t.tolist()
    # roofline over the streaming stages of every instance this rank processed (rank 0's share)
    peaks = {}
    try:
        with open(os.path.join(Bn.ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    sb = {k: sum(r["stage_bytes"].get(k, 0) for r in all_recs) for k in sigma.STREAMING_STAGES}
    sm = {k: sum(r["stage_ms"].get(k, 0.0) for r in all_recs) for k in sigma.STREAMING_STAGES}
    sl = {k: sum(r["stage_launches"].get(k, 0) for r in all_recs) for k in sigma.STREAMING_STAGES}
    streaming = [k for k in sigma.STREAMING_STAGES if sb[k] and sm[k]]
    single = sigma.SINGLE_KERNEL_STAGES
    dom = max([k for k in streaming if k in single], key=lambda k: sm[k])
    ach = sb[dom] / 1e9 / (sm[dom] / 1e3)
    tb, tm = sum(sb[k] for k in streaming), sum(sm[k] for k in streaming)
    roof = {"bound": "hbm", "kernel": single[dom], "stage": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": None, "peak_source": peak_src, "launches_timed": sl[dom],
            "note": "dominant single streaming kernel by time over rank 0's share of the batch",
            "all_streaming_aggregate": {"GB/s": tb / 1e9 / (tm / 1e3), "frac": tb / 1e9 / (tm / 1e3) / peak}}
    if rank == 0:
        value = lits / (total_ms / 1e3)
        line = {
            "metric": Bn.METRIC, "value": value, "unit": Bn.UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"batch of {len(items)} mixed CNF instances (C2/C3/C4 generators, 0.5-5M clauses, seeds 100..{99 + len(items)}) "
                                   "as a largest-first work queue, reference default options (BASELINE.json configs[4], SURVEY 8e C5)",
                       "config": "c5", "instances_per_step": len(items),
                       "input": "DIMACS files: parsed, extracted and simplified on the device (sigma_parse_dimacs -> sigma_upload_parsed)" if from_dimacs
                                else "the SCNF boundary awaken() hands over (sigma_upload)", "queue": "shared counter (torch.distributed TCPStore add), two contexts + two host threads per GPU",
                       "instances_per_rank": per_rank, "l2": "inputs larger than L2 (every instance is 40-450 MB)",
                       "value_is": "whole-queue wall-clock rate = e2e (two contexts share each GPU, so the passes' device times overlap "
                                   "and do not add up to a device-only figure; their sum on the busiest GPU is device_ms_busiest_gpu_per_step)"},
            "instances_per_s": ninst / (total_ms / 1e3),
            "roofline": roof,
            "e2e": {"value": value, "unit": Bn.UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                    "instances_per_s": ninst / (total_ms / 1e3), "ms_per_step": total_ms / args.steps,
                    "what": "wall clock of draining the whole queue: read from the page cache into pinned buffers, H2D, pass, D2H, max over ranks"},
            "device_ms_busiest_gpu_per_step": dev_ms / args.steps,
            "gpu_launches": int(launches), "clocks": clocks, "verified": verified,
            "cpu_baseline": None, "host": {"cpus": len(Bn.host_cpus()), "cpu_model": Bn.cpu_model(), "pinned_to": pinned_to},
        }
        print(json.dumps(line))
    for c in ctxs:
        c.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _bench_c5_reference(args, Bn, items, cache):
    """The reference draining the same queue: min(instances, host cores) worker slots, each running one unmodified
    single-threaded reference process at a time (largest first), pinned to its own core."""
    if not os.path.exists(Bn.REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_sigma not built (make -C oracle ref needs /root/reference)"}))
        return
    c5_prepare(items, cache, 0, 1, want_cnf=True)
    cpus = Bn.host_cpus()
    slots = min(len(items), len(cpus))
    lock, state = threading.Lock(), {"next": 0, "lits": 0, "n": 0, "pass_ms": 0.0}
    errs = []

    def slot(cpu):
        try:
            while True:
                with lock:
                    i = state["next"]
                    state["next"] += 1
                if i >= len(items):
                    return
                r = Bn.ref_result(Bn.ref_sigma(c5_path(cache, items[i], "cnf"), ere=True, cpu=cpu))
                with lock:
                    state["lits"] += int(r.get("literals_in", 0)); state["n"] += 1; state["pass_ms"] += r["pass_ms"]
        except Exception as e:      # noqa: BLE001
            errs.append(e)
    t0 = time.perf_counter()
    th = [threading.Thread(target=slot, args=(cpus[k],)) for k in range(slots)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    wall_s = time.perf_counter() - t0
    if errs:
        raise errs[0]
    value = state["lits"] / (state["pass_ms"] / 1e3 / slots)      # pass-only rate of the `slots` cores (parse time excluded)
    line = {"impl": "reference", "metric": Bn.METRIC, "value": value, "unit": Bn.UNIT, "n_gpus": args.gpus, "steps": 1, "warmup": 0,
            "ms_per_step": state["pass_ms"] / slots, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": f"batch of {len(items)} mixed CNF instances as a largest-first queue (C5)", "config": "c5", "same_config": True,
                       "instances_per_step": len(items), "wall_s_including_parse": wall_s},
            "instances_per_s": state["n"] / (state["pass_ms"] / 1e3 / slots),
            "cpu_baseline": {"value": value, "unit": Bn.UNIT, "cores": slots, "kind": "reference",
                             "sample": f"all {len(items)} instances once, {slots} single-threaded reference processes at a time, region "
                                       f"simplify.cpp:180-206 (sum of pass ms / {slots} cores), {Bn.cpu_model()}, {os.cpu_count()} host cores"},
            "e2e": {"value": value, "unit": Bn.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
