#!/usr/bin/env python
"""bench.py - SIGmA pass throughput on B200 (BASELINE.json metric: pass ms and literals/s per GPU,
HBM GB/s vs peak, instances/s at 1/2/4/8 GPUs).

A "step" is one whole SIGmA pass (reference simplify.cpp:180-210: every phase of createOT, prop,
LCVE, sortOT, SUB, VE, ERE, counts, compaction, export) over one synthetic instance.

  --config c2 (default)  BASELINE.json configs[1]: random 5-SAT, 2M variables / 10M clauses with planted
                         subsumption work - the configuration the metric is quoted on
  --config c1 | c3 | c4  configs[0], [2], [3] at their stated sizes (same JSON shape; results under profiles/)
  --config u1            prop() stress case (cnfgen.unit_storm), not a BASELINE config
  --config c5            configs[4]: the batch of 64 mixed instances as a largest-first work queue (seqfrost_b200/batch.py)

At N>1 every rank owns an independent instance of the same shape (different seed): the formula does not
shard, so there is no collective on the data path (scaling = weak); torch.distributed is used only for the
barrier and the max-over-ranks reduce.

  value : literals/s with the formula already resident in HBM (sigma_execute only), timed with
          the library's CUDA events on its own stream, max over ranks
  e2e   : the same through the C ABI with pinned HOST buffers, ONE context, blocking like the solver's own call:
          sigma_upload (H2D) + sigma_execute + sigma_download (D2H) inside the timed region.
          e2e.pipelined is the batch-mode figure (two contexts per GPU, copies of one instance under the kernels
          of the other)
  binding_e2e : "Simplifier time" of `seqfrost_sigma <cnf> -no-solve` (reference solver + this pass, pageable solver
          arenas pinned for the copies) next to the reference's own whole simplifying() call on the same file
  --impl reference : the unmodified reference (oracle/_ref/ref_sigma, single-threaded like SeqFROST itself) on the
          host cores, on the FULL instance of the config (same_config), one process per rank's instance at --gpus N
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_sigma")
HOST_BIN = os.path.join(ROOT, "seqfrost_b200", "host", "_build", "seqfrost_sigma")
METRIC, UNIT = "sigma_pass_literals_per_s", "literals/s"
CACHE = os.environ.get("SIGMA_BENCH_CACHE", "/tmp/seqfrost_b200_bench")
REF_BUDGET_S = float(os.environ.get("SIGMA_REF_BUDGET_S", "240"))     # wall-clock cap of the reference arm's timed steps

CONFIGS = {
    "c1": dict(gen="random_ksat", kw=dict(nvars=100_000, nclauses=426_000, k=3, seed=1),
               name="random 3-SAT 100k vars / 426k clauses, ratio 4.26 (BASELINE.json configs[0], SURVEY 8d C1)"),
    "c2": dict(gen="planted_subsumption_ksat", kw=dict(nvars=2_000_000, nclauses=10_000_000, k=5, seed=7),
               name="random 5-SAT 2M vars / 10M clauses + planted subsumption (BASELINE.json configs[1], SURVEY 8d C2)"),
    "c3": dict(gen="gate_circuit", kw=dict(nvars=5_000_000, seed=11),
               name="BMC-style gate circuit (AND/XOR/ITE/equivalence), 5M vars (BASELINE.json configs[2], SURVEY 8d C3)"),
    "c4": dict(gen="xor_php_mix", kw=dict(nclauses=1_000_000, seed=3),
               name="XOR chains + pigeonhole mix, 1M clauses (BASELINE.json configs[3], SURVEY 8d C4)"),
    # not a BASELINE config: the prop() stress case (propelim.cpp:46-81), 810k units through one prop() call
    "u1": dict(gen="unit_storm", kw=dict(nseeds=100_000, depth=8, seed=5),
               name="unit storm: 100k units hidden behind self-subsumption, 8 implication layers, 1M vars / 2.6M clauses"),
}


def config_kw(args, rank=0):
    from seqfrost_b200 import batch
    c = CONFIGS[args.config]
    kw = dict(c["kw"])
    if args.config == "c2":
        kw.update(nvars=args.vars, nclauses=args.clauses)
    kw["seed"] = batch.rank_seed(args.seed if args.seed is not None else kw["seed"], rank)
    return c["gen"], kw


def workload(args, rank=0):
    from seqfrost_b200 import cnfgen
    gen, kw = config_kw(args, rank)
    return getattr(cnfgen, gen)(**kw)


def workload_name(args):
    name = CONFIGS[args.config]["name"]
    if args.config == "c2" and (args.vars, args.clauses) != (2_000_000, 10_000_000):
        name = f"random 5-SAT {args.vars / 1e6:g}M vars / {args.clauses / 1e6:g}M clauses + planted subsumption (C2 generator, reduced size)"
    return name + ", " + ("-no-redundancy" if args.no_ere else "reference default options (ERE on)")


def dimacs_file(args, rank=0, inst=None):
    """The DIMACS file of rank `rank`'s instance, written once per (config, size, seed) under CACHE."""
    from seqfrost_b200 import cnfgen
    gen, kw = config_kw(args, rank)
    key = gen + "_" + "_".join(f"{k}{v}" for k, v in sorted(kw.items()))
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, key + ".cnf")
    if not os.path.exists(path):
        nv, off, lits = inst if inst is not None else getattr(cnfgen, gen)(**kw)
        tmp = path + f".{os.getpid()}.tmp"
        cnfgen.write_dimacs(tmp, nv, off, lits)
        os.replace(tmp, path)
    return path


REF_LINE = re.compile(r"(\w+)=([0-9.]+)")


def ref_sigma(cnf, ere=True, cpu=None, full=False):
    """One run of the unmodified reference on `cnf`; returns the numbers its harness prints (pass_ms = the replaced
    region simplify.cpp:180-206; with full=True also awaken / newBeginning / rebuildWT = the whole simplifying())."""
    cmd = [REF_BIN, cnf, "-quiet"] + ([] if ere else ["-no-redundancy"]) + (["--full-timing"] if full else [])
    if cpu is not None:
        cmd = ["taskset", "-c", str(cpu)] + cmd
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def ref_result(p):
    out, err = p.communicate()
    d = {}
    for l in out.splitlines():
        if l.startswith("ref_sigma:"):
            d.update({k: float(v) for k, v in REF_LINE.findall(l)})
    if "pass_ms" not in d:
        raise RuntimeError("ref_sigma failed: " + out[-300:] + err[-300:])
    return d


def host_cpus():
    return sorted(os.sched_getaffinity(0))


def pin_to_gpu_numa_node(torch, dev):
    """Run this rank's host threads (and first-touch its pinned buffers) on the CPUs of the NUMA node the GPU hangs off:
    with 8 ranks, 0.8 GB per step and rank crosses the host, and remote-node pinned memory halves the copy rate.
    Returns the cpulist used, or None when the topology cannot be read."""
    try:
        pr = torch.cuda.get_device_properties(dev)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return spec
    except (OSError, AttributeError, ValueError):
        return None


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for l in f:
                if l.startswith("model name"):
                    return l.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def clocks_start(dev):
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        return subprocess.Popen(["nvidia-smi", "-i", str(dev), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        return None


def clocks_stop(p):
    if p is None:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    p.terminate()
    try:
        out = p.communicate(timeout=5)[0]
    except Exception:
        p.kill()
        out = ""
    sm, mx, reasons = [], None, set()
    for line in out.splitlines():
        f = [x.strip() for x in line.split(",")]
        if len(f) < 7:
            continue
        try:
            sm.append(float(f[0])); mx = float(f[1])
        except ValueError:
            continue
        for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
            if val.lower().startswith("active"):
                reasons.add(name)
    return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world):
    """The reference arm: the unmodified reference's own pass (simplify.cpp:180-206) on the FULL instance of the config.
    SeqFROST is single-threaded, so "all the host threads it can use" is one per process; at --gpus N rank 0 starts N
    reference processes, each pinned to its own core, one per instance of the N-GPU batch (the same DIMACS file: every
    rank's instance has the same shape), and reports their aggregate.  A step is one such run; the number of timed steps is
    capped so that the arm ends within REF_BUDGET_S (there is no warm-up to speak of on a CPU: --warmup runs are skipped
    once a single run takes more than a few seconds)."""
    if rank != 0:
        return
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_sigma not built (make -C oracle ref needs /root/reference)"}))
        return
    cnf = dimacs_file(args)
    cpus = host_cpus()
    nproc = min(args.gpus, len(cpus))
    steps_ms, lits_in, t_begin = [], 0, time.perf_counter()
    n_warm = args.warmup
    i = 0
    while True:
        t0 = time.perf_counter()
        procs = [ref_sigma(cnf, ere=not args.no_ere, cpu=cpus[(k * max(1, len(cpus) // nproc)) % len(cpus)]) for k in range(nproc)]
        res = [ref_result(p) for p in procs]
        run_s = time.perf_counter() - t0
        lits_in = int(res[0]["literals_in"]) if "literals_in" in res[0] else lits_in
        step_ms = max(r["pass_ms"] for r in res)
        if i == 0 and run_s > 5.0:
            n_warm = 0                                   # a run this long is its own warm-up
        if i >= n_warm:
            steps_ms.append(step_ms)
        i += 1
        elapsed = time.perf_counter() - t_begin
        if len(steps_ms) >= args.steps or (steps_ms and elapsed + run_s > REF_BUDGET_S):
            break
    if not lits_in:
        from seqfrost_b200 import boundary as B  # noqa: F401
        nv, off, lits = workload(args)
        lits_in = int(off[-1])
    total_ms = sum(steps_ms)
    value = nproc * lits_in * len(steps_ms) / (total_ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / len(steps_ms), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args), "same_config": True, "literals_per_instance": lits_in, "instances_per_step": nproc,
                   "steps_timed": len(steps_ms), "warmup_run": n_warm,
                   "capped": None if len(steps_ms) >= args.steps else f"timed steps capped at {len(steps_ms)} to stay within {REF_BUDGET_S:.0f} s"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": "reference",
                         "sample": f"the full instance ({lits_in} literals), {nproc} process(es) of the unmodified reference pinned to "
                                   f"{nproc} core(s), region simplify.cpp:180-206, {cpu_model()}, {os.cpu_count()} host cores "
                                   f"(SeqFROST is single-threaded)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def binding_e2e(cnf, ref):
    """`seqfrost_sigma <cnf> -no-solve -report`: the reference CLI with this pass behind it, on the reference's own
    (pageable, malloc'ed) arenas - what a maintainer who links the binding sees as "Simplifier time"."""
    if not os.path.exists(HOST_BIN):
        return {"unavailable": "seqfrost_b200/host/_build/seqfrost_sigma not built (needs the reference headers)"}
    out = {}
    for tag, env in (("pageable", {}), ("pinned_for_copies", {"SIGMA_HOST_REGISTER": "1"})):
        runs = []
        for _ in range(2):          # two process starts: the child creates its CUDA context while it parses, and on a short file the simplifier waits for that
            r = subprocess.run([HOST_BIN, cnf, "-no-solve", "-report"], capture_output=True, text=True, env={**os.environ, **env})
            m = re.search(r"Simplifier time\s*:\s*(?:\x1b\[[0-9;]*m)*([0-9.]+)", r.stdout)
            if not m:
                return {"unavailable": "could not parse Simplifier time: " + (r.stdout[-200:] + r.stderr[-200:]).replace("\n", " ")}
            runs.append(float(m.group(1)))
        out[tag] = {"simplifier_s": min(runs), "runs_s": runs}
    out["what"] = ("timer.simplify of Solver::simplifying() (simplify.cpp:170,230): awaken/extract, pass, newBeginning, rebuildWT, "
                   "retrail; ours = seqfrost_sigma (device extract + pass + device newBeginning / map + device watch table, host retrail; best of two process starts), "
                   "reference = the same call of the unmodified reference")
    if ref and "simplifier_ms" in ref:
        out["reference_simplifier_s"] = ref["simplifier_ms"] / 1e3
        out["reference_parts_ms"] = {k: ref[k] for k in ("awaken_ms", "pass_ms", "final_compact_ms", "newbeginning_ms", "rebuildwt_ms") if k in ref}
        out["speedup"] = out["reference_simplifier_s"] / out["pageable"]["simplifier_s"]     # the binding's default: arenas stay pageable
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS) + ["c5"])
    ap.add_argument("--vars", type=int, default=2_000_000, help="c2 only")
    ap.add_argument("--clauses", type=int, default=10_000_000, help="c2 only")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-binding", action="store_true")
    ap.add_argument("--no-ere", action="store_true", help="run both sides with -no-redundancy")
    ap.add_argument("--instances", type=int, default=64, help="c5 only: size of the batch")
    ap.add_argument("--from-dimacs", action="store_true", help="c5 only: the queue starts from the DIMACS files (device parser + extract + pass)")
    ap.add_argument("--verify", action="store_true", help="c5 only: check every output against tests/golden/c5_digests.json")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.config == "c5":
        from seqfrost_b200 import batch
        return batch.bench_c5(args, rank, world, local)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    from seqfrost_b200 import boundary as B, sigma, batch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local if world > 1 else 0
    numa = pin_to_gpu_numa_node(torch, dev) if world > 1 else None      # host buffers next to the GPU they feed

    # ---- instance (synthetic, seeded; one independent instance per rank) ------------------------
    inst = workload(args, rank)
    nv, off, lits = inst
    b = B.from_cnf(nv, off, lits)
    b.opts["ere_en"] = 0 if args.no_ere else 1          # reference default: ERE on (options.cpp:28)
    cnf = dimacs_file(args, 0, inst) if (rank == 0 and world == 1 and not (args.no_cpu_baseline and args.no_binding)) else None
    del off, lits, inst
    s = sigma.Sigma(dev)
    bi, bufs = s.alloc_pinned_like(b)
    lits_in = int(b.n_literals)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        return batch.max_over_ranks(x, dist, f"cuda:{dev}")

    # ---- device-resident arm: sigma_execute on the uploaded formula ------------------------------
    s.upload(bi)
    for _ in range(args.warmup):
        s.execute()
    clk = clocks_start(dev)
    barrier()
    dev_ms, launches, stage_ms, stage_bytes, stage_launches, rep = 0.0, 0, {}, {}, {}, None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.execute()
        rep = s.report()
        dev_ms += rep["stage_ms"]["total_device"]
        launches += rep["kernel_launches"]
        for k, v in rep["stage_ms"].items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v
        for k, v in rep["stage_bytes"].items():
            stage_bytes[k] = stage_bytes.get(k, 0) + v
        for k, v in rep["stage_launches"].items():
            stage_launches[k] = stage_launches.get(k, 0) + v
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = clocks_stop(clk)
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)

    # ---- end-to-end arm: pinned host buffers, H2D + pass + D2H per step ---------------------------
    # (a) the blocking call the solver makes: one context, upload -> execute -> download, step after step
    for _ in range(min(args.warmup, 2)):
        s.upload(bi); s.execute(); s.download_into(bufs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s.upload(bi)
        s.execute()
        s.download_into(bufs)
    barrier()
    e2e_serial_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    rep2 = s.report()
    # (b) batch mode: two contexts per GPU (the ABI's unit of concurrency), one host thread each, taking
    # alternate steps, so that the PCIe copies of one instance overlap the kernels of the other.  Every step
    # still uploads its input from pinned host memory and downloads its result.
    import threading
    s2 = sigma.Sigma(dev)
    _, bufs2 = s2.alloc_pinned_like(b)
    for ctx_, out_ in ((s, bufs), (s2, bufs2)):
        ctx_.upload(bi); ctx_.execute(); ctx_.download_into(out_)
    errs = []
    gpu_turn = threading.Lock()     # one instance computes at a time; the other one's copies run under it

    def worker(ctx_, out_, n):
        try:
            for _ in range(n):
                ctx_.upload(bi)
                with gpu_turn:
                    ctx_.execute()
                ctx_.download_into(out_)
        except Exception as e:      # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=worker, args=(s, bufs, (args.steps + 1) // 2)),
          threading.Thread(target=worker, args=(s2, bufs2, args.steps // 2))]
    barrier()
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    barrier()
    e2e_pipe_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    if errs:
        raise errs[0]
    same = all(np.array_equal(bufs[k][:n_], bufs2[k][:n_]) for k, n_ in (("arena", 1024), ("refs", 1024))) if args.steps > 1 else True
    s2.close()

    total_lits = batch.sum_over_ranks(lits_in, dist, f"cuda:{dev}") * args.steps     # all ranks' instances
    value = batch.throughput(total_lits, dev_ms)
    # ---- roofline of the dominant streaming kernel -----------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    # stages that are ONE kernel per launch group (its CUDA-event time / launches = that kernel's launch duration)
    single = sigma.SINGLE_KERNEL_STAGES
    streaming = [k for k in sigma.STREAMING_STAGES if stage_bytes.get(k) and stage_ms.get(k)]
    dom = max([k for k in streaming if k in single], key=lambda k: stage_ms[k])
    dom_launches = max(1, stage_launches[dom])
    ach = stage_bytes[dom] / 1e9 / (stage_ms[dom] / 1e3)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            traffic = json.load(f)["traffic_bytes_per_launch"].get(single[dom])
    except (OSError, KeyError, ValueError):
        pass
    sb = sum(stage_bytes[k] for k in streaming)
    sm = sum(stage_ms[k] for k in streaming)
    roof = {"bound": "hbm", "kernel": single[dom], "stage": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic, "peak_source": peak_src, "launches_timed": dom_launches,
            "algorithmic_bytes_per_launch": stage_bytes[dom] / dom_launches, "ms_per_launch": stage_ms[dom] / dom_launches,
            "note": "dominant single streaming kernel by time; the irregular stages (election, SUB, BVE, ERE) are latency bound and listed under stage_ms_per_step",
            "all_streaming_aggregate": {"GB/s": sb / 1e9 / (sm / 1e3), "frac": sb / 1e9 / (sm / 1e3) / peak, "ms_per_step": sm / args.steps,
                                        "algorithmic_bytes_per_step": sb / args.steps},
            "all_streaming": {k: {"kernel": single.get(k, "group of kernels"), "GB/s": stage_bytes[k] / 1e9 / (stage_ms[k] / 1e3),
                                  "frac": stage_bytes[k] / 1e9 / (stage_ms[k] / 1e3) / peak, "ms_per_step": stage_ms[k] / args.steps,
                                  "launches_per_step": stage_launches[k] / args.steps} for k in streaming}}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args), "config": args.config, "literals_per_instance": lits_in,
                   "instances_per_step": world, "l2": "inputs larger than L2 (arena %.0f MB per instance)" % (b.arena.nbytes / 1e6)
                   if b.arena.nbytes > 126e6 else "the instance (arena %.0f MB) fits the 126 MB L2: between steps the pass itself "
                   "rewrites the whole store and its occurrence table" % (b.arena.nbytes / 1e6),
                   "options": "reference defaults" + (", -no-redundancy" if args.no_ere else "")},
        "roofline": roof,
        "e2e": {"value": batch.throughput(total_lits, e2e_serial_ms), "unit": UNIT, "ms_per_step": e2e_serial_ms / args.steps,
                "mode": "one context, blocking: upload -> execute -> download back to back (the call the solver makes)",
                "pipelined": {"value": batch.throughput(total_lits, e2e_pipe_ms), "ms_per_step": e2e_pipe_ms / args.steps,
                              "note": "batch mode: 2 contexts per GPU on alternate steps; every step uploads and downloads"},
                "outputs_agree": bool(same),
                "h2d_bytes_per_step": rep2["h2d_bytes"], "d2h_bytes_per_step": rep2["d2h_bytes"],
                "h2d_ms": rep2["stage_ms"]["h2d"], "d2h_ms": rep2["stage_ms"]["d2h"]},
        "instances_per_s": {"device": world * args.steps / (dev_ms / 1e3), "e2e": world * args.steps / (e2e_serial_ms / 1e3),
                            "e2e_pipelined": world * args.steps / (e2e_pipe_ms / 1e3),
                            "note": "BASELINE metric 'instances/s at 1/2/4/8': one instance of this config per GPU per step; "
                                    "the mixed 64-instance work queue is --config c5"},
        "host_affinity": numa,
        "gpu_launches": launches, "launches_per_step": launches / args.steps, "clocks": clocks,
        "wall_ms_per_step": wall_ms / args.steps,
        "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
        "pass": {"phases": rep["phases_run"], "elected_per_phase": rep["elected_per_phase"], "election_rounds": rep["election_rounds"],
                 "clauses_out": rep["n_clauses"], "literals_out": rep["n_literals"], "eliminated": rep["max_melted_delta"],
                 "subsumed": rep["subsumed"], "strengthened": rep["strengthened"], "host_round_trips": rep["host_round_trips"]},
    }
    # ---- widening #1: extract() on the device (sigma_upload_cnf), informational ------------------
    if rank == 0 and world == 1:
        db = B.clause_db_from_boundary(b, seed=args.seed or 0, dead_frac=0.05)
        ex_us, ex_bytes = [], 0
        for _ in range(4):
            s.upload_cnf(db, b)
            r = s.report()
            ex_us.append(r["extract_us"]); ex_bytes = r["extract_bytes"]
        ex_ms = float(np.mean(ex_us[1:])) / 1e3
        line["extract"] = {"what": "extract(orgs)+extract(learnts) on the device: clause database (5 % deleted records, literals in "
                                   "watch order) -> SCNF with sorted literals and signatures; kernels k_extract_sizes + 2 scans + k_extract_write",
                           "ms": ex_ms, "algorithmic_bytes": ex_bytes, "GB/s": ex_bytes / 1e9 / (ex_ms / 1e3),
                           "frac": ex_bytes / 1e9 / (ex_ms / 1e3) / peak, "h2d_bytes": r["h2d_bytes"], "h2d_ms": r["stage_ms"]["h2d"]}
        del db
    s.close()
    ref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(REF_BIN):
        ref = ref_result(ref_sigma(cnf, ere=not args.no_ere, cpu=host_cpus()[0], full=True))
        line["cpu_baseline"] = {"value": lits_in / (ref["pass_ms"] / 1e3), "unit": UNIT, "cores": 1, "kind": "reference",
                                "sample": f"the full instance ({lits_in} literals), one run of the unmodified reference, region "
                                          f"simplify.cpp:180-206, {cpu_model()}, {os.cpu_count()} host cores (SeqFROST is single-threaded)",
                                "pass_s": ref["pass_ms"] / 1e3, "same_config": True}
    elif rank == 0:
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "measured at N=1 only"}
    if rank == 0 and world == 1 and not args.no_binding:
        line["binding_e2e"] = binding_e2e(cnf, ref)
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
