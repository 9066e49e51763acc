#!/usr/bin/env python
"""bench.py - SIGmA pass throughput on B200 (BASELINE.json metric: pass ms and literals/s per GPU,
HBM GB/s vs peak, instances at 1/2/4/8 GPUs).

A "step" is one whole SIGmA pass (reference simplify.cpp:180-210: every phase of createOT, prop,
LCVE, sortOT, SUB, VE, counts, compaction, export) over one synthetic instance.  At N=1 the
workload is BASELINE.json configs[1]: random 5-SAT, 2M variables / 10M clauses with planted
subsumption work.  At N>1 every rank owns an independent instance of the same
shape (different seed): the formula does not shard, so there is no collective on the data path
(scaling = weak); torch.distributed is used only for the barrier and the max-over-ranks reduce.

  value : literals/s with the formula already resident in HBM (sigma_execute only), timed with
          the library's CUDA events on its own stream, max over ranks
  e2e   : the same through the C ABI with pinned HOST buffers: sigma_upload (H2D) +
          sigma_execute + sigma_download (D2H) inside the timed region
  --impl reference : the unmodified reference (oracle/_ref/ref_sigma, single-threaded like
          SeqFROST itself) timed on the host cores on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_sigma")
METRIC, UNIT = "sigma_pass_literals_per_s", "literals/s"


def workload(nvars, nclauses, seed):
    from seqfrost_b200 import cnfgen
    return cnfgen.planted_subsumption_ksat(nvars, nclauses, k=5, seed=seed)


def workload_name(nvars, nclauses, ere=True):
    return (f"random 5-SAT {nvars / 1e6:g}M vars / {nclauses / 1e6:g}M clauses + planted subsumption "
            f"(BASELINE.json configs[1], SURVEY 8d C2), " + ("reference default options (ERE on)" if ere else "-no-redundancy"))


def reference_pass(nvars, nclauses, seed, runs=1, warm=0, ere=True):
    """Runs the unmodified reference on a DIMACS file of the given shape; returns
    (literals_in, [pass seconds per run]) for the replaced region simplify.cpp:180-206."""
    from seqfrost_b200 import cnfgen
    nv, off, lits = workload(nvars, nclauses, seed)
    secs, lits_in = [], int(off[-1])
    with tempfile.TemporaryDirectory() as td:
        cnf = os.path.join(td, "sample.cnf")
        cnfgen.write_dimacs(cnf, nv, off, lits)
        for i in range(warm + runs):
            r = subprocess.run([REF_BIN, cnf, "-quiet"] + ([] if ere else ["-no-redundancy"]), capture_output=True, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("ref_sigma:")]
            if not line:
                raise RuntimeError("ref_sigma failed: " + r.stdout[-300:] + r.stderr[-300:])
            ms = float(line[-1].split("pass_ms=")[1].split()[0])
            if i >= warm:
                secs.append(ms / 1e3)
    return lits_in, secs


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
    if rank != 0:
        return
    sv, sc = args.vars // args.sample_div, args.clauses // args.sample_div
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_sigma not built (make -C oracle ref needs /root/reference)"}))
        return
    lits_in, secs = reference_pass(sv, sc, args.seed, runs=args.steps, warm=args.warmup, ere=not args.no_ere)
    total = sum(secs)
    value = lits_in * len(secs) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(secs), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args.vars, args.clauses, not args.no_ere), "sample": f"1/{args.sample_div} of the workload per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "reference",
                         "sample": f"same generator at 1/{args.sample_div} size ({sv} vars / {sc} clauses / {lits_in} literals), "
                                   f"replaced region simplify.cpp:180-206, {cpu_model()}, {os.cpu_count()} host cores, SeqFROST is single-threaded"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vars", type=int, default=2_000_000)
    ap.add_argument("--clauses", type=int, default=10_000_000)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--sample-div", type=int, default=10, help="CPU baseline sample = workload / this")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ere", action="store_true", help="run both sides with -no-redundancy")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    from seqfrost_b200 import boundary as B, sigma
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local if world > 1 else 0
    numa = pin_to_gpu_numa_node(torch, dev) if world > 1 else None      # host buffers next to the GPU they feed

    # ---- instance (synthetic, seeded; one independent instance per rank) ------------------------
    from seqfrost_b200 import batch
    nv, off, lits = workload(args.vars, args.clauses, batch.rank_seed(args.seed, rank))
    b = B.from_cnf(nv, off, lits)
    b.opts["ere_en"] = 0 if args.no_ere else 1          # reference default: ERE on (options.cpp:28)
    del off, lits
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
    # (a) serial: one context, upload -> execute -> download, step after step (single-instance latency)
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
    # (b) pipelined: two contexts per GPU (the ABI's unit of concurrency), one host thread each, taking
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
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    if errs:
        raise errs[0]
    same = all(np.array_equal(bufs[k][:n_], bufs2[k][:n_]) for k, n_ in (("arena", 1024), ("refs", 1024))) if args.steps > 1 else True
    s2.close()
    rep2 = s.report()

    total_lits = batch.sum_over_ranks(lits_in, dist, f"cuda:{dev}") * args.steps     # all ranks' instances
    value = batch.throughput(total_lits, dev_ms)
    # ---- roofline of the dominant streaming stage ------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    # stages that are one kernel (launched once, or once per slice) vs groups of small kernels
    single = {"ingest": "k_ingest_fused", "ot_hist": "k_hist", "ot_scatter": "k_scatter", "ot_order": "k_ot_order_staged", "count": "k_count_live"}
    streaming = [k for k in ("ingest", "ot_hist", "ot_scan", "ot_scatter", "ot_order", "ot_update", "vo", "count", "gc", "export") if stage_bytes.get(k)]
    dom = max([k for k in streaming if k in single], key=lambda k: stage_ms[k])
    dom_launches = max(1, stage_launches[dom])
    ach = stage_bytes[dom] / 1e9 / (stage_ms[dom] / 1e3)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            traffic = json.load(f)["traffic_bytes_per_launch"].get(single[dom])
    except (OSError, KeyError, ValueError):
        pass
    roof = {"bound": "hbm", "kernel": single[dom], "stage": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic, "peak_source": peak_src, "launches_timed": dom_launches,
            "algorithmic_bytes_per_launch": stage_bytes[dom] / dom_launches, "ms_per_launch": stage_ms[dom] / dom_launches,
            "note": "dominant single streaming kernel by time; the irregular stages (election, SUB, BVE, ERE) are latency bound and listed under stage_ms_per_step",
            "all_streaming": {k: {"kernel": single.get(k, "group of kernels"), "GB/s": stage_bytes[k] / 1e9 / (stage_ms[k] / 1e3),
                                  "frac": stage_bytes[k] / 1e9 / (stage_ms[k] / 1e3) / peak, "ms_per_step": stage_ms[k] / args.steps,
                                  "launches_per_step": stage_launches[k] / args.steps} for k in streaming}}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args.vars, args.clauses, not args.no_ere), "literals_per_instance": lits_in,
                   "instances_per_step": world, "l2": "inputs larger than L2 (arena %.0f MB per instance)" % (b.arena.nbytes / 1e6),
                   "options": "reference defaults" + (", -no-redundancy" if args.no_ere else "")},
        "roofline": roof,
        "e2e": {"value": batch.throughput(total_lits, e2e_ms), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "mode": "pipelined over 2 contexts per GPU (alternate steps; every step uploads and downloads)",
                "serial": {"value": batch.throughput(total_lits, e2e_serial_ms), "ms_per_step": e2e_serial_ms / args.steps,
                           "note": "one context, upload -> execute -> download back to back"},
                "outputs_agree": bool(same),
                "h2d_bytes_per_step": rep2["h2d_bytes"], "d2h_bytes_per_step": rep2["d2h_bytes"],
                "h2d_ms": rep2["stage_ms"]["h2d"], "d2h_ms": rep2["stage_ms"]["d2h"]},
        "instances_per_s": {"device": world * args.steps / (dev_ms / 1e3), "e2e": world * args.steps / (e2e_ms / 1e3),
                            "note": "BASELINE metric 'instances/s at 1/2/4/8': one 10M-clause instance per GPU per step"},
        "host_affinity": numa,
        "gpu_launches": launches, "clocks": clocks,
        "wall_ms_per_step": wall_ms / args.steps,
        "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
        "pass": {"phases": rep["phases_run"], "elected_per_phase": rep["elected_per_phase"], "election_rounds": rep["election_rounds"],
                 "clauses_out": rep["n_clauses"], "literals_out": rep["n_literals"], "eliminated": rep["max_melted_delta"],
                 "subsumed": rep["subsumed"], "strengthened": rep["strengthened"]},
    }
    # ---- widening #1: extract() on the device (sigma_upload_cnf), informational ------------------
    if rank == 0 and world == 1:
        db = B.clause_db_from_boundary(b, seed=args.seed, dead_frac=0.05)
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
    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(REF_BIN):
        sv, sc = args.vars // args.sample_div, args.clauses // args.sample_div
        li, secs = reference_pass(sv, sc, args.seed, runs=1, ere=not args.no_ere)
        line["cpu_baseline"] = {"value": li / secs[0], "unit": UNIT, "cores": 1, "kind": "reference",
                                "sample": f"same generator at 1/{args.sample_div} size ({sv} vars / {sc} clauses / {li} literals), one run of the "
                                          f"unmodified reference, region simplify.cpp:180-206, {cpu_model()}, {os.cpu_count()} host cores "
                                          f"(SeqFROST is single-threaded)", "pass_s": secs[0]}
    elif rank == 0:
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "measured at N=1 only"}
    if rank == 0:
        print(json.dumps(line))
    s.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
