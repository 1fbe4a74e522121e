#!/usr/bin/env python
"""Benchmark of the phylonium distance pipeline on B200 (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path — process(): ESA build of the reference, anchoring
of every genome, all-pairs comparison — over one batch of simulated genomes.

Workload (BASELINE.json configs[1]): 8 genomes x 5 Mbp from the simf generator
(`simf -s 2 -l 5000000 -d .001 -d .002 -d .005 -d .01 -d .02 -d .03 -d .05`), genome 0 is
the reference.  With N GPUs the work is sharded weak-scaling style: every rank maps 8
genomes of the same family (further mutation seeds), the index is built on rank 0 and
broadcast, the reference-coordinate rows are all-gathered and the 8N x 8N matrix is tiled
over the ranks.

metric/value: query Mbp/s = bases of all mapped genomes / device time of the whole step,
inputs resident in HBM.  e2e: the same through the host-buffer C ABI (phylo_process) from
pinned host memory, H2D and D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DISTS = [0.001, 0.002, 0.005, 0.01, 0.02, 0.03, 0.05]
SIMF_SEED = 2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--length", type=int, default=5_000_000, help="genome length (configs[1]: 5 Mbp)")
    ap.add_argument("--genomes", type=int, default=8, help="genomes per GPU (configs[1]: 8)")
    ap.add_argument("--chunk", type=int, default=0, help="walker chunk length (0 = library default)")
    ap.add_argument("--kmer-k", type=int, default=None, help="K of the descent table (library default: from m)")
    ap.add_argument("--sort-path", type=int, default=0, help="suffix sorter: 0 = pick (packed words), 1 = general")
    ap.add_argument("--table-direct", type=int, default=0, help="K-mer table build: 0/1 entry by entry, 2 level by level (A/B)")
    ap.add_argument("--index", default="replicate", choices=["replicate", "broadcast"],
                    help="multi-GPU: every rank builds the index, or rank 0 builds and broadcasts it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile-pass", action="store_true")
    return ap.parse_args()


def genome_spec(global_index: int):
    """(mutation seed, JC distance) of genome `global_index` in the benchmark family"""
    if global_index == 0:
        return SIMF_SEED, 0.0
    k = global_index % 8
    d = DISTS[k - 1] if k else 0.04
    return SIMF_SEED + global_index, d


def workload_name(args, world):
    return (f"{args.genomes * world} simulated {args.length / 1e6:g} Mbp genomes "
            f"(simf -s {SIMF_SEED}, d={DISTS[0]}..{DISTS[-1]}), reference = genome 0")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            fields = self.FIELDS
            probe = subprocess.run(["nvidia-smi", f"--query-gpu={fields}", "--format=csv,noheader,nounits", "-i",
                                    str(self.gpu_index)], capture_output=True, text=True, timeout=20)
            if probe.returncode != 0 or "not a valid field" in (probe.stdout + probe.stderr).lower():
                fields = fields.replace("clocks_event_reasons", "clocks_throttle_reasons")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={fields}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_traffic(kernel: str, m: int):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/traffic.json), if it was taken at this text length"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        for row in json.load(open(p)):
            if row["kernel"] == kernel and int(row["m"]) == int(m):
                return float(row["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def measured_random_access(kernel: str, m: int):
    """sectors/request and cache hit rates of the random-access kernel from the committed
    ncu capture (profiles/traffic.json)"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        for row in json.load(open(p)):
            if row["kernel"] == kernel and int(row["m"]) == int(m) and "random_access" in row:
                return dict(row["random_access"], kernel=kernel, source=row["source"])
    except Exception:
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------- reference arm

def run_reference(args):
    """The reference's own CPU implementation of process() on the host cores: the
    unmodified sources compiled into oracle/_ref (kind "reference"), else the CPU
    restatement (kind "port").  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from phylonium_b200 import simgen

    lib = oracle_lib.best()
    cores = os.cpu_count() or 1
    genomes = []
    for g in range(args.genomes):  # bounded sample: one GPU's batch
        seed, d = genome_spec(g)
        genomes.append(simgen.simf(SIMF_SEED, seed, args.length, d))
    bases = sum(len(g) for g in genomes)
    times, phases = [], None
    # one step is ~1.4 s on 16 cores (the index build is single threaded): keep the whole run
    # within a few minutes whatever K is, and say how many steps were really timed
    budget_s, spent, steps_done = 150.0, 0.0, 0
    for it in range(args.warmup + args.steps):
        timed = it >= args.warmup
        if spent > budget_s and (not timed or steps_done >= 3):
            if not timed:
                continue
            break
        t0 = time.perf_counter()
        res = lib.process(genomes, 0, 0, threads=cores, timed=True)
        dt = time.perf_counter() - t0
        spent += dt
        if timed:
            times.append(dt)
            steps_done += 1
            phases = res["timings"]
    ms = 1e3 * sum(times) / len(times)
    value = bases / 1e6 / (ms / 1e3)
    sample = f"{args.genomes} x {args.length / 1e6:g} Mbp (one GPU's batch), whole process(), {cores} OpenMP threads"
    line = {
        "impl": "reference", "metric": "query_Mbp_per_s_anchored", "value": value, "unit": "Mbp/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "steps_timed": len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args, 1), "note": "CPU arm always runs one GPU's batch on the host cores; "
                   "at most ~150 s of steps are timed (steps_timed)"},
        "cpu_baseline": {"value": value, "unit": "Mbp/s", "cores": cores, "kind": lib.kind, "sample": sample},
        "phases": {"esa_ms": 1e3 * phases["esa"], "esa_sa_sort_standin_ms": 1e3 * phases["sa_sort"],
                   "anchor_ms": 1e3 * phases["anchor"], "matrix_ms": 1e3 * phases["compare"],
                   "anchor_mbp_s": bases / 1e6 / phases["anchor"],
                   "note": "suffix sort is oracle/sa_standin.cxx, not libdivsufsort (absent from the image)"},
        "e2e": {"value": value, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- B200 arm

def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import phylonium_b200 as pb
    from phylonium_b200 import sharding, simgen

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    G, L = args.genomes, args.length
    total = G * world
    plan = sharding.make_plan(total, world, rank)

    # ---- inputs: pinned host buffer, then resident in HBM --------------------------------
    stride = (L + 1 + 15) // 16 * 16
    host = torch.zeros(G * stride + 64, dtype=torch.uint8).pin_memory()
    hbase = host.data_ptr()
    for k in range(G):
        seed, d = genome_spec(plan.first + k)
        simgen.simf(SIMF_SEED, seed, L, d, out=hbase + k * stride)
    offs = np.arange(G, dtype=np.uint64) * np.uint64(stride)
    lens = np.full(G, L, dtype=np.uint64)
    dQ = host.to(dev, non_blocking=False)
    ref_host_ptr = None
    if world > 1:
        # every rank needs the reference length only; rank 0 holds the reference (its genome 0)
        pass
    bases_local = int(lens.sum())
    bases_total = bases_local * world

    stream = torch.cuda.current_stream()
    ctx = pb.Context(local_rank)
    ctx.set_stream(stream.cuda_stream)
    if args.chunk:
        ctx.set_option("chunk", args.chunk)
    if args.kmer_k is not None:
        ctx.set_option("kmer_k", args.kmer_k)
    ctx.set_option("sort_path", args.sort_path)
    if args.table_direct:
        ctx.set_option("table_direct", args.table_direct)
    d_counts = torch.zeros(2, total * total, dtype=torch.int64, device=dev)
    d_subst, d_homol = d_counts[0], d_counts[1]
    if rank == 0:
        d_ref = dQ[:L]
    else:
        ref_host = torch.zeros(L, dtype=torch.uint8).pin_memory()
        simgen.simf(SIMF_SEED, SIMF_SEED, L, 0.0, out=ref_host.data_ptr())
        d_ref = ref_host.to(dev)
    if world > 1:
        ctx.rows_configure(plan.padded_total, plan.first)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def threshold():
        gc = ctx.stat("esa.gc_count") / L
        return pb.min_anchor_length(0.025, gc, 2 * L + 1)

    thr_box = [None]

    def step():
        """one pass of the hot path, inputs resident on the device"""
        if world == 1 or args.index == "replicate":
            # every rank indexes the (shared) reference itself: no rank waits for another
            ctx.esa_build_dev(d_ref.data_ptr(), L)
        else:
            # north-star layout: built once on rank 0, broadcast over NVLink
            if rank == 0:
                ctx.esa_build_dev(d_ref.data_ptr(), L)
            sharding.broadcast_index(ctx, L, 0, rank, local_rank)
        thr_box[0] = threshold()
        ctx.map_queries_dev(dQ.data_ptr(), offs, lens, thr_box[0])
        if world > 1:
            sharding.allgather_rows(ctx, plan, local_rank)
            ctx.compare_tiles_dev(d_counts[0].data_ptr(), d_counts[1].data_ptr(), rank, world)
            dist.all_reduce(d_counts, op=dist.ReduceOp.SUM)  # every cell is written by one rank
        else:
            ctx.compare_all_dev(d_subst.data_ptr(), d_homol.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    launches0 = ctx.stat("launches")
    step_ms = []
    barrier()
    for _ in range(args.steps):
        flush.zero_()  # evict the inputs from L2 (they are smaller than L2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = int(round((ctx.stat("launches") - launches0) / max(1, args.steps)))
    ms = sum(step_ms) / len(step_ms)
    srt = sorted(step_ms)
    step_stats = {"min": srt[0], "median": srt[len(srt) // 2], "max": srt[-1]}
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = bases_total / 1e6 / (ms / 1e3)
    subst_ref = d_subst.clone()
    homol_ref = d_homol.clone()

    # ---- e2e: host buffers, H2D and D2H inside the timed region --------------------------------
    e2e = None
    e2e_steps = min(max(3, args.steps), 50)
    if world == 1:
        # the reference-facing call: phylo_process() on host pointers (pinned here)
        ptrs = [hbase + k * stride for k in range(G)]
        out = (np.zeros((G, G), np.uint64), np.zeros((G, G), np.uint64))
        ctx.process_ptrs(ptrs, lens, 0, 0, out)  # warm-up
        ts = []
        for _ in range(e2e_steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.process_ptrs(ptrs, lens, 0, 0, out)
            ts.append(time.perf_counter() - t0)
        e_ms = 1e3 * sum(ts) / len(ts)
        same = bool((torch.from_numpy(out[0].astype(np.int64)).reshape(-1) == subst_ref.cpu()).all()
                    and (torch.from_numpy(out[1].astype(np.int64)).reshape(-1) == homol_ref.cpu()).all())
        e2e = {"value": bases_total / 1e6 / (e_ms / 1e3), "unit": "Mbp/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(bases_local), "d2h_bytes_per_step": int(2 * G * G * 8),
               "same_counts_as_device_path": same, "call": "phylo_process (C ABI, host pointers)"}
    else:
        # sharded: every rank uploads its own genomes (and the shared reference) from pinned
        # memory, runs the sharded step, and rank 0 reads the count matrices back
        h_counts = torch.zeros(2, total * total, dtype=torch.int64).pin_memory()
        ts = []
        for it in range(e2e_steps + 1):
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            dQ.copy_(host, non_blocking=True)
            if rank != 0:
                d_ref.copy_(ref_host, non_blocking=True)
            step()
            if rank == 0:
                h_counts.copy_(d_counts, non_blocking=True)
            barrier()
            if it:
                ts.append(time.perf_counter() - t0)
        e_ms = 1e3 * sum(ts) / len(ts)
        t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item())
        same = bool((h_counts.reshape(-1) == torch.stack([subst_ref, homol_ref]).reshape(-1).cpu()).all())
        e2e = {"value": bases_total / 1e6 / (e_ms / 1e3), "unit": "Mbp/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(bases_local + (0 if rank == 0 else L)) * world,
               "d2h_bytes_per_step": int(2 * total * total * 8), "same_counts_as_device_path": same,
               "call": "sharded step of phylonium_b200.sharding (pinned host buffers per rank)"}

    # ---- profile pass: per-phase device times and the roofline of the dominant kernel -------
    phases, roofline = None, None
    if rank == 0 and world == 1 and not args.no_profile_pass:
        ctx.set_option("timings", 1)
        acc = {}
        reps = 3
        for _ in range(reps):
            flush.zero_()
            torch.cuda.synchronize()
            step()
            for k in ("esa.text_ms", "esa.keys_ms", "esa.sort_ms", "esa.refine_ms", "esa.lcp_ms", "esa.cld_ms",
                      "esa.table_ms", "esa.total_ms", "anchor.walk_ms", "anchor.open_ms", "anchor.bridge_ms",
                      "anchor.path_ms", "anchor.assemble_ms", "anchor.filter_ms", "anchor.total_ms", "rows.ms",
                      "compare.ms", "esa.scatter_ms_avg", "esa.scatter_launches", "esa.first_pass_ms", "esa.hist_ms_avg",
                      "esa.scan_ms_avg"):
                acc[k] = acc.get(k, 0.0) + ctx.stat(k) / reps
        ctx.set_option("timings", 0)
        phases = {k: round(v, 4) for k, v in acc.items()}
        phases["anchor_mbp_s"] = bases_local / 1e6 / (acc["anchor.total_ms"] / 1e3)
        phases["matrix_ms"] = acc["compare.ms"]
        phases["threshold"] = thr_box[0]
        peak, peak_src = measured_peak()
        m = 2 * L + 1
        packed = ctx.stat("esa.packed") == 1
        if acc.get("esa.scatter_ms_avg", -1) > 0:
            # dominant kernel: one radix pass of the suffix sort.  pk_scatter moves one packed
            # 64-bit word per suffix (8 B read + 8 B written); the general sorter's rs_scatter a
            # u64 key and a u32 index (12 B + 12 B)
            bytes_per_launch = (16.0 if packed else 24.0) * m
            achieved = bytes_per_launch / (acc["esa.scatter_ms_avg"] * 1e-3) / 1e9
            kname = "pk_scatter<false>" if packed else "rs_scatter"
            roofline = {"kernel": kname + " (radix pass of the suffix sort)",
                        "bound": "hbm",
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": measured_traffic(kname, m), "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": bytes_per_launch,
                        "ms_per_launch": acc["esa.scatter_ms_avg"],
                        "random_access_kernel": measured_random_access("k_walk_chunks", m)}

    # ---- CPU baseline next to it (rank 0, N = 1) ------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib

            lib = oracle_lib.best()
            cores = os.cpu_count() or 1
            genomes = [bytes((ctypes.c_char * L).from_address(hbase + k * stride)) for k in range(G)]
            t0 = time.perf_counter()
            res = lib.process(genomes, 0, 0, threads=cores, timed=True)
            dt = time.perf_counter() - t0
            ok = bool((res["subst"].astype(np.int64).reshape(-1) == subst_ref.cpu().numpy()).all()
                      and (res["homologs"].astype(np.int64).reshape(-1) == homol_ref.cpu().numpy()).all())
            cpu = {"value": bases_local / 1e6 / dt, "unit": "Mbp/s", "cores": cores, "kind": lib.kind,
                   "sample": f"the full workload once ({G} x {L / 1e6:g} Mbp), whole process(), {cores} OpenMP threads",
                   "seconds": dt, "esa_s": res["timings"]["esa"], "sa_sort_standin_s": res["timings"]["sa_sort"],
                   "anchor_s": res["timings"]["anchor"], "matrix_s": res["timings"]["compare"],
                   "counts_equal_gpu": ok}
        except Exception as e:  # the checker is optional for the measurement itself
            cpu = {"unavailable": str(e)}

    if rank == 0:
        line = {
            "metric": "query_Mbp_per_s_anchored", "value": value, "unit": "Mbp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "genomes_per_gpu": G, "genome_length": L,
                       "l2": "flushed between timed steps (512 MiB memset)",
                       "step": "ESA build + anchoring of all genomes + all-pairs counts (process())",
                       "parallelism": (f"queries sharded x{world}, index {args.index}d, rows all-gathered, "
                                       f"matrix tiles dealt to ranks + all-reduce") if world > 1 else "single GPU"},
            "gpu_launches": launches, "step_ms": step_stats, "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
            "phases": phases,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
