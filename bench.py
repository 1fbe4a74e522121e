#!/usr/bin/env python
"""Benchmark of the phylonium distance pipeline on B200 (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path — process(): ESA build of the reference, anchoring
of every genome, all-pairs comparison — over one batch of simulated genomes.

Headline workload (BASELINE.json configs[1]): 8 genomes x 5 Mbp from the simf generator
(`simf -s 2 -l 5000000 -d .001 -d .002 -d .005 -d .01 -d .02 -d .03 -d .05`), genome 0 is
the reference.  With N GPUs the headline is weak-scaled: every rank maps 8 genomes of the same
family (further mutation seeds), every rank indexes the shared reference itself, each batch
of reference-coordinate rows is pushed into the peers' row stores over NVLink while the next
batch is mapped, and the 8N x 8N matrix is tiled over the ranks.  The `--impl reference` arm
runs the SAME 8N genomes through the unmodified reference on the host cores.

metric/value: query Mbp/s = bases of all mapped genomes / device time of the whole step,
inputs resident in HBM.  e2e: the same through the host-buffer C ABI (phylo_process) from
pinned host memory, H2D and D2H inside the timed region; at N > 1 through
phylo_esa_build / phylo_map_queries on host buffers per rank.

`north_star`: BASELINE.json's larger configurations — configs[2] 100 x 5 Mbp, configs[3]
1000 x 3 Mbp and the north-star target 1000 x 5 Mbp — STRONG-scaled over the same N GPUs
(genomes dealt round-robin to the ranks), each with per-phase times (index, loop A, matrix),
an end-to-end time, and 32 sampled rows of the matrix checked against the reference's CPU code.
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
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DISTS = [0.001, 0.002, 0.005, 0.01, 0.02, 0.03, 0.05]
SIMF_SEED = 2

# strong-scaled workloads of the `north_star` section: (name, simf seed, genomes, length, largest distance)
NORTH_STAR = [
    ("configs[2]: 100 x 5 Mbp", 3, 100, 5_000_000, 0.05),
    ("configs[3]: 1000 x 3 Mbp", 4, 1000, 3_000_000, 0.05),
    ("north-star target: 1000 x 5 Mbp", 6, 1000, 5_000_000, 0.05),
]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--length", type=int, default=5_000_000, help="genome length (configs[1]: 5 Mbp)")
    ap.add_argument("--genomes", type=int, default=8, help="genomes per GPU (configs[1]: 8)")
    ap.add_argument("--chunk", type=int, default=0, help="walker chunk length (0 = library default)")
    ap.add_argument("--kmer-k", type=int, default=None, help="K of the descent table (library default: from m)")
    ap.add_argument("--sort-path", type=int, default=0, help="suffix sorter: 0 = pick (packed words), 1 = general")
    ap.add_argument("--index", default="replicate", choices=["replicate", "broadcast"],
                    help="multi-GPU: every rank builds the index, or rank 0 builds and broadcasts it")
    ap.add_argument("--exchange", default="push", choices=["push", "nccl"],
                    help="multi-GPU rows: pushed into the peers' stores batch by batch, or one NCCL all-gather")
    ap.add_argument("--no-alternatives", action="store_true", help="multi-GPU: skip timing the other index/exchange modes")
    ap.add_argument("--north-star", default="all", help="all | none | comma-separated indices into the list (0,1,2)")
    ap.add_argument("--ns-steps", type=int, default=5, help="timed steps per north-star workload")
    ap.add_argument("--ns-scale", type=float, default=1.0, help="scale the north-star genome lengths (smoke runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the CPU checks of the results")
    ap.add_argument("--no-profile-pass", action="store_true")
    return ap.parse_args()


def genome_spec(global_index: int):
    """(mutation seed, JC distance) of genome `global_index` in the headline family"""
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
        # the median over the samples taken under load (idle samples between the sections of
        # this script sit at the idle clock and say nothing about the timed steps)
        busy = sorted(x for x in sm if x >= 0.8 * mx) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "samples_under_load": len(busy)}


def _traffic_rows():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return []


def measured_traffic(kernel: str, m: int):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/traffic.json), if it was taken at this text length"""
    for row in _traffic_rows():
        if row.get("kernel") == kernel and int(row.get("m", -1)) == int(m) and "dram_bytes_per_launch" in row:
            return float(row["dram_bytes_per_launch"])
    return None


def measured_random_access(kernel: str, m: int):
    """sectors/request and cache hit rates of the random-access kernel from the committed
    ncu capture (profiles/traffic.json)"""
    for row in _traffic_rows():
        if row.get("kernel") == kernel and int(row.get("m", -1)) == int(m) and "random_access" in row:
            return dict(row["random_access"], kernel=kernel, source=row["source"])
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads(world: int) -> int:
    return max(1, (os.cpu_count() or 1) // max(1, world))


# ------------------------------------------------------------------------------- reference arm

def run_reference(args):
    """The reference's own CPU implementation of process() on the host cores: the
    unmodified sources compiled into oracle/_ref (kind "reference"), else the CPU
    restatement (kind "port"), on the SAME workload as the B200 arm at this N: 8 N genomes.
    The timed call is the reference's untouched process() (no clocks between its phases); the
    per-phase split comes from one extra, instrumented call.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from phylonium_b200 import simgen

    lib = oracle_lib.best()
    cores = os.cpu_count() or 1
    world = max(1, args.gpus)
    total = args.genomes * world
    specs = [genome_spec(g) for g in range(total)]
    genomes = [None] * total

    def gen(i):
        genomes[i] = simgen.simf(SIMF_SEED, specs[i][0], args.length, specs[i][1])

    with ThreadPoolExecutor(max_workers=cores) as ex:
        list(ex.map(gen, range(total)))
    bases = sum(len(g) for g in genomes)
    times = []
    # one step is ~1.4 s (8 genomes) to ~3 s (64 genomes) on 16 cores: keep the whole run within
    # a few minutes whatever K is, and say how many steps were really timed
    budget_s, spent, steps_done = 150.0, 0.0, 0
    for it in range(args.warmup + args.steps):
        timed = it >= args.warmup
        if spent > budget_s and (not timed or steps_done >= 3):
            if not timed:
                continue
            break
        t0 = time.perf_counter()
        lib.process(genomes, 0, 0, threads=cores, timed=False)  # the untouched process() call
        dt = time.perf_counter() - t0
        spent += dt
        if timed:
            times.append(dt)
            steps_done += 1
    phases = lib.process(genomes, 0, 0, threads=cores, timed=True)["timings"]
    ms = 1e3 * sum(times) / len(times)
    value = bases / 1e6 / (ms / 1e3)
    sample = (f"{total} x {args.length / 1e6:g} Mbp (the B200 arm's whole workload at {world} GPU(s)), "
              f"process() of the unmodified reference, {cores} OpenMP threads")
    line = {
        "impl": "reference", "metric": "query_Mbp_per_s_anchored", "value": value, "unit": "Mbp/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "steps_timed": len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "genomes_per_gpu": args.genomes, "genome_length": args.length,
                   "note": "same genomes as the B200 arm at this N; at most ~150 s of steps are timed (steps_timed)"},
        "cpu_baseline": {"value": value, "unit": "Mbp/s", "cores": cores, "kind": lib.kind, "sample": sample},
        "phases": {"esa_ms": 1e3 * phases["esa"], "esa_sa_sort_standin_ms": 1e3 * phases["sa_sort"],
                   "anchor_ms": 1e3 * phases["anchor"], "matrix_ms": 1e3 * phases["compare"],
                   "anchor_mbp_s": bases / 1e6 / phases["anchor"],
                   "note": "from one extra instrumented call; suffix sort is oracle/sa_standin.cxx, not libdivsufsort "
                           "(absent from the image)"},
        "e2e": {"value": value, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- B200 arm

class Shard:
    """this rank's genomes: pinned host buffer (every genome 16-byte aligned and followed by a
    zero byte) and its copy in HBM"""

    def __init__(self, torch, simgen, dev, base_seed, specs, length, threads):
        import numpy as np

        self.count, self.length = len(specs), length
        self.stride = (length + 1 + 15) // 16 * 16
        self.host = torch.zeros(max(1, self.count) * self.stride + 64, dtype=torch.uint8).pin_memory()
        hbase = self.host.data_ptr()

        def gen(k):
            simgen.simf(base_seed, specs[k][0], length, specs[k][1], out=hbase + k * self.stride)

        if self.count:
            with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
                list(ex.map(gen, range(self.count)))
        self.ptrs = [hbase + k * self.stride for k in range(self.count)]
        self.offs = np.arange(self.count, dtype=np.uint64) * np.uint64(self.stride)
        self.lens = np.full(self.count, length, dtype=np.uint64)
        self.dQ = self.host.to(dev, non_blocking=False)
        self.bases = self.count * length


class Pipeline:
    """one rank's share of a (possibly sharded) process() pass"""

    def __init__(self, mods, args, dev, local_rank, rank, world, plan, shard, ref_host, length, index_mode, exchange,
                 batch_bytes=None):
        self.torch, self.dist, self.pb, self.sharding = mods
        torch, pb, sharding = self.torch, self.pb, self.sharding
        self.dev, self.local_rank, self.rank, self.world = dev, local_rank, rank, world
        self.plan, self.shard, self.L = plan, shard, length
        self.index_mode, self.exchange = index_mode, exchange
        self.stream = torch.cuda.current_stream()
        self.ctx = pb.Context(local_rank)
        self.ctx.set_stream(self.stream.cuda_stream)
        if args.chunk:
            self.ctx.set_option("chunk", args.chunk)
        if args.kmer_k is not None:
            self.ctx.set_option("kmer_k", args.kmer_k)
        self.ctx.set_option("sort_path", args.sort_path)
        # the ranks share the host's cores; at one rank the library's own choice (two cores left for
        # the calling thread and the driver: 16 packers on 16 cores measured 2.2 ms against 1.9 ms with 14)
        self.ctx.set_option("stage_threads", 0 if world == 1 else max(1, min(16, host_threads(world))))
        if world > 1 and batch_bytes:
            # several batches per rank: the rows of one batch cross NVLink while the next is mapped
            self.ctx.set_option("map_batch_bytes", batch_bytes)
        self.n = plan.padded_total if world > 1 else plan.total
        self.d_counts = torch.zeros(2, self.n * self.n, dtype=torch.int64, device=dev)
        self.ref_host = ref_host  # pinned copy of the reference on this rank
        self.d_ref = shard.dQ[:length] if (rank == 0 and shard.count) else ref_host.to(dev)
        self.thr = None
        if world > 1:
            self.ctx.esa_build_dev(self.d_ref.data_ptr(), length)  # the row store needs the reference length
            self.ctx.rows_configure(plan.padded_total, plan.first)
            if exchange == "push":
                sharding.setup_push(self.ctx, rank, world)

    def set_exchange(self, exchange):
        if exchange == self.exchange:
            return
        if exchange == "push":
            self.sharding.setup_push(self.ctx, self.rank, self.world)
        else:
            self.ctx.rows_set_peers(None, 0)
        self.exchange = exchange

    def threshold(self):
        gc = self.ctx.stat("esa.gc_count") / self.L
        return self.pb.min_anchor_length(0.025, gc, 2 * self.L + 1)

    def _index(self, host=False):
        ctx, L = self.ctx, self.L
        if self.world == 1 or self.index_mode == "replicate":
            # every rank indexes the (shared) reference itself: no rank waits for another
            if host:
                ctx.esa_build_ptr(self.ref_host.data_ptr(), L)
            else:
                ctx.esa_build_dev(self.d_ref.data_ptr(), L)
        else:
            # north-star layout: built once on rank 0, broadcast over NVLink
            if self.rank == 0:
                if host:
                    ctx.esa_build_ptr(self.ref_host.data_ptr(), L)
                else:
                    ctx.esa_build_dev(self.d_ref.data_ptr(), L)
            self.sharding.broadcast_index(ctx, L, 0, self.rank, self.local_rank)
        self.thr = self.threshold()

    def _matrix(self):
        ctx, c = self.ctx, self.d_counts
        if self.world > 1:
            if self.exchange == "push":
                self.sharding.rows_barrier(self.dev)  # every rank's pushes have landed
            else:
                self.sharding.allgather_rows(ctx, self.plan, self.local_rank)
            ctx.compare_tiles_dev(c[0].data_ptr(), c[1].data_ptr(), self.rank, self.world)
            self.dist.all_reduce(c, op=self.dist.ReduceOp.SUM)  # also the barrier before the next step's pushes
        else:
            ctx.compare_all_dev(c[0].data_ptr(), c[1].data_ptr())

    def step(self):
        """one pass of the hot path, inputs resident on the device"""
        self._index()
        sh = self.shard
        self.ctx.map_queries_dev(sh.dQ.data_ptr(), sh.offs, sh.lens, self.thr)
        self._matrix()

    def step_from_host(self):
        """the same from (pinned) host buffers through phylo_esa_build / phylo_map_queries: the
        sequences cross PCIe inside the call, batch by batch, next to the mapping"""
        self._index(host=True)
        sh = self.shard
        self.ctx.map_queries_ptrs(sh.ptrs, sh.lens, self.thr)
        self._matrix()

    def counts(self):
        """(2, total, total) int64 in genome order"""
        if self.world > 1 and (self.plan.layout != "block" or self.plan.padded_total != self.plan.total):
            return self.sharding.genome_order(self.d_counts, self.plan)
        return self.d_counts.reshape(2, self.n, self.n)  # slot order is genome order

    def close(self):
        self.ctx.close()


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
    mods = (torch, dist, pb, sharding)
    # one explicit stream for everything: the library's kernels (phylo_set_stream), torch's own
    # work, the NCCL collectives and the timing events.  (The default stream has the handle 0,
    # which phylo_set_stream reads as "use the context's own stream" — events recorded on
    # torch's stream would then not see the library's kernels.)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    threads = host_threads(world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def time_steps(fn, steps, warmup, wall=False):
        """per-step times in ms; device events on the launching stream, or (wall) host clock
        around the call with a synchronisation on both sides"""
        for _ in range(warmup):
            fn()
        out = []
        barrier()
        for _ in range(steps):
            flush.zero_()  # evict the inputs from L2 (they are smaller than L2)
            barrier()
            if wall:
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                out.append(1e3 * (time.perf_counter() - t0))
            else:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                e1.synchronize()
                out.append(e0.elapsed_time(e1))
        barrier()
        return out

    # =============================================================== headline: configs[1], weak
    G, L = args.genomes, args.length
    total = G * world
    plan = sharding.make_plan(total, world, rank)
    specs = [genome_spec(g) for g in plan.genomes()]
    shard = Shard(torch, simgen, dev, SIMF_SEED, specs, L, threads)
    if rank == 0:
        ref_host = shard.host[:L]
    else:
        ref_host = torch.zeros(L, dtype=torch.uint8).pin_memory()
        simgen.simf(SIMF_SEED, SIMF_SEED, L, 0.0, out=ref_host.data_ptr())
    # (8 genomes per rank are one batch: their rows, 25 MB, are pushed in one go behind the mapping)
    pipe = Pipeline(mods, args, dev, local_rank, rank, world, plan, shard, ref_host, L, args.index, args.exchange)
    ctx = pipe.ctx
    bases_local, bases_total = shard.bases, G * L * world

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        pipe.step()
    launches0 = ctx.stat("launches")
    step_ms = time_steps(pipe.step, args.steps, 0)
    launches = int(round((ctx.stat("launches") - launches0) / max(1, args.steps)))
    ms = max_over_ranks(sum(step_ms) / len(step_ms))
    srt = sorted(step_ms)
    step_stats = {"min": srt[0], "median": srt[len(srt) // 2], "max": srt[-1]}
    value = bases_total / 1e6 / (ms / 1e3)
    counts_dev = pipe.counts().clone()

    # ---- the other index / exchange modes, a few steps each (N > 1) -----------------------------
    alternatives = None
    if world > 1 and not args.no_alternatives:
        alternatives = {}
        alt_steps = min(args.steps, 10)
        for index_mode, exchange in (("replicate", "push"), ("broadcast", "push"), ("replicate", "nccl"), ("broadcast", "nccl")):
            if (index_mode, exchange) == (args.index, args.exchange):
                alternatives[f"index {index_mode}, rows {exchange}"] = {"ms_per_step": ms, "steps": args.steps, "primary": True}
                continue
            pipe.index_mode = index_mode
            pipe.set_exchange(exchange)
            t = time_steps(pipe.step, alt_steps, 2)
            same = bool((pipe.counts() == counts_dev).all())
            alternatives[f"index {index_mode}, rows {exchange}"] = {
                "ms_per_step": max_over_ranks(sum(t) / len(t)), "steps": alt_steps, "same_counts": same}
        pipe.index_mode = args.index
        pipe.set_exchange(args.exchange)

    # ---- e2e: host buffers, H2D and D2H inside the timed region --------------------------------
    e2e_steps = min(max(3, args.steps), 50)
    if world == 1:
        # the reference-facing call: phylo_process() on host pointers (pinned here)
        out = (np.zeros((G, G), np.uint64), np.zeros((G, G), np.uint64))
        ts = time_steps(lambda: ctx.process_ptrs(shard.ptrs, shard.lens, 0, 0, out), e2e_steps, 1, wall=True)
        e_ms = sum(ts) / len(ts)
        same = bool((torch.from_numpy(out[0].astype(np.int64)) == counts_dev[0].cpu()).all()
                    and (torch.from_numpy(out[1].astype(np.int64)) == counts_dev[1].cpu()).all())
        e2e = {"value": bases_total / 1e6 / (e_ms / 1e3), "unit": "Mbp/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(bases_local), "d2h_bytes_per_step": int(2 * G * G * 8),
               "same_counts_as_device_path": same, "call": "phylo_process (C ABI, pinned host pointers)",
               "bus_bytes_per_step": int(ctx.stat("process.h2d_bytes")),
               "host_timeline_ms": {"index_done": ctx.stat("process.host_index_done_ms"), "map_done": ctx.stat("process.host_map_done_ms"),
                                    "done": ctx.stat("process.host_done_ms")},
               "note": "h2d_bytes_per_step counts the caller's buffers; the library packs the query sequences to 2 bits per "
                       "base on the host, so bus_bytes_per_step cross PCIe"}
        # the same call on ordinary (pageable) memory, as the C++ host's std::string storage is
        pageable = [np.frombuffer((ctypes.c_char * L).from_address(p), dtype=np.uint8).copy() for p in shard.ptrs]
        pts = time_steps(lambda: ctx.process_ptrs([a.ctypes.data for a in pageable], shard.lens, 0, 0, out), e2e_steps, 1,
                         wall=True)
        p_ms = sum(pts) / len(pts)
        e2e["pageable"] = {"value": bases_total / 1e6 / (p_ms / 1e3), "ms_per_step": p_ms,
                           "seen_as_pageable": ctx.stat("process.pageable") == 1,
                           "same_counts": bool((torch.from_numpy(out[1].astype(np.int64)) == counts_dev[1].cpu()).all())}
        del pageable
    else:
        # sharded: every rank hands its genomes (and the shared reference) over as pinned host
        # buffers; rank 0 reads the count matrices back
        h_counts = torch.zeros(2, total, total, dtype=torch.int64).pin_memory()

        def e2e_step():
            pipe.step_from_host()
            if rank == 0:
                h_counts.copy_(pipe.counts(), non_blocking=True)

        ts = time_steps(e2e_step, e2e_steps, 1, wall=True)
        e_ms = max_over_ranks(sum(ts) / len(ts))
        same = bool((h_counts == counts_dev.cpu()).all()) if rank == 0 else None
        e2e = {"value": bases_total / 1e6 / (e_ms / 1e3), "unit": "Mbp/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(bases_total + (L * (world - 1) if args.index == "replicate" else 0)),
               "d2h_bytes_per_step": int(2 * total * total * 8), "same_counts_as_device_path": same,
               "call": "phylo_esa_build + phylo_map_queries on pinned host buffers per rank, rows pushed to the peers, "
                       "phylo_compare_tiles_dev + all-reduce, matrix read back on rank 0"}

    clocks = sampler.stop() if rank == 0 else None  # sampled over the timed steps and the e2e steps

    # ---- profile pass: per-phase device times and the roofline of the dominant streaming kernel ---
    phases, roofline = None, None
    if not args.no_profile_pass:
        if rank == 0:
            ctx.set_option("timings", 1)
        acc, reps = {}, 3
        keys = ("esa.text_ms", "esa.keys_ms", "esa.sort_ms", "esa.refine_ms", "esa.lcp_ms", "esa.cld_ms", "esa.table_ms",
                "esa.total_ms", "anchor.walk_ms", "anchor.open_ms", "anchor.bridge_ms", "anchor.path_ms",
                "anchor.assemble_ms", "anchor.filter_ms", "anchor.total_ms", "rows.ms", "compare.ms",
                "esa.scatter_ms_avg", "esa.scatter_launches", "esa.first_pass_ms", "esa.hist_ms_avg", "esa.scan_ms_avg")
        for _ in range(reps):
            flush.zero_()
            barrier()
            pipe.step()
            torch.cuda.synchronize()
            for k in keys:
                acc[k] = acc.get(k, 0.0) + ctx.stat(k) / reps
        ctx.set_option("timings", 0)
        if rank == 0:
            phases = {k: round(v, 4) for k, v in acc.items()}
            phases["anchor_mbp_s_per_gpu"] = bases_local / 1e6 / (acc["anchor.total_ms"] / 1e3)
            phases["index_mbp_s"] = L / 1e6 / (acc["esa.total_ms"] / 1e3)
            phases["matrix_ms"] = acc["compare.ms"]
            phases["threshold"] = pipe.thr
            peak, peak_src = measured_peak()
            m = 2 * L + 1
            packed = ctx.stat("esa.packed") == 1
            if acc.get("esa.scatter_ms_avg", -1) > 0:
                # dominant streaming kernel: one radix pass of the suffix sort.  pk_scatter moves one
                # packed 64-bit word per suffix (8 B read + 8 B written); the general sorter's
                # rs_scatter a u64 key and a u32 index (12 B + 12 B) (kernels)
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

    # ---- the CPU next to it: checks the counts (every N), cpu_baseline (N = 1) -------------------
    cpu, check = None, None
    if rank == 0 and not args.no_check:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_lib

            lib = oracle_lib.best()
            cores = os.cpu_count() or 1
            all_specs = [genome_spec(g) for g in range(total)]
            genomes = [None] * total

            def gen(i):
                genomes[i] = simgen.simf(SIMF_SEED, all_specs[i][0], L, all_specs[i][1])

            with ThreadPoolExecutor(max_workers=cores) as ex:
                list(ex.map(gen, range(total)))
            t0 = time.perf_counter()
            res = lib.process(genomes, 0, 0, threads=cores, timed=True)
            dt = time.perf_counter() - t0
            got = counts_dev.cpu().numpy()
            ok = bool((res["subst"].astype(np.int64) == got[0]).all() and (res["homologs"].astype(np.int64) == got[1]).all())
            check = {"counts_equal_reference": ok, "checker": lib.kind, "genomes": total, "pairs": total * (total - 1) // 2,
                     "what": "every cell of both count matrices of the timed device path against process() of the checker"}
            if world == 1 and not args.no_cpu_baseline:
                cpu = {"value": bases_total / 1e6 / dt, "unit": "Mbp/s", "cores": cores, "kind": lib.kind,
                       "sample": f"the full workload once ({G} x {L / 1e6:g} Mbp), whole process(), {cores} OpenMP threads",
                       "seconds": dt, "esa_s": res["timings"]["esa"], "sa_sort_standin_s": res["timings"]["sa_sort"],
                       "anchor_s": res["timings"]["anchor"], "matrix_s": res["timings"]["compare"],
                       "counts_equal_gpu": ok}
            del genomes
        except Exception as e:  # the checker is optional for the measurement itself
            check = {"unavailable": str(e)}
    pipe.close()
    del pipe, shard
    torch.cuda.empty_cache()
    barrier()

    # =============================================================== north star: strong scaling
    north = None
    if args.north_star != "none":
        picks = range(len(NORTH_STAR)) if args.north_star == "all" else [int(x) for x in args.north_star.split(",")]
        north = []
        for idx in picks:
            name, seed, ntot, length, dmax = NORTH_STAR[idx]
            length = max(1000, int(length * args.ns_scale))
            north.append(run_north_star(mods, args, dev, local_rank, rank, world, name, seed, ntot, length, dmax, threads,
                                        time_steps, max_over_ranks, barrier, flush))
            torch.cuda.empty_cache()
            barrier()

    if rank == 0:
        line = {
            "metric": "query_Mbp_per_s_anchored", "value": value, "unit": "Mbp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "genomes_per_gpu": G, "genome_length": L,
                       "l2": "flushed between timed steps (512 MiB memset)",
                       "step": "ESA build + anchoring of all genomes + all-pairs counts (process())",
                       "parallelism": (f"queries sharded x{world}, index {args.index}d, rows "
                                       f"{'pushed into the peers row stores (NVLink, per batch)' if args.exchange == 'push' else 'all-gathered (NCCL)'}, "
                                       f"matrix tiles dealt to ranks + all-reduce") if world > 1 else "single GPU"},
            "gpu_launches": launches, "step_ms": step_stats, "clocks": clocks, "e2e": e2e, "roofline": roofline,
            "cpu_baseline": cpu, "check": check, "alternatives": alternatives, "phases": phases, "north_star": north,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_north_star(mods, args, dev, local_rank, rank, world, name, seed, ntot, length, dmax, threads, time_steps,
                   max_over_ranks, barrier, flush):
    """one of BASELINE.json's larger configurations, strong-scaled over the ranks"""
    import numpy as np

    torch, dist, pb, sharding = mods
    from phylonium_b200 import simgen

    def spec(i):
        return (seed + i, dmax * i / (ntot - 1) if i else 0.0)

    plan = sharding.make_plan(ntot, world, rank, "interleaved")
    t0 = time.perf_counter()
    shard = Shard(torch, simgen, dev, seed, [spec(g) for g in plan.genomes()], length, threads)
    if rank == 0:
        ref_host = shard.host[:length]
    else:
        ref_host = torch.zeros(length, dtype=torch.uint8).pin_memory()
        simgen.simf(seed, seed, length, 0.0, out=ref_host.data_ptr())
    gen_s = time.perf_counter() - t0
    pipe = Pipeline(mods, args, dev, local_rank, rank, world, plan, shard, ref_host, length, args.index, args.exchange,
                    batch_bytes=max(16 * length, (shard.bases + 2) // 3))  # three batches per rank
    ctx = pipe.ctx
    bases_total = ntot * length
    steps = max(1, args.ns_steps)

    t = time_steps(pipe.step, steps, 2)
    ms = max_over_ranks(sum(t) / len(t))
    counts_dev = pipe.counts().clone()

    # end to end: host buffers in, matrix out
    if world == 1:
        out = (np.zeros((ntot, ntot), np.uint64), np.zeros((ntot, ntot), np.uint64))
        te = time_steps(lambda: ctx.process_ptrs(shard.ptrs, shard.lens, 0, 0, out), min(steps, 3), 1, wall=True)
        e_ms = sum(te) / len(te)
        same = bool((torch.from_numpy(out[1].astype(np.int64)) == counts_dev[1].cpu()).all()
                    and (torch.from_numpy(out[0].astype(np.int64)) == counts_dev[0].cpu()).all())
        call = "phylo_process (C ABI, pinned host pointers)"
    else:
        h_counts = torch.zeros(2, ntot, ntot, dtype=torch.int64).pin_memory()

        def e2e_step():
            pipe.step_from_host()
            if rank == 0:
                h_counts.copy_(pipe.counts(), non_blocking=True)

        te = time_steps(e2e_step, min(steps, 3), 1, wall=True)
        e_ms = max_over_ranks(sum(te) / len(te))
        same = bool((h_counts == counts_dev.cpu()).all()) if rank == 0 else None
        call = "phylo_esa_build + phylo_map_queries on pinned host buffers per rank, rows pushed, tiles + all-reduce"

    # per-phase device times of this rank (rank 0 reports its own)
    if rank == 0:
        ctx.set_option("timings", 1)
    flush.zero_()
    barrier()
    pipe.step()
    torch.cuda.synchronize()
    ph = {k: ctx.stat(k) for k in ("esa.total_ms", "anchor.total_ms", "anchor.walk_ms", "rows.ms", "compare.ms", "map.batches")}
    ctx.set_option("timings", 0)
    barrier()

    entry = None
    if rank == 0:
        got = counts_dev.cpu().numpy()
        pairs = ntot * (ntot - 1) // 2
        alg_bytes = 2.0 * float(np.triu(got[1], 1).sum())  # 2 B per (pair, homologous column): what seqcmp reads
        peak, peak_src = measured_peak()
        W = (length + 31) // 32
        matrix_ms = ph["compare.ms"]
        entry = {
            "workload": name, "genomes": ntot, "genome_length": length, "n_gpus": world, "scaling": "strong",
            "layout": "genomes dealt round-robin to the ranks" if world > 1 else "single GPU",
            "steps": steps, "ms_per_step": ms, "value": bases_total / 1e6 / (ms / 1e3), "unit": "Mbp/s",
            "e2e": {"ms_per_step": e_ms, "value": bases_total / 1e6 / (e_ms / 1e3), "unit": "Mbp/s",
                    "h2d_bytes_per_step": int(bases_total + (length * (world - 1))), "d2h_bytes_per_step": int(2 * ntot * ntot * 8),
                    "same_counts_as_device_path": same, "call": call},
            "phases_rank0": {"index_ms": ph["esa.total_ms"], "index_mbp_s": length / 1e6 / (ph["esa.total_ms"] / 1e3),
                             "anchor_ms": ph["anchor.total_ms"], "walk_ms": ph["anchor.walk_ms"],
                             "loopA_mbp_s_per_gpu": shard.bases / 1e6 / (ph["anchor.total_ms"] / 1e3),
                             "rows_ms": ph["rows.ms"], "matrix_ms": matrix_ms, "batches": ph["map.batches"]},
            "roofline_matrix": {
                "kernel": "k_compare_tiles", "bound": "issue (logic + POPC pipes), not hbm: bit-plane tiles are reused from shared memory",
                "algorithmic_bytes": alg_bytes, "achieved": alg_bytes / (matrix_ms * 1e-3) / 1e9 if matrix_ms > 0 else None,
                "unit": "GB/s per GPU (this rank's tiles; x n_gpus for the job)" if world > 1 else "GB/s",
                "hbm_peak": peak, "peak_source": peak_src,
                "note": "achieved counts the bytes seqcmp/revseqcmp would read (2 B per pair and homologous column) "
                        "for the whole matrix over this rank's matrix time; the packed, tiled kernel moves ~0.1 B per "
                        "pair and column, so this legitimately exceeds the HBM peak",
                "pair_words": pairs * W,
                "alu_floor_ms": pairs * W / 32.0 * (16.0 / 3.0) * 0.5 / (148 * 1.965e6) / world,
            },
            "generate_s": gen_s,
        }
        if entry["roofline_matrix"]["achieved"] and world > 1:
            entry["roofline_matrix"]["achieved"] /= world  # this rank did 1 / world of the pairs
    counts_host = counts_dev.cpu().numpy() if rank == 0 else None
    pipe.close()
    del pipe, shard, counts_dev
    torch.cuda.empty_cache()

    # rank 0: sampled check against the CPU checker, and the same workload on one GPU alone
    if rank == 0:
        all_host = None
        if not args.no_check or world > 1:
            stride = (length + 1 + 15) // 16 * 16
            all_host = torch.zeros(ntot * stride + 64, dtype=torch.uint8)
            if world > 1:
                all_host = all_host.pin_memory()
            base = all_host.data_ptr()

            def gen(i):
                s = spec(i)
                simgen.simf(seed, s[0], length, s[1], out=base + i * stride)

            with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
                list(ex.map(gen, range(ntot)))
            ptrs = [base + i * stride for i in range(ntot)]
            lens = np.full(ntot, length, dtype=np.uint64)
        if not args.no_check:
            try:
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                import oracle_lib

                lib = oracle_lib.best()
                rng = np.random.default_rng(12345 + seed)
                rows = sorted(set([0, ntot - 1] + [int(x) for x in rng.integers(0, ntot, size=30)]))
                t0 = time.perf_counter()
                res = lib.process_rows(ptrs, rows, 0, 0, threads=os.cpu_count() or 1, lens=lens)
                dt = time.perf_counter() - t0
                ok = bool((res["subst"].astype(np.int64) == counts_host[0][rows]).all()
                          and (res["homologs"].astype(np.int64) == counts_host[1][rows]).all())
                entry["check"] = {"sampled_rows": len(rows), "cells": len(rows) * ntot, "rows_equal_reference": ok,
                                  "checker": lib.kind, "seconds": dt, "cpu_anchor_s": res["timings"]["anchor"],
                                  "cpu_esa_s": res["timings"]["esa"],
                                  "what": "every sequence mapped by the checker, the sampled rows of both count matrices compared cell by cell"}
            except Exception as e:
                entry["check"] = {"unavailable": str(e)}
        if world > 1:
            # the whole workload on this one GPU, same run: what the N-GPU time is measured against
            try:
                solo = pb.Context(local_rank)
                solo.set_stream(torch.cuda.current_stream().cuda_stream)
                out = (np.zeros((ntot, ntot), np.uint64), np.zeros((ntot, ntot), np.uint64))
                ts = []
                for it in range(3):
                    flush.zero_()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    solo.process_ptrs(ptrs, lens, 0, 0, out)
                    ts.append(1e3 * (time.perf_counter() - t0))
                entry["single_gpu_same_run"] = {
                    "e2e_ms": min(ts[1:]), "call": "phylo_process on rank 0 alone (pinned host pointers)",
                    "same_counts": bool((out[0].astype(np.int64) == counts_host[0]).all()
                                        and (out[1].astype(np.int64) == counts_host[1]).all()),
                    "e2e_speedup": min(ts[1:]) / entry["e2e"]["ms_per_step"]}
                solo.close()
            except Exception as e:
                entry["single_gpu_same_run"] = {"unavailable": str(e)}
        del all_host
    barrier()
    return entry


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
