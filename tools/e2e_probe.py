"""Where phylo_process spends its time on configs[1] (8 x 5 Mbp, pinned host buffers): wall clock
per call and the device-clock marks the library records (index built / compared / done), for the
raw and the packed upload and for the resident second pass.

    python tools/e2e_probe.py [--genomes 8] [--length 5000000] [--reps 20]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=8)
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--options", default="", help="k=v,k=v set on the context")
    ap.add_argument("--staged", action="store_true")
    a = ap.parse_args()
    if a.staged:
        return staged(a)

    import numpy as np
    import torch

    import bench
    import phylonium_b200 as pb
    from phylonium_b200 import sharding, simgen

    sys.argv = [sys.argv[0]]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    G, L = a.genomes, a.length
    plan = sharding.make_plan(G, 1, 0)
    specs = [bench.genome_spec(g) for g in plan.genomes()]
    shard = bench.Shard(torch, simgen, dev, bench.SIMF_SEED, specs, L, 8)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = (np.zeros((G, G), np.uint64), np.zeros((G, G), np.uint64))
    keys = ("process.gpu_index_ms", "process.gpu_compared_ms", "process.gpu_done_ms", "process.host_index_done_ms",
            "process.host_map_done_ms", "process.host_done_ms", "process.h2d_bytes")
    import ctypes
    pageable = [np.frombuffer((ctypes.c_char * L).from_address(p), dtype=np.uint8).copy() for p in shard.ptrs]
    pptrs = [x.ctypes.data for x in pageable]
    for name, opts, again in (("raw", {"upload_raw": 1}, False), ("packed", {"upload_raw": -1}, False), ("resident", {}, True),
                              ("pageable", {}, False), ("pageable", {}, False)):
        ctx = pb.Context(0)
        ctx.set_stream(stream.cuda_stream)
        for kv in filter(None, a.options.split(",")):
            k, v = kv.split("=")
            ctx.set_option(k, int(v))
        for k, v in opts.items():
            ctx.set_option(k, v)
        ptrs = pptrs if name == "pageable" else shard.ptrs
        fn = (lambda: ctx.process_again(0, 0, out)) if again else (lambda: ctx.process_ptrs(ptrs, shard.lens, 0, 0, out))
        ctx.process_ptrs(ptrs, shard.lens, 0, 0, out)
        for _ in range(3):
            fn()
        acc = {k: 0.0 for k in keys}
        wall = 0.0
        for _ in range(a.reps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            wall += 1e3 * (time.perf_counter() - t0)
            for k in keys:
                acc[k] += ctx.stat(k)
        print("graph: instantiated", ctx.stat("esa.graph_instantiated"), "updated", ctx.stat("esa.graph_updated"))
        print(name, "wall_ms %.3f" % (wall / a.reps), " ".join("%s %.3f" % (k.split(".")[1], v / a.reps) for k, v in acc.items()), flush=True)
        ctx.close()


def staged(a):
    """the calls a rank of a sharded run makes (bench.py: Pipeline.step_from_host), on one GPU, timed one by one"""
    import numpy as np
    import torch

    import bench
    import phylonium_b200 as pb
    from phylonium_b200 import sharding, simgen

    sys.argv = [sys.argv[0]]
    args = bench.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    G, L = a.genomes, a.length
    plan = sharding.make_plan(G, 1, 0)
    specs = [bench.genome_spec(g) for g in plan.genomes()]
    shard = bench.Shard(torch, simgen, dev, bench.SIMF_SEED, specs, L, 8)
    pipe = bench.Pipeline((torch, None, pb, sharding), args, dev, 0, 0, 1, plan, shard, shard.host[:L], L, "replicate", "push")
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(3):
        pipe.step_from_host()
    t = [0.0] * 5
    for _ in range(a.reps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe._index(host=True)
        t1 = time.perf_counter()
        pipe.ctx.map_queries_ptrs(shard.ptrs, shard.lens, pipe.thr)
        t2 = time.perf_counter()
        pipe._matrix()
        t3 = time.perf_counter()
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        for k, (x, y) in enumerate(((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t0, t4))):
            t[k] += 1e3 * (y - x) / a.reps
    print("staged: index call %.3f  map call %.3f  matrix call %.3f  final sync %.3f  total %.3f ms" % tuple(t))


if __name__ == "__main__":
    main()
