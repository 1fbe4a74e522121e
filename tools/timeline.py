"""Kernel timeline of one headline step (configs[1]) through CUPTI (torch.profiler collects the
activity records of every kernel in the process, the library's included): where the GPU idles.

    python tools/timeline.py [--mode step|again] > gpurun_out/timeline.txt

mode step : bench.py's device-resident step (three library calls from Python)
mode again: phylo_process_again (one C call on the resident sequences)
mode process: phylo_process on pinned host buffers (the e2e call)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="step")
    ap.add_argument("--genomes", type=int, default=8)
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--csv", default=None)
    a = ap.parse_args()

    import numpy as np
    import torch
    from torch.profiler import ProfilerActivity, profile

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
    ctx = pipe.ctx
    out = (np.zeros((G, G), np.uint64), np.zeros((G, G), np.uint64))
    if a.mode == "again":
        ctx.process_ptrs(shard.ptrs, shard.lens, 0, 0, out)
        fn = lambda: ctx.process_again(0, 0, out)
    elif a.mode == "process":
        fn = lambda: ctx.process_ptrs(shard.ptrs, shard.lens, 0, 0, out)
    else:
        fn = pipe.step
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            flush.zero_()
            torch.cuda.synchronize()
            fn()
            torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    # split into steps at the flush memsets (>= 512 MiB fill: by far the longest memset)
    steps, cur = [], None
    for e in ev:
        dur = e.time_range.end - e.time_range.start
        if ("emset" in e.name or "FillFunctor" in e.name) and dur > 50:
            cur = []
            steps.append(cur)
            continue
        if cur is not None:
            cur.append(e)
    rows = steps[-1]
    t0 = rows[0].time_range.start
    busy_until, gaps, busy = t0, [], 0.0
    lines = []
    for e in rows:
        s, t = e.time_range.start, e.time_range.end
        gap = s - busy_until
        if gap > 0:
            gaps.append((gap, e.name, s - t0))
        busy += max(0.0, t - max(s, busy_until))
        busy_until = max(busy_until, t)
        lines.append((s - t0, t - s, gap, e.name[:70]))
    span = busy_until - t0
    print(f"# mode {a.mode}: {len(rows)} device activities, span {span:.1f} us, busy {busy:.1f} us, idle {span - busy:.1f} us")
    print(f"{'start':>9} {'dur':>8} {'gap':>7}  name")
    for s, d, g, n in lines:
        print(f"{s:9.1f} {d:8.1f} {max(g, 0):7.1f}  {n}")
    print("# largest gaps")
    for g, n, s in sorted(gaps, reverse=True)[:12]:
        print(f"#   {g:7.1f} us before {n[:60]} at {s:.1f}")


if __name__ == "__main__":
    main()
