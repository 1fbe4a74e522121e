#!/usr/bin/env python
"""Run one of BASELINE.json's configurations at full (or scaled) size through the C ABI on one
GPU, time it, and check every count against the CPU checker (oracle/_ref or the port).

    python tests/run_full_size.py --config 3            # 100 x 5 Mbp
    python tests/run_full_size.py --config 4            # 1000 x 3 Mbp
    python tests/run_full_size.py --config 5 --scale 0.1   # 16 x 25 Mbp, multi-contig

Genomes come from the product-side simf generator (phylonium_b200/simgen.py); config 5 is
split into 25 contigs per genome as BASELINE.md §2 / SURVEY.md §8d describe (cut points every
len/25 bases, shifted by 1000 g for odd g; contig c reverse-complemented iff (c + g) % 3 == 0).
This script is test infrastructure like the rest of tests/ (it is not collected by pytest: the
full sizes take minutes): the checker (oracle/) is only used to verify, never on the timed path.
Prints one JSON line (also appended to gpurun_out/configs.jsonl)."""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # oracle_lib

_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT!", b"TGCA!"):
    _COMP[_a] = _b


def revcomp(a: np.ndarray) -> np.ndarray:
    return _COMP[a][::-1]


def config_spec(cfg: int, scale: float):
    if cfg == 1:
        return dict(seed=1, length=int(100_000 * scale), dists=[0.01], contigs=1)
    if cfg == 2:
        return dict(seed=2, length=int(5_000_000 * scale), dists=[0.001, 0.002, 0.005, 0.01, 0.02, 0.03, 0.05], contigs=1)
    if cfg == 3:
        return dict(seed=3, length=int(5_000_000 * scale), dists=[0.05 * i / 99 for i in range(1, 100)], contigs=1)
    if cfg == 4:
        return dict(seed=4, length=int(3_000_000 * scale), dists=[0.05 * i / 999 for i in range(1, 1000)], contigs=1)
    if cfg == 5:
        return dict(seed=5, length=int(250_000_000 * scale), dists=[0.02 * i / 15 for i in range(1, 16)], contigs=25)
    raise SystemExit("config must be 1..5")


def split_contigs(seq: np.ndarray, g: int, ncontig: int) -> np.ndarray:
    L = len(seq)
    step = L // ncontig
    cuts = [c * step + (1000 * g if g % 2 else 0) for c in range(1, ncontig)] + [L]
    out = np.empty(L + ncontig - 1, dtype=np.uint8)
    w, prev = 0, 0
    for c, cut in enumerate(cuts):
        piece = seq[prev:cut]
        if (c + g) % 3 == 0:
            piece = revcomp(piece)
        out[w : w + len(piece)] = piece
        w += len(piece)
        if c + 1 < len(cuts):
            out[w] = ord("!")
            w += 1
        prev = cut
    assert w == len(out)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--scale", type=float, default=1.0, help="scale the genome length (1.0 = BASELINE.json's size)")
    ap.add_argument("--genomes", type=int, default=0, help="use only the first N genomes (0 = all)")
    ap.add_argument("--no-check", action="store_true", help="skip the CPU checker")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--flags", type=int, default=0, help="4 = complete deletion")
    ap.add_argument("--pageable", action="store_true", help="hand over ordinary (pageable) host memory instead of pinned")
    ap.add_argument("--option", action="append", default=[], help="library option key=value (repeatable)")
    args = ap.parse_args()

    import torch

    import phylonium_b200 as pb
    from phylonium_b200 import simgen

    spec = config_spec(args.config, args.scale)
    ds = [0.0] + spec["dists"]
    if args.genomes:
        ds = ds[: args.genomes]
    N, L = len(ds), spec["length"]
    threads = os.cpu_count() or 1

    t0 = time.perf_counter()
    genomes = [None] * N

    def gen(i):
        buf = np.empty(L, dtype=np.uint8)
        simgen.simf(spec["seed"], spec["seed"] + i, L, ds[i], out=buf.ctypes.data)
        if spec["contigs"] > 1:
            buf = split_contigs(buf, i, spec["contigs"])
        genomes[i] = buf

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(gen, range(N)))
    gen_s = time.perf_counter() - t0
    bases = int(sum(len(g) for g in genomes))

    # pinned host copies: what a host program would hand to phylo_process
    stride = [(len(g) + 1 + 15) // 16 * 16 for g in genomes]
    host = torch.zeros(sum(stride) + 64, dtype=torch.uint8).pin_memory()
    hview = host.numpy()
    ptrs, off = [], 0
    for g, st in zip(genomes, stride):
        hview[off : off + len(g)] = g
        ptrs.append(host.data_ptr() + off)
        off += st
    lens = np.array([len(g) for g in genomes], dtype=np.uint64)
    if args.pageable:  # numpy's own allocations: what a std::string would be
        ptrs = [g.ctypes.data for g in genomes]

    ctx = pb.Context(0)
    for kv in args.option:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    out = (np.zeros((N, N), np.uint64), np.zeros((N, N), np.uint64))
    times = []
    for rep in range(args.reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.process_ptrs(ptrs, lens, 0, args.flags, out)
        times.append(time.perf_counter() - t0)
    first_s, best_s = times[0], min(times[1:])
    ctx.set_option("timings", 1)
    ctx.process_ptrs(ptrs, lens, 0, args.flags, out)
    phases = {k: round(ctx.stat(k), 3) for k in (
        "esa.total_ms", "esa.sort_ms", "esa.lcp_ms", "esa.cld_ms", "esa.table_ms", "anchor.total_ms", "anchor.walk_ms",
        "rows.ms", "compare.ms", "esa.packed", "esa.dirty", "esa.tie_groups", "esa.tied", "esa.key_chars")}
    ctx.set_option("timings", 0)
    subst, homol = out[0].copy(), out[1].copy()
    mem_gb = torch.cuda.mem_get_info(0)
    used_gb = (mem_gb[1] - mem_gb[0]) / 2**30

    # size-independent properties (hold at any size, checked even without the CPU checker)
    props = {
        "symmetric": bool((subst == subst.T).all() and (homol == homol.T).all()),
        "diagonal_zero": bool((np.diag(subst) == 0).all() and (np.diag(homol) == 0).all()),
        "subst_le_homologs": bool((subst <= homol).all()),
        "reference_row_covers": float(homol[0, 1:].min() / max(1, L)) if N > 1 else None,
    }

    check = None
    if not args.no_check:
        import oracle_lib

        lib = oracle_lib.best()
        t0 = time.perf_counter()
        want = lib.process([g.tobytes() for g in genomes], 0, args.flags, threads=threads, timed=True)
        cpu_s = time.perf_counter() - t0
        check = {
            "kind": lib.kind, "cores": threads, "seconds": round(cpu_s, 3),
            "phases_s": {k: round(float(v), 3) for k, v in want["timings"].items()},
            "homologs_equal": bool(np.array_equal(homol, want["homologs"])),
            "subst_equal": bool(np.array_equal(subst, want["subst"])),
        }

    line = {
        "config": args.config, "scale": args.scale, "genomes": N, "genome_length": L, "contigs": spec["contigs"],
        "bases": bases, "flags": args.flags, "generate_s": round(gen_s, 2),
        "gpu_first_call_s": round(first_s, 4), "gpu_best_s": round(best_s, 4),
        "e2e_mbp_s": round(bases / 1e6 / best_s, 1), "device_mem_used_gb": round(used_gb, 2),
        "phases_ms": phases, "properties": props, "check": check,
        "call": "phylo_process (host pointers, %s), wall clock around the call" % ("pageable" if args.pageable else "pinned"),
    }
    print(json.dumps(line), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")
    ok = all(v for k, v in props.items() if isinstance(v, bool)) and (check is None or (check["homologs_equal"] and check["subst_equal"]))
    ctx.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
