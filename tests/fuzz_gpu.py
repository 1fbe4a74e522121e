#!/usr/bin/env python
"""Randomised parity hunt on a GPU (not collected by pytest; run it by hand):

    python tests/fuzz_gpu.py --seconds 240 --seed 1

Every round draws a small family of genomes (random reference with optional contigs, repeats
and low-complexity stretches; queries that are mutated, cut, inverted, shuffled, truncated,
reverse complemented, unrelated or empty), random tuning options (chunk, cap, K, key length,
sorter, batch size, scan mode, complete deletion) and requires the counts of phylo_process —
and, every few rounds, every homology list and the index arrays — to equal the checker's.
Prints the failing recipe (seed, round) and exits 1 on the first difference."""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import datasets  # noqa: E402
import oracle_lib  # noqa: E402


def make_family(rng, big=False):
    n = int(rng.integers(1, 40000)) if rng.random() < 0.85 else int(rng.integers(1, 40))
    if big:
        n = int(rng.integers(50000, 400000))
    kind = rng.random()
    if kind < 0.15:
        unit = datasets.random_dna(rng, int(rng.integers(1, 12)))
        ref = (unit * (n // len(unit) + 1))[:n]
    else:
        ref = datasets.random_dna(rng, n)
    if n > 200 and rng.random() < 0.4:  # a repeat
        a, l = int(rng.integers(0, n // 2)), int(rng.integers(20, max(21, n // 4)))
        b = int(rng.integers(0, n))
        ref = ref[:b] + ref[a : a + l] + ref[b:]
    if n > 50 and rng.random() < 0.3:  # low complexity
        b = int(rng.integers(0, len(ref)))
        ref = ref[:b] + b"A" * int(rng.integers(5, 120)) + ref[b:]
    if len(ref) > 10 and rng.random() < 0.4:  # contigs (sometimes more than the packed sorter takes)
        for _ in range(int(rng.integers(1, 8)) if not big or rng.random() < 0.7 else int(rng.integers(900, 1300))):
            b = int(rng.integers(0, len(ref) + 1))
            ref = ref[:b] + b"!" + ref[b:]
    flat = ref.replace(b"!", b"") or b"A"
    genomes = [ref]
    many = rng.random() < 0.12 and n < 6000  # enough genomes for 16 x 16 tiles and several tile columns
    for _ in range(int(rng.integers(25, 45)) if many else int(rng.integers(1, 9))):
        r = rng.random()
        q = datasets.mutate(rng, flat, float(rng.choice([0.0, 0.001, 0.01, 0.03, 0.08, 0.2])))
        if r < 0.15:
            q = datasets.indel(rng, q, int(rng.integers(1, 12)), int(rng.integers(1, 80)))
        elif r < 0.3:
            bl = max(1, len(q) // int(rng.integers(2, 9)))
            blocks = [q[i : i + bl] for i in range(0, len(q), bl)]
            order = rng.permutation(len(blocks))
            q = b"".join(datasets.revcomp(blocks[k]) if rng.random() < 0.4 else blocks[k] for k in order)
        elif r < 0.4:
            q = datasets.revcomp(q)
        elif r < 0.5:
            q = q[: int(rng.integers(0, len(q) + 1))]
        elif r < 0.55:
            q = datasets.random_dna(rng, int(rng.integers(1, 3000)))
        elif r < 0.6:
            q = b""
        if len(q) > 4 and rng.random() < 0.25:
            b = int(rng.integers(0, len(q) + 1))
            q = q[:b] + b"!" + q[b:]
        genomes.append(q)
    if rng.random() < 0.3:
        genomes.append(ref)
    return genomes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--big", action="store_true", help="references of 50..400 kbp, sometimes with > 1000 contigs")
    args = ap.parse_args()
    import phylonium_b200 as pb

    oracle = oracle_lib.best()
    t_end = time.time() + args.seconds
    rounds = 0
    print_recipe_on_error = ""
    try:
        while time.time() < t_end:
            rng = np.random.default_rng([args.seed, rounds])
            genomes = make_family(rng, args.big)
            ref_index = int(rng.integers(0, len(genomes)))
            if len(genomes[ref_index]) == 0:
                ref_index = 0
            opts = dict(
                chunk=int(rng.choice([32, 64, 256, 2048])), cap=int(rng.choice([0, 64, 300])),
                kmer_k=int(rng.choice([-1, 0, 1, 5, 9])), key_chars=int(rng.choice([0, 0, 2, 7, 16, 21])),
                sort_path=int(rng.choice([0, 0, 1])), scan_mode=int(rng.choice([1, 1, 0])),
                map_batch_bytes=int(rng.choice([1, 5000, 512 << 20])), table_direct=int(rng.choice([0, 2])),
                # (with keep_raw the rows of a batch are built after its lists are final; without,
                # before the host has seen them — and again if they were not)
                keep_raw=1 if rounds % 4 == 0 else 0, esa_graph=int(rng.choice([1, 1, 0])),
                map_graph=int(rng.choice([0, 1, 2, 2])),
                upload_raw=int(rng.choice([0, 0, 1, -1])), compare_path=int(rng.choice([0, 0, 1])),
                esa_speculative=int(rng.choice([1, 1, 0])), stage_threads=int(rng.choice([0, 1, 3])),
            )
            via_ingest = rng.random() < 0.25
            flags = int(rng.choice([0, 0, 4]))
            recipe = (f"seed={args.seed} round={rounds} ref_index={ref_index} flags={flags} opts={opts} "
                      f"lens={[len(g) for g in genomes]}")
            want = oracle.process(genomes, ref_index, flags, threads=4)
            print_recipe_on_error = recipe
            with pb.Context(**opts) as ctx:
                if via_ingest:  # sequences handed over one by one, then process() on what is resident
                    ctx.ingest(genomes, max_lens=[len(g) + int(rng.integers(0, 50)) for g in genomes],
                               lanes=int(rng.integers(1, 4)), threads=int(rng.integers(1, 4)))
                    subst, homol = ctx.process_again(ref_index, flags)
                else:
                    subst, homol = ctx.process(genomes, ref_index, flags)
                if not (np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])):
                    print("COUNTS DIFFER:", recipe)
                    return 1
                if rounds % 4 == 0:
                    ref = genomes[ref_index]
                    thr = oracle.threshold(ref)
                    esa = oracle.esa(ref)
                    arr, got = esa.arrays(), ctx.esa_arrays()
                    for k in ("SA", "LCP", "CLD", "FVC"):
                        if not np.array_equal(got[k], arr[k]):
                            print("INDEX DIFFERS:", k, recipe)
                            return 1
                    for k, q in enumerate(genomes):
                        raw = esa.anchor_homologies(thr, q) if len(q) else np.zeros(0, oracle_lib.HOM_DTYPE)
                        if not np.array_equal(ctx.homologies(k, raw=True), raw):
                            print("RAW HOMOLOGIES DIFFER:", k, recipe)
                            return 1
                        if not np.array_equal(ctx.homologies(k), oracle.sort_filter(raw)):
                            print("FILTERED HOMOLOGIES DIFFER:", k, recipe)
                            return 1
                if rounds % 3 == 1:
                    # second pass on the resident sequences with another reference: the index build's
                    # graph of the first pass is brought up to date or instantiated anew
                    ref2 = int(rng.integers(0, len(genomes)))
                    if len(genomes[ref2]):
                        want2 = oracle.process(genomes, ref2, flags, threads=4)
                        s2, h2 = ctx.process_again(ref2, flags)
                        if not (np.array_equal(s2, want2["subst"]) and np.array_equal(h2, want2["homologs"])):
                            print("COUNTS OF THE SECOND PASS DIFFER:", f"ref2={ref2}", recipe)
                            return 1
            rounds += 1
    except Exception:
        print("EXCEPTION:", print_recipe_on_error)
        raise
    finally:
        with pb.Context(sort_path=0, scan_mode=1, map_batch_bytes=512 << 20, table_direct=0):
            pass
    print(f"fuzz ok: {rounds} rounds, seed {args.seed}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
