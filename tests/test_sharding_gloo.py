"""Host-side logic of the multi-GPU path on CPU: two processes, gloo backend.

Checks the shard plan, that the work-unit enumeration gives every (genome pair, column chunk)
to exactly one rank (so that summing the partial matrices reproduces the full matrix) and the
in-place all-gather of the genome-major row store."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phylonium_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = sharding.make_plan(total, world, rank)
        # --- rows: every rank fills its slice, then all ranks must hold everything
        bpg = 24
        store = torch.zeros(plan.padded_total * bpg, dtype=torch.uint8)
        for g in range(plan.first, plan.first + plan.count):
            store[g * bpg : (g + 1) * bpg] = (g * 7 + 1) % 251
        sharding.allgather_store(store, plan, bpg)
        want = torch.zeros_like(store)
        for g in range(total):
            want[g * bpg : (g + 1) * bpg] = (g * 7 + 1) % 251
        ok_rows = bool((store == want).all())
        # --- matrix: each rank adds up the contributions of its work units only
        n = plan.padded_total
        words = 1000
        pairs, chunks, _ = sharding.compare_units(n, words, world)
        rng = np.random.default_rng(5)
        contrib = rng.integers(1, 1000, size=(chunks, n, n))  # what chunk c adds to cell (i, j)
        contrib = np.triu(contrib, 1)
        full = contrib.sum(axis=0)
        full = full + full.T
        part = np.zeros_like(full)
        T = sharding.tile_side(n)
        for ti, tj, c in sharding.units_of_rank(n, words, rank, world):
            for i in range(ti * T, min(n, ti * T + T)):
                for j in range(tj * T, min(n, tj * T + T)):
                    if i < j:
                        part[i, j] += contrib[c, i, j]
                        part[j, i] += contrib[c, i, j]
        a = torch.from_numpy(part.copy())
        b = torch.from_numpy(part.copy())
        sharding.reduce_matrix(a, b)
        ok_matrix = bool((a.numpy() == full).all() and (b.numpy() == full).all())
        results[rank] = (ok_rows, ok_matrix, plan.first, plan.count)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [16, 13, 3])
def test_two_ranks(total):
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, total, results), nprocs=world, join=True)
        got = dict(results)
    assert set(got) == {0, 1}
    assert all(v[0] and v[1] for v in got.values()), got
    assert sum(v[3] for v in got.values()) == total
    assert got[0][2] == 0 and got[1][2] == (total + 1) // 2


def test_plan_and_tiles_cover_everything():
    for total in (1, 2, 7, 8, 64, 1000):
        for world in (1, 2, 4, 8):
            plans = [sharding.make_plan(total, world, r) for r in range(world)]
            assert sum(p.count for p in plans) == total
            owned = [g for p in plans for g in range(p.first, p.first + p.count)]
            assert owned == list(range(total))
            n = plans[0].padded_total
            for words in (1, 255, 4000, 156252):
                pairs, chunks, chunk_words = sharding.compare_units(n, words, world)
                assert chunks * chunk_words >= words > (chunks - 1) * chunk_words
                seen = set()
                for r in range(world):
                    for u in sharding.units_of_rank(n, words, r, world):
                        assert u not in seen
                        seen.add(u)
                side = (n + sharding.tile_side(n) - 1) // sharding.tile_side(n)
                assert len(seen) == side * (side + 1) // 2 * chunks


def test_interleaved_plan_and_genome_order():
    """interleaved layout: genome g lives on rank g % world at slot rank * per_rank + g // world;
    the slot-ordered matrix indexed with plan.slots() is the matrix in genome order"""
    for total in (1, 2, 7, 8, 13, 100):
        for world in (1, 2, 4, 8):
            plans = [sharding.make_plan(total, world, r, "interleaved") for r in range(world)]
            assert sum(p.count for p in plans) == total
            owned = sorted(g for p in plans for g in p.genomes())
            assert owned == list(range(total))
            slots = plans[0].slots()
            assert len(set(slots)) == total and max(slots) < plans[0].padded_total
            for p in plans:
                for k, g in enumerate(p.genomes()):
                    assert p.slot_of(g) == p.first + k and sharding.owner_of(p, g) == p.rank
                    assert p.first + k < p.first + p.per_rank
            # a matrix whose cell (slot a, slot b) encodes the genomes that live there
            n = plans[0].padded_total
            genome_at = {s: g for g, s in enumerate(slots)}
            m = torch.full((2, n * n), -1, dtype=torch.int64)
            for a in range(n):
                for b in range(n):
                    if a in genome_at and b in genome_at:
                        m[:, a * n + b] = genome_at[a] * 1000 + genome_at[b]
            out = sharding.genome_order(m, plans[0])
            assert out.shape == (2, total, total)
            want = torch.arange(total)[:, None] * 1000 + torch.arange(total)[None, :]
            assert bool((out[0] == want).all() and (out[1] == want).all())
