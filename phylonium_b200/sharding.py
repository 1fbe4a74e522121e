"""One-process-per-GPU sharding of the pipeline.

  index    built on rank 0, broadcast in place (S, SA, LCP, CLD, FVC = 14 B per suffix);
           each receiver then builds its descent table locally
  queries  genome g belongs to rank g // per_rank (contiguous blocks, so that a rank's
           reference-coordinate rows are one contiguous slice of the row store)
  rows     all-gather of the bit-plane rows: afterwards every GPU holds all rows
  matrix   work units (16x16 genome tile pair, chunk of reference columns) dealt round-robin
           to ranks; the partial N x N count matrices are summed with one all-reduce

torch.distributed is plumbing only; the collectives move buffers that live inside the
phylo contexts (wrapped without copying).  The same helpers run under the gloo backend
with CPU tensors in the tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch
import torch.distributed as dist


@dataclass
class ShardPlan:
    world: int
    rank: int
    total: int      # genomes over all ranks
    per_rank: int   # genomes per rank (the last ranks may hold padding rows)
    first: int      # global index of this rank's first genome
    count: int      # genomes this rank really owns

    @property
    def padded_total(self) -> int:
        return self.per_rank * self.world


def make_plan(total: int, world: int, rank: int) -> ShardPlan:
    per_rank = (total + world - 1) // world
    first = min(rank * per_rank, total)
    count = max(0, min(per_rank, total - first))
    return ShardPlan(world, rank, total, per_rank, rank * per_rank, count)


def owner_of(plan: ShardPlan, genome: int) -> int:
    return genome // plan.per_rank


NUM_SMS = 148
STEP_WORDS = 32  # words per step (compare.cu: CMP_STEP)


def tile_side(n_genomes: int) -> int:
    """genomes per side of a block's tile (compare.cu: CT)"""
    return 8 if n_genomes <= 24 else 16


def compare_units(n_genomes: int, words: int, world: int):
    """(tile pairs, chunks, chunk_words) of the all-pairs stage — the same arithmetic as
    compare_all_device (compare.cu); `words` = 32-bit words per row plane"""
    tile = tile_side(n_genomes)
    side = (n_genomes + tile - 1) // tile
    pairs = [(ti, tj) for ti in range(side) for tj in range(ti, side)]
    want_blocks = NUM_SMS * 8 * world
    chunks = (want_blocks + len(pairs) - 1) // max(1, len(pairs))
    chunks = max(1, min(chunks, (words + 255) // 256))
    chunk_words = (words + chunks - 1) // chunks
    chunk_words = (chunk_words + STEP_WORDS - 1) // STEP_WORDS * STEP_WORDS
    chunks = (words + chunk_words - 1) // chunk_words
    return pairs, chunks, chunk_words


def units_of_rank(n_genomes: int, words: int, rank: int, world: int) -> List[tuple]:
    """the (ti, tj, chunk) work units rank `rank` computes: unit u = pair * chunks + chunk
    belongs to rank u mod world — same enumeration as k_compare_tiles"""
    pairs, chunks, _ = compare_units(n_genomes, words, world)
    units = [(ti, tj, c) for (ti, tj) in pairs for c in range(chunks)]
    return units[rank::world]


class DeviceBuffer:
    """A region of device memory owned by a phylo context, viewable as a torch tensor."""

    def __init__(self, ptr: int, nbytes: int, device_index: int):
        self.ptr, self.nbytes, self.device_index = ptr, nbytes, device_index
        self.__cuda_array_interface__ = {
            "shape": (nbytes,),
            "typestr": "|u1",
            "data": (ptr, False),
            "version": 2,
        }

    def tensor(self) -> torch.Tensor:
        return torch.as_tensor(self, device=torch.device("cuda", self.device_index))


def broadcast_index(ctx, n: int, src: int, rank: int, device_index: int) -> int:
    """In-place broadcast of the five ESA arrays from `src`; returns the bytes moved."""
    if rank != src:
        ctx.esa_alloc(n)
    moved = 0
    for name, (ptr, nbytes) in ctx.esa_device_arrays().items():
        t = DeviceBuffer(ptr, nbytes, device_index).tensor()
        dist.broadcast(t, src=src)
        moved += nbytes
    if rank != src:
        ctx.esa_finish_import()
    return moved


def allgather_store(store: torch.Tensor, plan: ShardPlan, bytes_per_genome: int) -> None:
    """In-place all-gather of a genome-major byte store: rank r owns rows
    [r * per_rank, (r + 1) * per_rank)."""
    mine = store[plan.first * bytes_per_genome : (plan.first + plan.per_rank) * bytes_per_genome]
    if dist.get_backend() == "gloo":
        parts = [torch.empty_like(mine) for _ in range(plan.world)]
        dist.all_gather(parts, mine.clone())
        for r, part in enumerate(parts):
            store[r * plan.per_rank * bytes_per_genome : (r + 1) * plan.per_rank * bytes_per_genome] = part
    else:
        dist.all_gather_into_tensor(store, mine)


def allgather_rows(ctx, plan: ShardPlan, device_index: int) -> int:
    """All-gather of the row store (each rank contributed rows [first, first + per_rank))."""
    ptr, bytes_per_genome, total = ctx.rows_device()
    assert total == plan.padded_total
    store = DeviceBuffer(ptr, bytes_per_genome * total, device_index).tensor()
    allgather_store(store, plan, bytes_per_genome)
    return bytes_per_genome * total


def reduce_matrix(subst: torch.Tensor, homol: torch.Tensor) -> None:
    """Sum of the per-rank partial count matrices (int64 views of the uint64 counts)."""
    dist.all_reduce(subst, op=dist.ReduceOp.SUM)
    dist.all_reduce(homol, op=dist.ReduceOp.SUM)
