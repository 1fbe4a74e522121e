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
    per_rank: int   # row-store slots per rank (the last ones of a rank may be padding)
    first: int      # first row-store slot of this rank
    count: int      # genomes this rank really owns
    layout: str = "block"  # "block": genome g on rank g // per_rank; "interleaved": on rank g % world

    @property
    def padded_total(self) -> int:
        return self.per_rank * self.world

    def genomes(self) -> List[int]:
        """global indices of this rank's genomes, in the order of its row-store slots"""
        if self.layout == "block":
            return list(range(self.first, self.first + self.count))
        return list(range(self.rank, self.total, self.world))

    def slot_of(self, genome: int) -> int:
        """row-store slot (= row/column of the count matrix) of a genome"""
        if self.layout == "block":
            return genome
        return (genome % self.world) * self.per_rank + genome // self.world

    def slots(self) -> List[int]:
        """slot of genome 0, 1, ... total - 1: indexing the slot-ordered matrix with it on both
        axes gives the matrix in genome order without the padding slots"""
        return [self.slot_of(g) for g in range(self.total)]


def make_plan(total: int, world: int, rank: int, layout: str = "block") -> ShardPlan:
    """block: contiguous runs of genomes per rank.  interleaved: genome g on rank g % world —
    when the divergence from the reference grows with the genome index (BASELINE.json's
    configs) the anchoring time does too, and dealing the genomes round-robin is what keeps
    the ranks balanced.  Either way a rank's rows are one contiguous slice of the store."""
    per_rank = (total + world - 1) // world
    if layout == "block":
        first = min(rank * per_rank, total)
        count = max(0, min(per_rank, total - first))
    elif layout == "interleaved":
        count = len(range(rank, total, world))
    else:
        raise ValueError(layout)
    return ShardPlan(world, rank, total, per_rank, rank * per_rank, count, layout)


def owner_of(plan: ShardPlan, genome: int) -> int:
    return genome // plan.per_rank if plan.layout == "block" else genome % plan.world


NUM_SMS = 148
STEP_WORDS = 96  # chunk granularity in words (compare.cu: whole steps of both paths, ROW_BLK)


def tile_side(n_genomes: int) -> int:
    """genomes per side of a block's tile (compare.cu: CT)"""
    return 8 if n_genomes <= 24 else 16


BAND = 12  # tile rows per band (CMP_BAND, csrc/tile_order.h)


def tile_pairs(tile_begin: int, tile_end: int) -> List[tuple]:
    """the tile pairs (ti <= tj, tile_begin <= tj < tile_end) in the order k_compare_tiles visits
    them: bands of BAND tile rows, column by column inside a band — so that the blocks running at
    the same time work on a compact patch of the matrix (csrc/tile_order.h)"""
    out = []
    for ti0 in range(0, tile_end, BAND):
        for tj in range(max(tile_begin, ti0), tile_end):
            for ti in range(ti0, min(ti0 + BAND, tj + 1)):
                out.append((ti, tj))
    return out


def compare_units(n_genomes: int, words: int, world: int):
    """(tile pairs, chunks, chunk_words) of the all-pairs stage — the same arithmetic as
    compare_all_device (compare.cu); `words` = 32-bit words per row plane"""
    tile = tile_side(n_genomes)
    side = (n_genomes + tile - 1) // tile
    pairs = tile_pairs(0, side)
    want_blocks = NUM_SMS * 8 * world
    chunks = (want_blocks + len(pairs) - 1) // max(1, len(pairs))
    chunks = max(1, min(chunks, (words + 255) // 256))
    chunk_words = (words + chunks - 1) // chunks
    chunk_words = (chunk_words + STEP_WORDS - 1) // STEP_WORDS * STEP_WORDS
    chunks = (words + chunk_words - 1) // chunk_words
    return pairs, chunks, chunk_words


def units_of_rank(n_genomes: int, words: int, rank: int, world: int) -> List[tuple]:
    """the (ti, tj, chunk) work units rank `rank` computes: unit u = pair * chunks + chunk; every
    rank takes one contiguous range of ceil(units / world) — same as compare_all_device"""
    pairs, chunks, _ = compare_units(n_genomes, words, world)
    units = [(ti, tj, c) for (ti, tj) in pairs for c in range(chunks)]
    per_rank = (len(units) + world - 1) // world
    return units[rank * per_rank:(rank + 1) * per_rank]


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


def setup_push(ctx, rank: int, world: int) -> None:
    """Row exchange without a collective (include/phylonium_b200.h, phylo_rows_ipc_*): every
    rank learns the address of every other rank's row store; from then on the mapping copies
    each batch of rows into all of them over NVLink while it maps the next batch.  Call after
    ctx.rows_configure (with an index in place) and again whenever the store changes size."""
    mine = ctx.rows_ipc_export()
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    ctx.rows_ipc_import(handles, rank)


_token = {}


def rows_barrier(device) -> None:
    """stream-ordered barrier across ranks: behind it, on this rank's current stream, the row
    pushes of all ranks have landed (each rank's stream is ordered behind its own pushes)"""
    t = _token.get(device)
    if t is None:
        t = _token[device] = torch.zeros(1, dtype=torch.int32, device=device)
    dist.all_reduce(t)


def genome_order(counts: torch.Tensor, plan: ShardPlan) -> torch.Tensor:
    """counts: (..., padded_total * padded_total) in row-store slot order -> (..., total, total)
    in genome order"""
    n = plan.padded_total
    idx = torch.as_tensor(plan.slots(), device=counts.device)
    m = counts.reshape(*counts.shape[:-1], n, n)
    return m.index_select(-2, idx).index_select(-1, idx)


def reduce_matrix(subst: torch.Tensor, homol: torch.Tensor) -> None:
    """Sum of the per-rank partial count matrices (int64 views of the uint64 counts)."""
    dist.all_reduce(subst, op=dist.ReduceOp.SUM)
    dist.all_reduce(homol, op=dist.ReduceOp.SUM)
