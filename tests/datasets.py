"""Deterministic input sets for the parity tests (numpy RNG, fixed seeds).

Each set is a list of genome byte strings over {A,C,G,T,!} ('!' joins contigs,
/root/reference/src/sequence.cxx:171-199); genome 0 is used as the reference unless
a test says otherwise.  The sets cover what SURVEY.md §4 says the reference never
tests: reverse-strand homologies, multi-contig inputs, repeats, indels, identical and
unrelated genomes, and the degenerate sizes.
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT!", b"TGCA!"):
    _COMP[_a] = _b


def random_dna(rng: np.random.Generator, n: int) -> bytes:
    return ACGT[rng.integers(0, 4, size=n)].tobytes()


def revcomp(s: bytes) -> bytes:
    return _COMP[np.frombuffer(s, dtype=np.uint8)][::-1].tobytes()


def mutate(rng: np.random.Generator, s: bytes, rate: float) -> bytes:
    """Substitutions only, like simf."""
    a = np.frombuffer(s, dtype=np.uint8).copy()
    hit = np.nonzero((rng.random(len(a)) < rate) & (a != ord("!")))[0]
    for p in hit:
        choices = [c for c in b"ACGT" if c != a[p]]
        a[p] = choices[rng.integers(0, 3)]
    return a.tobytes()


def indel(rng: np.random.Generator, s: bytes, events: int, maxlen: int) -> bytes:
    b = bytearray(s)
    for _ in range(events):
        p = int(rng.integers(0, max(1, len(b))))
        l = int(rng.integers(1, maxlen + 1))
        if rng.random() < 0.5:
            del b[p : p + l]
        else:
            b[p:p] = random_dna(rng, l)
    return bytes(b)


def simple_pair(seed=1, n=20000, d=0.02):
    rng = np.random.default_rng(seed)
    r = random_dna(rng, n)
    return [r, mutate(rng, r, d)]


def divergent_set(seed=2, n=30000, rates=(0.001, 0.01, 0.03, 0.06, 0.12)):
    rng = np.random.default_rng(seed)
    r = random_dna(rng, n)
    return [r] + [mutate(rng, r, d) for d in rates]


def multi_contig_set(seed=3, n=24000):
    """3-contig reference; queries with reversed / swapped / re-cut contigs."""
    rng = np.random.default_rng(seed)
    c = [random_dna(rng, n // 3) for _ in range(3)]
    ref = b"!".join(c)
    q1 = b"!".join([mutate(rng, c[1], 0.02), revcomp(mutate(rng, c[0], 0.02)), mutate(rng, c[2], 0.02)])
    q2 = revcomp(mutate(rng, ref, 0.03))  # whole genome on the other strand, '!' kept
    flat = mutate(rng, b"".join(c), 0.01)
    q3 = flat[: n // 2] + b"!" + flat[n // 2 :]  # different cut points
    q4 = b"!".join([mutate(rng, c[0], 0.04), mutate(rng, c[1], 0.04), mutate(rng, c[2], 0.04)])  # same cuts
    q5 = b"!".join([revcomp(mutate(rng, c[2], 0.02)), revcomp(mutate(rng, c[1], 0.02))])  # contig 0 missing
    return [ref, q1, q2, q3, q4, q5]


def bang_vs_base_set(seed=4, n=3000):
    """SURVEY.md A.6 regression: reference X!Y, one genome XAY in one piece, one X!Y."""
    rng = np.random.default_rng(seed)
    x, y = random_dna(rng, n), random_dna(rng, n)
    return [x + b"!" + y, x + b"A" + y, x + b"!" + y, revcomp(x + b"A" + y), revcomp(x + b"!" + y)]


def repeat_set(seed=5, n=20000):
    """Reference with dispersed, tandem and inverted repeats and a low-complexity run."""
    rng = np.random.default_rng(seed)
    unit = random_dna(rng, 700)
    short = random_dna(rng, 37)
    parts = [
        random_dna(rng, n // 4), unit, random_dna(rng, n // 5), short * 12, random_dna(rng, n // 6),
        revcomp(unit), random_dna(rng, n // 7), b"A" * 150, unit, random_dna(rng, n // 8), b"AC" * 90,
        random_dna(rng, n // 9),
    ]
    ref = b"".join(parts)
    q1 = mutate(rng, ref, 0.01)
    q2 = mutate(rng, b"".join(parts[:3] + parts[5:]), 0.02)  # tandem block deleted
    q3 = revcomp(mutate(rng, ref, 0.015))
    q4 = mutate(rng, unit + random_dna(rng, 500) + unit + random_dna(rng, 500) + revcomp(unit), 0.005)
    return [ref, q1, q2, q3, q4]


def indel_set(seed=6, n=30000):
    rng = np.random.default_rng(seed)
    r = random_dna(rng, n)
    return [r] + [indel(rng, mutate(rng, r, d), ev, 60) for d, ev in ((0.005, 10), (0.02, 40), (0.05, 25))]


def rearranged_set(seed=7, n=40000, block=2500):
    """Blocks shuffled, some reversed, some duplicated: many homologies, overlaps on the reference."""
    rng = np.random.default_rng(seed)
    r = random_dna(rng, n)
    blocks = [r[i : i + block] for i in range(0, n, block)]
    out = [r]
    for k in range(3):
        order = rng.permutation(len(blocks))
        q = []
        for b in order:
            piece = mutate(rng, blocks[b], 0.01 * (k + 1))
            if rng.random() < 0.4:
                piece = revcomp(piece)
            q.append(piece)
            if rng.random() < 0.2:
                # duplicate, shifted, so that two homologies overlap on the reference
                s = int(rng.integers(0, block // 2))
                q.append(mutate(rng, blocks[b][s:], 0.01))
        out.append(b"".join(q))
    return out


def identical_and_unrelated_set(seed=8, n=15000):
    rng = np.random.default_rng(seed)
    r = random_dna(rng, n)
    return [r, r, revcomp(r), random_dna(rng, n), r[: n // 2], r[n // 3 :] + random_dna(rng, 2000)]


def tiny_set(seed=9):
    rng = np.random.default_rng(seed)
    r = random_dna(rng, 300)
    return [r, r[:40], b"A", b"ACGT", mutate(rng, r, 0.05), b"!" + r[:100], r[100:200] + b"!"]


def substitution_in_first_bases(seed=10, n=5000):
    """first anchor on diagonal 0 at pos > 0 swallows the prefix."""
    rng = np.random.default_rng(seed)
    r = random_dna(rng, n)
    a = bytearray(r)
    a[5] = ord("A") if a[5] != ord("A") else ord("C")
    return [r, bytes(a), revcomp(bytes(a))]


ALL_SETS = {
    "simple_pair": simple_pair,
    "divergent": divergent_set,
    "multi_contig": multi_contig_set,
    "bang_vs_base": bang_vs_base_set,
    "repeats": repeat_set,
    "indels": indel_set,
    "rearranged": rearranged_set,
    "identical_unrelated": identical_and_unrelated_set,
    "tiny": tiny_set,
    "first_bases": substitution_in_first_bases,
}
