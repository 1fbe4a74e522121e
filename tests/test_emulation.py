"""CPU emulation of the device-side logic (tests/emul/emul.cxx drives the host/device
headers of phylonium_b200/csrc sequentially) against the oracle.

This is where the exactness argument of walk.h is exercised without a GPU: chunked
speculative walking with tiny chunks and caps (so that bridging, open matches and the
serial continuation all occur) must reproduce the reference's sequential walk."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import datasets
import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "emul.cxx")
OUT = os.path.join(HERE, "emul", "_build", "libemul.so")
CSRC = os.path.join(os.path.dirname(HERE), "phylonium_b200", "csrc")


@pytest.fixture(scope="module")
def emul():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("esa_types.h", "esa_search.h", "walk.h", "cld_search.h", "filter.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", SRC, "-o", OUT], check=True)
    return C.CDLL(OUT)


@pytest.fixture(scope="module")
def oracle():
    return oracle_lib.best()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _esa_args(arr):
    return [_ptr(arr["S"]), _ptr(arr["SA"]), _ptr(arr["LCP"]), _ptr(arr["CLD"]), _ptr(arr["FVC"]), C.c_int32(len(arr["SA"]))]


@pytest.mark.parametrize("name", sorted(datasets.ALL_SETS))
def test_cld_closed_form(emul, oracle, name):
    ref = datasets.ALL_SETS[name]()[0]
    arr = oracle.esa(ref).arrays()
    m = len(arr["SA"])
    out = np.zeros(m + 1, np.int64)
    emul.emul_cld(_ptr(arr["LCP"]), C.c_int32(m), _ptr(out))
    assert np.array_equal(out, arr["CLD"])


@pytest.mark.parametrize("name", ["multi_contig", "repeats", "tiny", "bang_vs_base", "divergent"])
@pytest.mark.parametrize("K", [0, 1, 3, 6, 8])
def test_table_descent(emul, oracle, name, K):
    genomes = datasets.ALL_SETS[name]()
    ref = genomes[0]
    esa = oracle.esa(ref)
    arr = esa.arrays()
    rng = np.random.default_rng(K)
    text = b"".join(genomes[1:4]) + b"ACGTAC!GT"
    pos = np.unique(np.concatenate([rng.integers(0, len(text), size=300), np.arange(len(text) - 12, len(text))]))
    lens = np.minimum(rng.integers(1, 3000, size=len(pos)), len(text) - pos).astype(np.int64)
    offs = pos.astype(np.int64)
    tb = np.frombuffer(text, np.uint8)
    for use_table in (0, 1):
        out = np.zeros(3 * len(pos), np.int64)
        emul.emul_matches(*_esa_args(arr), C.c_int32(K), _ptr(tb), _ptr(offs), _ptr(lens), C.c_int64(len(pos)), C.c_int(use_table), _ptr(out))
        for k in range(len(pos)):
            want = esa.get_match(text[offs[k] : offs[k] + lens[k]], cached=True)
            assert tuple(out[3 * k : 3 * k + 3]) == want, (k, offs[k], lens[k])


def _emul_anchor(emul, arr, K, thr, CH, CAP, q):
    cap = len(q) // max(1, thr) + 16
    out = np.zeros(5 * cap, np.int64)
    stats = np.zeros(5, np.int64)
    qb = np.frombuffer(q, np.uint8)
    emul.emul_anchor.restype = C.c_int64
    n = emul.emul_anchor(*_esa_args(arr), C.c_int32(K), C.c_int32(thr), C.c_int32(CH), C.c_int32(CAP), _ptr(qb), C.c_int32(len(q)), _ptr(out), C.c_int64(cap), _ptr(stats))
    assert n >= 0, f"emulation reported internal error {n}"
    assert n <= cap
    h = np.zeros(n, dtype=oracle_lib.HOM_DTYPE)
    o = out[: 5 * n].reshape(n, 5)
    for c, f in enumerate(("direction", "index_reference", "index_reference_projected", "index_query", "length")):
        h[f] = o[:, c]
    return h, dict(zip(("chunks", "events", "open", "merged", "unresolved"), stats.tolist()))


CONFIGS = [(32, 0, 6), (64, 64, 3), (256, 300, 0), (4096, 0, 8), (96, 200, 2)]


@pytest.mark.parametrize("name", sorted(datasets.ALL_SETS))
@pytest.mark.parametrize("CH,CAP,K", CONFIGS)
def test_speculative_walk_equals_sequential(emul, oracle, name, CH, CAP, K):
    genomes = datasets.ALL_SETS[name]()
    ref = genomes[0]
    thr = oracle.threshold(ref)
    esa = oracle.esa(ref)
    arr = esa.arrays()
    for q in genomes:
        if len(q) == 0:
            continue
        want = esa.anchor_homologies(thr, q)
        got, stats = _emul_anchor(emul, arr, K, thr, CH, CAP, q)
        assert np.array_equal(got, want), (name, len(q), stats)


def test_walk_paths_are_exercised(emul, oracle):
    """the small-chunk configurations must actually hit open matches, merges and continuations"""
    totals = {"open": 0, "merged": 0, "unresolved": 0}
    for name in ("identical_unrelated", "repeats", "rearranged", "divergent"):
        genomes = datasets.ALL_SETS[name]()
        ref = genomes[0]
        thr = oracle.threshold(ref)
        arr = oracle.esa(ref).arrays()
        for q in genomes:
            for CH, CAP, K in CONFIGS[:3]:
                _, st = _emul_anchor(emul, arr, K, thr, CH, CAP, q)
                for k in totals:
                    totals[k] += st[k]
    assert totals["open"] > 0 and totals["merged"] > 0 and totals["unresolved"] > 0, totals


def test_low_thresholds(emul, oracle):
    """thresholds far below what min_anchor_length would give make random anchors common"""
    genomes = datasets.divergent_set(seed=21, n=6000, rates=(0.05, 0.15, 0.3))
    genomes.append(datasets.random_dna(np.random.default_rng(3), 5000))
    ref = genomes[0]
    esa = oracle.esa(ref)
    arr = esa.arrays()
    for thr in (1, 2, 4, 7):
        for q in genomes:
            want = esa.anchor_homologies(thr, q)
            for CH, CAP, K in ((32, 0, 2), (64, 0, 5), (512, 0, 0)):
                got, stats = _emul_anchor(emul, arr, K, thr, CH, CAP, q)
                assert np.array_equal(got, want), (thr, CH, stats)


def _filter_via_emul(emul, homs, variant=None):
    """variant None: filter_overlaps_max as the kernels call it; 0: the cluster DP alone;
    1: the O(h log h) heap version alone"""
    h = len(homs)
    start = np.ascontiguousarray(homs["index_reference_projected"], dtype=np.int32)
    ln = np.ascontiguousarray(homs["length"], dtype=np.int32)
    keep = np.zeros(max(h, 1), np.uint8)
    if variant is None:
        emul.emul_filter(_ptr(start), _ptr(ln), C.c_int32(h), _ptr(keep))
    else:
        emul.emul_filter_variant(_ptr(start), _ptr(ln), C.c_int32(h), _ptr(keep), C.c_int32(variant))
    return homs[keep[:h] != 0]


def _homs(triples):
    h = np.zeros(len(triples), dtype=oracle_lib.HOM_DTYPE)
    for k, (r, q, l) in enumerate(triples):
        h[k] = (0, r, r, q, l)
    return h


def test_filter_known_answers(emul, oracle):
    """the four scenarios of /root/reference/test/Tprocess.cxx:54-94"""
    cases = [
        ([(0, 0, 10), (1, 1, 3)], [(0, 0, 10)]),
        ([(0, 0, 10), (10, 10, 10), (10, 10, 20), (40, 40, 5)], [(0, 0, 10), (10, 10, 20), (40, 40, 5)]),
        ([(0, 0, 10), (10, 10, 10), (10, 10, 20), (40, 40, 5), (42, 42, 2)], [(0, 0, 10), (10, 10, 20), (40, 40, 5)]),
        (
            sorted([(10, 10, 10), (0, 0, 10), (20, 20, 10), (5, 5, 10), (15, 15, 10), (25, 25, 10), (30, 30, 10)]),
            [(0, 0, 10), (10, 10, 10), (20, 20, 10), (30, 30, 10)],
        ),
    ]
    for pile, expected in cases:
        for variant in (None, 0, 1):
            got = _filter_via_emul(emul, _homs(pile), variant)
            assert np.array_equal(got, _homs(expected)), variant
        assert np.array_equal(oracle.sort_filter(_homs(pile), do_sort=False), _homs(expected))


def test_filter_random_lists(emul, oracle):
    rng = np.random.default_rng(11)
    for trial in range(300):
        h = int(rng.integers(0, 40))
        span = int(rng.integers(50, 2000))
        starts = np.sort(rng.integers(0, span, size=h))
        if trial % 3 == 0:
            starts = np.unique(starts)  # no equal starts
        lens = rng.integers(1, max(2, span // 6), size=len(starts))
        pile = _homs([(int(s), int(rng.integers(0, 10**6)), int(l)) for s, l in zip(starts, lens)])
        want = oracle.sort_filter(pile, do_sort=False)
        for variant in (None, 0, 1):
            got = _filter_via_emul(emul, pile, variant)
            assert np.array_equal(got, want), (trial, variant)


def test_filter_large_clusters_take_the_heap(emul, oracle):
    """thousands of mutually overlapping homologies (one cluster): the cluster DP gives up
    on its work budget and the O(h log h) version answers — same survivors as the reference's
    O(h^2) loop, including its tie rules (equal scores, equal starts, equal ends)"""
    rng = np.random.default_rng(12)
    for trial, (h, span, maxlen) in enumerate([(1500, 3000, 2500), (3000, 4000, 60), (2000, 500, 400), (800, 100000, 90000)]):
        starts = np.sort(rng.integers(0, span, size=h))
        lens = rng.integers(1, maxlen, size=h)
        if trial == 2:
            lens[:] = rng.integers(1, 4, size=h) * 100  # many equal scores
        pile = _homs([(int(s), int(rng.integers(0, 10**6)), int(l)) for s, l in zip(starts, lens)])
        want = oracle.sort_filter(pile, do_sort=False)
        for variant in (None, 0, 1):
            got = _filter_via_emul(emul, pile, variant)
            assert np.array_equal(got, want), (trial, variant)


def test_tile_pair_order_is_a_bijection_and_matches_the_python_mirror(emul):
    """csrc/tile_order.h: every tile pair of a launch exactly once, in the band order that
    sharding.tile_pairs spells out"""
    from phylonium_b200 import sharding

    emul.emul_unrank_pair.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    emul.emul_unrank_pair.restype = None
    for tb, te in [(0, 1), (0, 2), (0, 11), (0, 12), (0, 13), (0, 63), (5, 30), (12, 13), (23, 25), (24, 64), (40, 41), (0, 200)]:
        want = sharding.tile_pairs(tb, te)
        assert len(want) == te * (te + 1) // 2 - tb * (tb + 1) // 2
        assert len(set(want)) == len(want) and all(ti <= tj and tb <= tj < te for ti, tj in want)
        ti, tj = C.c_int32(), C.c_int32()
        got = []
        for p in range(len(want)):
            emul.emul_unrank_pair(p, tb, te, C.byref(ti), C.byref(tj))
            got.append((ti.value, tj.value))
        assert got == want, (tb, te)
