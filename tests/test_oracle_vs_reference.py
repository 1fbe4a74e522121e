"""Pins the CPU restatement (oracle/phylo_oracle.cxx) to the unmodified reference
(oracle/_ref/libphylo_ref.so, built from /root/reference by oracle/Makefile).

Skipped where the compiled reference is absent (it travels to the GPU box with the
snapshot, so normally it is present there as well)."""
import numpy as np
import pytest

import datasets
import oracle_lib

pytestmark = pytest.mark.skipif(not oracle_lib.have_reference(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def libs():
    return oracle_lib.port(), oracle_lib.reference()


def test_kinds(libs):
    p, r = libs
    assert p.kind == "port" and r.kind == "reference"


def test_sequence_helpers(libs):
    p, r = libs
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 33, 1000):
        s = datasets.random_dna(rng, n)
        if n > 3:
            s = s[: n // 2] + b"!" + s[n // 2 :]
        assert p.revcomp(s) == r.revcomp(s)
        assert p.revcomp(s) == datasets.revcomp(s)
    junk = b"tacgatc!gatc!gaa__agctagttcgcc#ccgagataNNNxyz\n>"
    assert p.filter_nucl(junk) == r.filter_nucl(junk)
    for s in (b"ACGT", b"GGGCC!AT", datasets.random_dna(rng, 999)):
        assert p.gc_content(s) == r.gc_content(s)


def test_leaf_comparators(libs):
    p, r = libs
    rng = np.random.default_rng(1)
    for n in (0, 1, 15, 16, 31, 32, 33, 63, 64, 65, 1000, 4097):
        a = datasets.mutate(rng, datasets.random_dna(rng, n), 0.0)
        b = datasets.mutate(rng, a, 0.2)
        if n > 40:
            a = a[:7] + b"!" + a[8:]
            b = b[:20] + b"!" + b[21:]
        assert p.seqcmp(a, b) == r.seqcmp(a, b)
        assert p.revseqcmp(a, b) == r.revseqcmp(a, b)
        assert p.revseqcmp(a, datasets.revcomp(a)) == r.revseqcmp(a, datasets.revcomp(a))


def test_threshold(libs):
    p, r = libs
    for gc in (0.3, 0.5, 0.65):
        for l in (11, 101, 2001, 200001, 10000001, 500000001):
            assert p.min_anchor_length(0.025, gc, l) == r.min_anchor_length(0.025, gc, l)


@pytest.mark.parametrize("name", sorted(datasets.ALL_SETS))
def test_esa_arrays_and_matches(libs, name):
    p, r = libs
    genomes = datasets.ALL_SETS[name]()
    ref = genomes[0]
    ep, er = p.esa(ref), r.esa(ref)
    ap, ar = ep.arrays(), er.arrays()
    for k in ("S", "SA", "LCP", "CLD"):
        assert np.array_equal(ap[k], ar[k]), k
    # FVC[0] reads S[SA[0]-1] in the reference (LCP[0] = -1); compared like the rest
    assert np.array_equal(ap["FVC"], ar["FVC"])
    rng = np.random.default_rng(5)
    for q in genomes[1:3]:
        if len(q) < 2:
            continue
        for pos in rng.integers(0, len(q), size=min(400, len(q))):
            sub = q[int(pos) :]
            got = ep.get_match(sub, cached=True)
            assert got == er.get_match(sub, cached=True)
            assert got == er.get_match(sub, cached=False)
            assert got == ep.get_match(sub, cached=False)


@pytest.mark.parametrize("name", sorted(datasets.ALL_SETS))
def test_homologies_and_counts(libs, name):
    p, r = libs
    genomes = datasets.ALL_SETS[name]()
    ref = genomes[0]
    thr = r.threshold(ref)
    assert thr == p.threshold(ref)
    ep, er = p.esa(ref), r.esa(ref)
    lists = []
    for q in genomes:
        hp, hr = ep.anchor_homologies(thr, q), er.anchor_homologies(thr, q)
        assert np.array_equal(hp, hr)
        fp, fr = p.sort_filter(hp), r.sort_filter(hr)
        assert np.array_equal(fp, fr)
        lists.append(fr)
    for i in range(len(genomes)):
        for j in range(i + 1, len(genomes)):
            assert p.compare(genomes[i], lists[i], genomes[j], lists[j]) == r.compare(
                genomes[i], lists[i], genomes[j], lists[j]
            )
    if all(len(l) for l in lists):
        cp, cr = p.complete_delete(lists), r.complete_delete(lists)
        for a, b in zip(cp, cr):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("name", sorted(datasets.ALL_SETS))
@pytest.mark.parametrize("flags", [0, 4])
def test_process_and_matrix(libs, name, flags):
    p, r = libs
    genomes = datasets.ALL_SETS[name]()
    if flags == 4 and name in ("tiny", "identical_unrelated"):
        pytest.skip("complete deletion dereferences empty lists in the reference")
    rp = p.process(genomes, 0, flags, threads=2)
    rr = r.process(genomes, 0, flags, threads=2)
    rt = r.process(genomes, 0, flags, threads=2, timed=True)
    for k in ("subst", "homologs"):
        assert np.array_equal(rp[k], rr[k]), k
        assert np.array_equal(rt[k], rr[k]), k
    names = [f"g{i}" for i in range(len(genomes))]
    for kind in (0, 1, 2):
        assert p.format_matrix(names, rp["subst"], rp["homologs"], kind) == r.format_matrix(
            names, rr["subst"], rr["homologs"], kind
        )


def test_simf_matches_reference_generator(libs, tmp_path):
    p, r = libs
    for seed, length, d in ((1, 1000, 0.01), (7, 5003, 0.1), (4, 70, 0.5), (11, 2000, 0.0)):
        assert p.simf(seed, seed + 1, length, d) == r.simf(seed, seed + 1, length, d)
    # and against the files the unmodified simf binary writes
    import os
    import subprocess

    simf = os.path.join(oracle_lib.ORACLE_DIR, "_ref", "simf")
    subprocess.run([simf, "-s", "3", "-l", "3000", "-d", "0.02", "-d", "0.05", "-p", str(tmp_path / "g")], check=True)
    want = p.simf_set(3, 3000, [0.02, 0.05])
    for i, w in enumerate(want):
        txt = (tmp_path / f"g{i}.fasta").read_text().split("\n", 1)[1].replace("\n", "")
        assert txt.encode() == w


def test_saturated_and_empty_pairs_print_like_libm(libs):
    """raw distance > 0.75: log() of a negative number; the reference prints what glibc hands
    back ("-nan"), an uncovered pair prints "nan", raw == 0.75 prints "inf" — port, reference
    and the host mirror (phylonium_b200.pipeline) must agree byte for byte"""
    import phylonium_b200.pipeline as pl

    p, r = libs
    names = ["a", "b", "c", "d"]
    subst = np.array([[0, 80, 75, 0], [80, 0, 10, 0], [75, 10, 0, 0], [0, 0, 0, 0]], np.uint64)
    homol = np.array([[0, 100, 100, 0], [100, 0, 100, 0], [100, 100, 0, 0], [0, 0, 0, 0]], np.uint64)
    for kind in (0, 1, 2):
        want = r.format_matrix(names, subst, homol, kind)
        assert p.format_matrix(names, subst, homol, kind) == want
        mat = [pl.EvoModel(int(subst[i, j]), int(homol[i, j])) for i in range(4) for j in range(4)]
        assert pl.format_matrix(names, mat, kind) == want
    jc = r.format_matrix(names, subst, homol, 1)
    assert "-nan" in jc and "inf" in jc


@pytest.mark.parametrize("flags", [0, 4])
def test_sampled_rows_equal_the_full_matrix(libs, flags):
    """po_process_rows (what the benchmark's sampled check uses at 1000 genomes) against the
    untouched process() call of the reference, and the port against both"""
    p, r = libs
    genomes = datasets.ALL_SETS["rearranged"]() if "rearranged" in datasets.ALL_SETS else datasets.divergent_set()
    full = r.process(genomes, 0, flags, threads=2)
    rows = [len(genomes) - 1, 0, 1]
    for lib in (p, r):
        got = lib.process_rows(genomes, rows, 0, flags, threads=2)
        for k, i in enumerate(rows):
            assert np.array_equal(got["subst"][k], full["subst"][i]), (lib.kind, i)
            assert np.array_equal(got["homologs"][k], full["homologs"][i]), (lib.kind, i)
