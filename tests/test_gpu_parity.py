"""Parity of the CUDA path with the oracle, through the C ABI (needs a B200).

Bar: bit-exact for SA, LCP, CLD, FVC, longest matches, homology lists (before and after
sort/filter) and per-pair counts; distances evaluated on the device within 1e-12
relative (BASELINE.json north_star); PHYLIP text identical."""
import numpy as np
import pytest

import datasets
import oracle_lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import phylonium_b200

    return phylonium_b200


@pytest.fixture(scope="module")
def oracle():
    return oracle_lib.best()


@pytest.fixture()
def ctx(pb):
    c = pb.Context(keep_raw=1)
    yield c
    c.close()


SETS = sorted(datasets.ALL_SETS)


@pytest.mark.parametrize("name", SETS)
def test_esa_arrays(ctx, oracle, name):
    ref = datasets.ALL_SETS[name]()[0]
    ctx.esa_build(ref)
    got = ctx.esa_arrays()
    want = oracle.esa(ref).arrays()
    for k in ("S", "SA", "LCP", "CLD", "FVC"):
        assert np.array_equal(got[k], want[k]), k


@pytest.mark.parametrize("name", SETS)
def test_esa_arrays_built_step_by_step(pb, oracle, name):
    """the careful build (a host decision after every stage) that the speculative one falls
    back to gives the same index; the speculative one is what every other test runs"""
    ref = datasets.ALL_SETS[name]()[0]
    with pb.Context(esa_speculative=0) as c:
        c.esa_build(ref)
        got = c.esa_arrays()
    want = oracle.esa(ref).arrays()
    for k in ("S", "SA", "LCP", "CLD", "FVC"):
        assert np.array_equal(got[k], want[k]), k


@pytest.mark.parametrize("n", [1, 2, 3, 31, 1023, 1024, 2047, 2048, 2049, 4095, 4096, 4097, 70001])
def test_esa_sizes_across_tile_boundaries(ctx, oracle, n):
    rng = np.random.default_rng(n)
    ref = datasets.random_dna(rng, n)
    ctx.esa_build(ref)
    got = ctx.esa_arrays()
    want = oracle.esa(ref).arrays()
    for k in ("SA", "LCP", "CLD", "FVC"):
        assert np.array_equal(got[k], want[k]), k


@pytest.mark.parametrize(
    "ref",
    [b"A" * 3000, b"AC" * 2500, b"ACGTTGCA" * 700, (b"ACGGTCA" * 50 + b"!") * 9 + b"AC"],
    ids=["polyA", "AC-repeat", "palindromic-8mer", "tandem-with-separators"],
)
def test_esa_degenerate_repeats(ctx, oracle, ref):
    """every suffix ties on its first 21 characters: all work goes through the refinement"""
    ctx.esa_build(ref)
    got = ctx.esa_arrays()
    want = oracle.esa(ref).arrays()
    for k in ("SA", "LCP", "CLD", "FVC"):
        assert np.array_equal(got[k], want[k]), k
    assert ctx.stat("esa.tied") > 0 and ctx.stat("esa.refine_rounds") >= 1


@pytest.mark.parametrize("name", ["multi_contig", "repeats", "rearranged", "tiny"])
@pytest.mark.parametrize("key_chars", [1, 2, 5, 11, 21])
def test_esa_any_sort_key_length(pb, oracle, name, key_chars):
    """short sort keys leave most suffixes tied: the refinement rounds must finish the order"""
    ref = datasets.ALL_SETS[name]()[0]
    want = oracle.esa(ref).arrays()
    with pb.Context(key_chars=key_chars) as ctx:
        ctx.esa_build(ref)
        got = ctx.esa_arrays()
        for k in ("SA", "LCP", "CLD", "FVC"):
            assert np.array_equal(got[k], want[k]), k
        assert ctx.stat("esa.key_chars") == key_chars


def _contig_torture():
    """references whose contig separators sit in repeats, poly-A runs and next to each other:
    the suffixes that start within 16 characters of a separator ("dirty" in suffix_sort.cuh)
    tie with each other and with clean suffixes in every possible way"""
    rng = np.random.default_rng(77)
    rep = datasets.random_dna(rng, 300)
    refs = {
        "seps_in_polyA": b"A" * 40 + b"!" + b"A" * 23 + b"!" + b"A" * 7 + b"!" + b"AAAC" + b"A" * 60,
        "same_contig_many_times": b"!".join([rep] * 12),
        "contigs_end_in_repeat": b"!".join(datasets.random_dna(rng, 50 + 7 * k) + rep for k in range(20)),
        "contigs_start_with_repeat": b"!".join(rep[: 20 + k] + datasets.random_dna(rng, 40) for k in range(30)),
        "short_contigs": b"!".join(datasets.random_dna(rng, 1 + (k % 19)) for k in range(400)),
        "adjacent_separators": b"ACGT!!ACGT!!!A!C!G!T!!" + datasets.random_dna(rng, 100) + b"!",
        "separator_first_and_last": b"!" + datasets.random_dna(rng, 500) + b"!",
        "all_T_then_seps": b"T" * 100 + b"!" + b"T" * 100 + b"!" + b"T" * 15 + b"!" + b"T" * 16 + b"!" + b"T" * 17,
        "many_contigs": b"!".join(datasets.random_dna(rng, 30 + (k % 11)) for k in range(900)),
    }
    return refs


@pytest.mark.parametrize("name", sorted(_contig_torture()))
@pytest.mark.parametrize("key_chars", [0, 3, 8, 16])
def test_esa_separators_packed_sorter(pb, oracle, name, key_chars):
    ref = _contig_torture()[name]
    want = oracle.esa(ref).arrays()
    with pb.Context(key_chars=key_chars) as ctx:
        ctx.esa_build(ref)
        got = ctx.esa_arrays()
        assert ctx.stat("esa.packed") == 1 and ctx.stat("esa.dirty") >= 3
        for k in ("SA", "LCP", "CLD", "FVC"):
            assert np.array_equal(got[k], want[k]), k


def test_esa_too_many_separators_take_the_general_sorter(pb, oracle):
    """more suffixes next to separators than the packed sorter lists (32768): the speculative
    build notices at its end and the index is built again with the general sorter; just under
    the old estimate (1100 contigs) the packed sorter now copes, its list is filled by what is
    really there"""
    rng = np.random.default_rng(5)
    for contigs, packed in ((1100, None), (3000, 0)):
        ref = b"!".join(datasets.random_dna(rng, 20) for _ in range(contigs))
        want = oracle.esa(ref).arrays()
        for spec in (1, 0):
            with pb.Context(esa_speculative=spec) as ctx:
                ctx.esa_build(ref)
                got = ctx.esa_arrays()
                if packed is not None:
                    assert ctx.stat("esa.packed") == packed
                for k in ("SA", "LCP", "CLD", "FVC"):
                    assert np.array_equal(got[k], want[k]), (contigs, spec, k)


@pytest.mark.parametrize("name", SETS)
def test_esa_general_sorter(pb, oracle, name):
    """sort_path 1: 3-bit codes, 64-bit keys + 32-bit indices (what long keys and
    separator-rich references use)"""
    ref = datasets.ALL_SETS[name]()[0]
    want = oracle.esa(ref).arrays()
    try:
        with pb.Context(sort_path=1) as ctx:
            ctx.esa_build(ref)
            got = ctx.esa_arrays()
            assert ctx.stat("esa.packed") == 0
            for k in ("SA", "LCP", "CLD", "FVC"):
                assert np.array_equal(got[k], want[k]), k
    finally:
        with pb.Context(sort_path=0):
            pass


def test_esa_packed_equals_general_on_a_large_text(pb):
    """1 Mbp with separators: both sorters must give the same index (no oracle needed)"""
    rng = np.random.default_rng(11)
    parts = [datasets.random_dna(rng, 100_000) for _ in range(9)]
    parts.append(parts[2][:5000])  # a repeat next to a separator
    parts.append(b"A" * 1000)
    ref = b"!".join(parts)
    res = []
    try:
        for path in (0, 1):
            with pb.Context(sort_path=path) as ctx:
                ctx.esa_build(ref)
                assert ctx.stat("esa.packed") == 1 - path
                res.append(ctx.esa_arrays())
    finally:
        with pb.Context(sort_path=0):
            pass
    for k in ("SA", "LCP", "CLD", "FVC"):
        assert np.array_equal(res[0][k], res[1][k]), k


@pytest.mark.parametrize("key_chars", [1, 2, 5, 11, 21])
@pytest.mark.parametrize("name", ["multi_contig", "repeats", "rearranged", "tiny"])
def test_esa_general_sorter_any_key_length(pb, oracle, name, key_chars):
    """the general sorter with short keys: nearly everything goes through the refinement,
    including the first suffixes of the array (found by tests/fuzz_gpu.py: FVC[0] follows the
    final SA[0])"""
    ref = datasets.ALL_SETS[name]()[0]
    want = oracle.esa(ref).arrays()
    try:
        with pb.Context(key_chars=key_chars, sort_path=1) as ctx:
            ctx.esa_build(ref)
            got = ctx.esa_arrays()
            assert ctx.stat("esa.packed") == 0
            for k in ("SA", "LCP", "CLD", "FVC"):
                assert np.array_equal(got[k], want[k]), k
    finally:
        with pb.Context(sort_path=0):
            pass


def test_esa_first_suffix_in_a_tie_group(pb, oracle):
    """several contigs that start alike: the smallest suffixes ('!...') tie on their sort key"""
    ref = b"!".join([b"ACGTACGTACGTACGTACGTACGTAAAC" + (b"A" if k % 2 else b"C") * k for k in range(1, 12)])
    want = oracle.esa(ref).arrays()
    try:
        for path in (0, 1):
            for kc in (0, 2, 21):
                with pb.Context(key_chars=kc, sort_path=path) as ctx:
                    ctx.esa_build(ref)
                    got = ctx.esa_arrays()
                    for k in ("SA", "LCP", "CLD", "FVC"):
                        assert np.array_equal(got[k], want[k]), (path, kc, k)
    finally:
        with pb.Context(sort_path=0):
            pass


@pytest.mark.parametrize("mode", [1, 2])
def test_both_radix_sort_schemes(pb, oracle, mode):
    """the look-back ("onesweep") passes are normally used only for very large inputs"""
    try:
        with pb.Context(sort_mode=mode, keep_raw=1) as ctx:
            for name in ("repeats", "rearranged"):
                genomes = datasets.ALL_SETS[name]()
                ref = genomes[0]
                ctx.esa_build(ref)
                got, want = ctx.esa_arrays(), oracle.esa(ref).arrays()
                for k in ("SA", "LCP", "CLD", "FVC"):
                    assert np.array_equal(got[k], want[k]), (name, k)
            genomes = oracle_lib.port().simf_set(9, 300000, [0.03])
            want = oracle.process(genomes, 0, 0, threads=2)
            subst, homol = ctx.process(genomes, 0, 0)
            assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])
    finally:
        with pb.Context(sort_mode=0):
            pass


@pytest.mark.parametrize("name", ["multi_contig", "repeats", "tiny", "bang_vs_base", "divergent"])
@pytest.mark.parametrize("K", [-1, 0, 1, 4, 9])
def test_longest_matches(pb, oracle, name, K):
    genomes = datasets.ALL_SETS[name]()
    ref = genomes[0]
    esa = oracle.esa(ref)
    with pb.Context(kmer_k=K) as ctx:
        ctx.esa_build(ref)
        rng = np.random.default_rng(7)
        text = b"".join(genomes[1:4]) + b"ACGTAC!GT"
        pos = np.unique(np.concatenate([rng.integers(0, len(text), size=400), np.arange(max(0, len(text) - 14), len(text))]))
        lens = np.minimum(rng.integers(1, 4000, size=len(pos)), len(text) - pos)
        for use_table in (True, False):
            got = ctx.get_matches(text, pos, lens, use_table)
            for k in range(len(pos)):
                want = esa.get_match(text[pos[k] : pos[k] + lens[k]], cached=True)
                assert tuple(got[k]) == want, (k, pos[k], lens[k])


@pytest.mark.parametrize("K", [1, 5, 6, 7, 8, 11])
def test_table_built_level_by_level(pb, K):
    """the K-mer table grown one level at a time (default) must serve exactly the same
    longest matches as the table made entry by entry from the root (table_direct), on a text
    with repeats and separators where many entries sit inside long edge labels"""
    rng = np.random.default_rng(K)
    parts = [datasets.random_dna(rng, 150_000) for _ in range(3)]
    parts += [parts[0][1000:9000], b"ACG" * 700, b"A" * 500]
    ref = b"!".join(parts)
    text = datasets.mutate(rng, ref.replace(b"!", b"A"), 0.03)
    pos = np.arange(0, len(text) - 40, 7)
    lens = np.minimum(2000, len(text) - pos)
    res = []
    try:
        for direct in (2, 1):
            with pb.Context(kmer_k=K, table_direct=direct) as ctx:
                ctx.esa_build(ref)
                res.append(ctx.get_matches(text, pos, lens, True))
        with pb.Context(kmer_k=0) as ctx:
            ctx.esa_build(ref)
            res.append(ctx.get_matches(text, pos, lens, False))
    finally:
        with pb.Context(table_direct=0):
            pass
    # a match of length 0 reports the root interval from the root, but whatever the table
    # holds otherwise; the consumers only use l, i == j and SA[i]
    assert np.array_equal(res[0], res[1])
    assert np.array_equal(res[0][:, 0], res[2][:, 0])
    same = res[2][:, 0] > 0
    assert np.array_equal(res[0][same], res[2][same])


@pytest.mark.parametrize("name", SETS)
@pytest.mark.parametrize("chunk,cap", [(32, 0), (64, 64), (256, 300), (4096, 0)])
def test_homologies(pb, oracle, name, chunk, cap):
    genomes = [g for g in datasets.ALL_SETS[name]() if len(g)]
    ref = genomes[0]
    thr = oracle.threshold(ref)
    assert thr == pb.threshold_for(ref)
    esa = oracle.esa(ref)
    with pb.Context(keep_raw=1, chunk=chunk, cap=cap) as ctx:
        ctx.esa_build(ref)
        ctx.map_queries(genomes, thr)
        for k, q in enumerate(genomes):
            raw = esa.anchor_homologies(thr, q)
            assert np.array_equal(ctx.homologies(k, raw=True), raw), (k, "raw")
            assert np.array_equal(ctx.homologies(k), oracle.sort_filter(raw)), (k, "filtered")


@pytest.mark.parametrize("thr", [1, 3, 6])
def test_homologies_low_threshold(pb, oracle, thr):
    genomes = datasets.divergent_set(seed=21, n=6000, rates=(0.05, 0.15, 0.3))
    ref = genomes[0]
    esa = oracle.esa(ref)
    with pb.Context(keep_raw=1, chunk=64) as ctx:
        ctx.esa_build(ref)
        ctx.map_queries(genomes, thr)
        for k, q in enumerate(genomes):
            raw = esa.anchor_homologies(thr, q)
            assert np.array_equal(ctx.homologies(k, raw=True), raw)
            assert np.array_equal(ctx.homologies(k), oracle.sort_filter(raw))


@pytest.mark.parametrize("name", SETS)
@pytest.mark.parametrize("flags", [0, 4])
def test_pair_counts_and_matrix(pb, oracle, ctx, name, flags):
    genomes = [g for g in datasets.ALL_SETS[name]() if len(g)]
    if flags == 4 and name in ("tiny", "identical_unrelated"):
        pytest.skip("complete deletion dereferences empty lists in the reference")
    want = oracle.process(genomes, 0, flags, threads=4)
    subst, homol = ctx.process(genomes, 0, flags)
    assert np.array_equal(homol, want["homologs"])
    assert np.array_equal(subst, want["subst"])
    names = [f"g{i}" for i in range(len(genomes))]
    mat = [pb.EvoModel(int(subst[i, j]), int(homol[i, j])) for i in range(len(genomes)) for j in range(len(genomes))]
    for kind in (0, 1, 2):
        assert pb.format_matrix(names, mat, kind) == oracle.format_matrix(names, want["subst"], want["homologs"], kind)
        dev = ctx.estimate(kind)
        for i in range(len(genomes)):
            for j in range(len(genomes)):
                ref_val = 0.0 if i == j else oracle.estimate(want["subst"][i, j], want["homologs"][i, j], kind)
                if np.isnan(ref_val):
                    assert np.isnan(dev[i, j])
                else:
                    assert abs(dev[i, j] - ref_val) <= 1e-12 * max(1.0, abs(ref_val))  # tolerance of the north star


def test_equal_starts_take_the_host_sort(pb, oracle):
    """two homologies of one query with the same reference start: the reference's result
    hangs on libstdc++'s unstable std::sort, so the library runs that very call"""
    rng = np.random.default_rng(41)
    r = datasets.random_dna(rng, 20000)
    dup = r[4000:9000]

    def other(b):  # a base different from b, so that the copies start and end exactly at the duplicate
        return b"C" if b != ord("C") else b"G"

    # first copy at query position 0; the second one right after a full match of another
    # reference stretch plus one mismatching base, so that the walk lands on its first base
    post, stretch, bad = other(r[9000]), r[12000:13000], other(r[13000])
    q = dup + post + stretch + bad + dup + post + datasets.random_dna(rng, 300)
    assert len(np.unique(esa_starts := oracle.esa(r).anchor_homologies(oracle.threshold(r), q)["index_reference_projected"])) < len(esa_starts)
    genomes = [r, q, datasets.mutate(rng, r, 0.01)]
    thr = oracle.threshold(r)
    esa = oracle.esa(r)
    with pb.Context(keep_raw=1) as ctx:
        ctx.esa_build(r)
        ctx.map_queries(genomes, thr)
        assert ctx.stat("anchor.tie_fallback") == 1 and ctx.stat("anchor.general_path") == 1
        for k, q in enumerate(genomes):
            raw = esa.anchor_homologies(thr, q)
            assert np.array_equal(ctx.homologies(k, raw=True), raw)
            assert np.array_equal(ctx.homologies(k), oracle.sort_filter(raw))
        subst, homol = ctx.compare_all()
    want = oracle.process(genomes, 0, 0, threads=2)
    assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])


@pytest.mark.parametrize("name", sorted(datasets.ALL_SETS))
def test_process_with_the_mapping_as_one_graph(pb, oracle, name):
    """map_graph = 2: validation, walk, path, lists, rows and comparison of every batch are
    captured and submitted as one CUDA graph whatever the size; twice on one context, so that the
    second call updates the instantiated graph of the first"""
    genomes = datasets.ALL_SETS[name]()
    want = oracle.process(genomes, 0, 0, threads=2)
    with pb.Context(map_graph=2, map_batch_bytes=20000) as ctx:
        for _ in range(2):
            subst, homol = ctx.process(genomes, 0, 0)
            assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])
        again = ctx.process_again(len(genomes) - 1, 4)
    want2 = oracle.process(genomes, len(genomes) - 1, 4, threads=2)
    assert np.array_equal(again[0], want2["subst"]) and np.array_equal(again[1], want2["homologs"])


def _general_path_genomes():
    """genome 2 and 3 carry more homologies than the per-query sort holds"""
    rng = np.random.default_rng(43)
    blocks = [datasets.random_dna(rng, 150) for _ in range(2600)]
    r = b"".join(blocks)
    order = rng.permutation(len(blocks))
    q1 = b"".join(blocks[i] for i in order)
    q2 = b"".join(datasets.revcomp(blocks[i]) if i % 3 == 0 else blocks[i] for i in order[::-1])
    return [r, datasets.mutate(rng, r, 0.02), q1, q2, datasets.mutate(rng, r, 0.01), datasets.mutate(rng, r, 0.03)]


@pytest.mark.parametrize("map_graph", [1, 2])
@pytest.mark.parametrize("batch_bytes", [0, 300_000, 700_000])
@pytest.mark.parametrize("flags", [0, 4])
def test_process_redoes_a_batch_whose_lists_were_not_final(pb, oracle, batch_bytes, flags, map_graph):
    """phylo_process queues the rows (and the comparison) of a batch before the host has seen
    whether its filtered lists are final; a batch that needs the global sort after all is
    done again — in the first, a middle or the only batch"""
    genomes = _general_path_genomes()
    with pb.Context(map_graph=map_graph) as ctx:  # 2: also for batches this small, the mapping as one CUDA graph
        if batch_bytes:
            ctx.set_option("map_batch_bytes", batch_bytes)
        subst, homol = ctx.process(genomes, 0, flags)
        assert ctx.stat("anchor.general_path") >= 1
        assert ctx.stat("map.batches") == {0: 1, 300_000: 6, 700_000: 3}[batch_bytes]  # 390 kbp per genome
        again = ctx.process_again(0, flags)
    want = oracle.process(genomes, 0, flags, threads=2)
    assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])
    assert np.array_equal(again[0], want["subst"]) and np.array_equal(again[1], want["homologs"])


@pytest.mark.parametrize("map_graph", [1, 2])
@pytest.mark.parametrize("kind", ["contigs", "repeats"])
def test_process_on_an_index_the_speculative_build_gives_up_on(pb, oracle, kind, map_graph):
    """the mapping is queued behind the index build; when the build's assumptions fail (too
    many contigs for the packed sorter's list, repeats that need the doubling rounds) its kernels
    do nothing, the index is built again step by step and the batch is mapped again"""
    rng = np.random.default_rng(47)
    if kind == "contigs":
        r = b"!".join(datasets.random_dna(rng, 40 + (k % 7)) for k in range(3000))
    else:
        unit = datasets.random_dna(rng, 700)
        r = unit * 6 + datasets.random_dna(rng, 3000) + unit * 3
    genomes = [r, datasets.mutate(rng, r, 0.02), datasets.mutate(rng, r, 0.05), datasets.random_dna(rng, 5000)]
    with pb.Context(map_graph=map_graph) as ctx:
        subst, homol = ctx.process(genomes, 0, 0)
        if kind == "contigs":
            assert ctx.stat("esa.packed") == 0
        else:
            assert ctx.stat("esa.refine_rounds") >= 1
        # ... and through the stage calls: the build returns early, the mapping finds out
        ctx.esa_build(r)
        ctx.map_queries(genomes, oracle.threshold(r))
        s2, h2 = ctx.compare_all()
    want = oracle.process(genomes, 0, 0, threads=2)
    assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])
    assert np.array_equal(s2, want["subst"]) and np.array_equal(h2, want["homologs"])


def test_thousands_of_homologies_take_the_global_sort(pb, oracle):
    """more homologies in one list than the per-query shared-memory sort holds"""
    rng = np.random.default_rng(43)
    blocks = [datasets.random_dna(rng, 150) for _ in range(2600)]
    r = b"".join(blocks)
    order = rng.permutation(len(blocks))
    q1 = b"".join(blocks[i] for i in order)
    q2 = b"".join(datasets.revcomp(blocks[i]) if i % 3 == 0 else blocks[i] for i in order[::-1])
    genomes = [r, q1, q2, datasets.mutate(rng, r, 0.02)]
    thr = oracle.threshold(r)
    esa = oracle.esa(r)
    with pb.Context(keep_raw=1) as ctx:
        ctx.esa_build(r)
        ctx.map_queries(genomes, thr)
        assert ctx.homology_counts(raw=True).max() > 2048
        assert ctx.stat("anchor.general_path") == 1 and ctx.stat("anchor.tie_fallback") == 0
        for k, q in enumerate(genomes):
            raw = esa.anchor_homologies(thr, q)
            assert np.array_equal(ctx.homologies(k, raw=True), raw)
            assert np.array_equal(ctx.homologies(k), oracle.sort_filter(raw))
        subst, homol = ctx.compare_all()
    want = oracle.process(genomes, 0, 0, threads=2)
    assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])


def test_reference_need_not_be_first(pb, oracle, ctx):
    genomes = datasets.multi_contig_set()
    for ref_index in (2, 5):
        want = oracle.process(genomes, ref_index, 0, threads=4)
        subst, homol = ctx.process(genomes, ref_index, 0)
        assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])


@pytest.mark.parametrize("batch_bytes", [1, 9000, 30000])
def test_mapping_in_batches(pb, oracle, batch_bytes):
    """phylo_process / phylo_map_queries map the sequences batch by batch (512 MiB by default);
    tiny batches here: every count, the per-sequence homology lists and the raw lists must not
    depend on where the batch borders fall"""
    rng = np.random.default_rng(31)
    r = datasets.random_dna(rng, 6000)
    genomes = [r] + [datasets.mutate(rng, r, 0.003 * (k + 1)) for k in range(10)]
    genomes[4] = datasets.revcomp(genomes[4])
    genomes[7] = genomes[7][:2500] + b"!" + genomes[7][2500:]
    genomes[9] = b""
    want = oracle.process(genomes, 0, 0, threads=4)
    try:
        with pb.Context(keep_raw=1, map_batch_bytes=batch_bytes) as ctx:
            subst, homol = ctx.process(genomes, 0, 0)
            assert ctx.stat("map.batches") > 1
            assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])
            with pb.Context(keep_raw=1, map_batch_bytes=1 << 40) as one:
                one.process(genomes, 0, 0)
                assert one.stat("map.batches") == 1
                assert np.array_equal(ctx.homology_counts(), one.homology_counts())
                assert np.array_equal(ctx.homology_counts(raw=True), one.homology_counts(raw=True))
                for k in range(len(genomes)):
                    assert np.array_equal(ctx.homologies(k), one.homologies(k)), k
                    assert np.array_equal(ctx.homologies(k, raw=True), one.homologies(k, raw=True)), k
    finally:
        with pb.Context(map_batch_bytes=512 << 20):
            pass


def _kitchen_sink(n=600_000, seed=41):
    """a 5-contig reference with a 20 kbp repeat and low-complexity stretches; genomes that are
    mutated, cut, inverted, shuffled, reverse complemented, truncated or unrelated"""
    rng = np.random.default_rng(seed)
    contigs = [datasets.random_dna(rng, n // 5) for _ in range(5)]
    repeat = contigs[0][5000:25000]
    contigs[2] = contigs[2][:40000] + repeat + contigs[2][40000:]
    contigs[3] = contigs[3][:1000] + b"A" * 300 + b"ACAC" * 100 + contigs[3][1000:]
    ref = b"!".join(contigs)
    flat = ref.replace(b"!", b"")
    genomes = [ref]
    for d in (0.001, 0.01, 0.03, 0.06):
        genomes.append(datasets.mutate(rng, flat, d))
    genomes.append(datasets.indel(rng, datasets.mutate(rng, flat, 0.01), 60, 200))
    blocks = [flat[i : i + 30000] for i in range(0, len(flat), 30000)]
    order = rng.permutation(len(blocks))
    genomes.append(b"".join(datasets.revcomp(blocks[k]) if k % 3 == 0 else blocks[k] for k in order))
    genomes.append(datasets.revcomp(datasets.mutate(rng, flat, 0.02)))
    genomes.append(b"!".join(datasets.mutate(rng, c, 0.005) for c in reversed(contigs)))
    genomes.append(datasets.mutate(rng, flat, 0.02)[: n // 3])
    genomes.append(datasets.random_dna(rng, 50000))
    genomes.append(ref)
    return genomes


def test_kitchen_sink_600kbp(pb, oracle):
    """everything at once at a size where chunks, batches and tiles are many: counts, with and
    without complete deletion, and every homology list"""
    genomes = _kitchen_sink()
    try:
        with pb.Context(keep_raw=1, map_batch_bytes=2_000_000) as ctx:
            for flags in (0, 4):
                want = oracle.process(genomes, 0, flags, threads=8)
                subst, homol = ctx.process(genomes, 0, flags)
                assert np.array_equal(homol, want["homologs"]) and np.array_equal(subst, want["subst"]), flags
            assert ctx.stat("map.batches") > 1
            thr = oracle.threshold(genomes[0])
            esa = oracle.esa(genomes[0])
            for k, q in enumerate(genomes):
                raw = esa.anchor_homologies(thr, q)
                assert np.array_equal(ctx.homologies(k, raw=True), raw), (k, "raw")
                assert np.array_equal(ctx.homologies(k), oracle.sort_filter(raw)), (k, "filtered")
    finally:
        with pb.Context(map_batch_bytes=512 << 20):
            pass


def test_many_genomes_tiles(pb, oracle, ctx):
    """more genomes than one 4x4 tile, odd count: diagonal and ragged tiles"""
    rng = np.random.default_rng(17)
    r = datasets.random_dna(rng, 4000)
    genomes = [r] + [datasets.mutate(rng, r, 0.002 * (k + 1)) for k in range(12)]
    genomes[5] = datasets.revcomp(genomes[5])
    genomes[9] = genomes[9][:2000]
    want = oracle.process(genomes, 0, 0, threads=4)
    subst, homol = ctx.process(genomes, 0, 0)
    assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])
    want = oracle.process(genomes, 0, 4, threads=4)
    subst, homol = ctx.process(genomes, 0, 4)
    assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])


def test_config4_shape_scaled(pb, oracle, ctx):
    """BASELINE.json configs[3] in small: many genomes (67 x 200 kbp, d up to 0.05), all 2211
    pairs; also with complete deletion"""
    n, N = 200000, 67
    dists = [0.05 * i / (N - 1) for i in range(1, N)]
    genomes = oracle_lib.port().simf_set(4, n, dists)
    for flags in (0, 4):
        want = oracle.process(genomes, 0, flags, threads=16)
        subst, homol = ctx.process(genomes, 0, flags)
        assert np.array_equal(homol, want["homologs"]) and np.array_equal(subst, want["subst"])


def test_config5_shape_scaled(pb, oracle, ctx):
    """BASELINE.json configs[4] in small: multi-contig genomes, every third contig reverse
    complemented, cut points shifted per genome"""
    L, N, ncontig = 600000, 6, 12
    base = oracle_lib.port().simf_set(5, L, [0.02 * i / (N - 1) for i in range(1, N)])
    genomes = []
    for g, seq in enumerate(base):
        step = L // ncontig
        cuts = [c * step + (1000 * g if g % 2 else 0) for c in range(1, ncontig)]
        pieces, prev = [], 0
        for c, cut in enumerate(cuts + [L]):
            piece = seq[prev:cut]
            pieces.append(datasets.revcomp(piece) if (c + g) % 3 == 0 else piece)
            prev = cut
        genomes.append(b"!".join(pieces))
    want = oracle.process(genomes, 0, 0, threads=16)
    subst, homol = ctx.process(genomes, 0, 0)
    assert np.array_equal(homol, want["homologs"]) and np.array_equal(subst, want["subst"])
    assert (homol[0, 1:] > 0.7 * L).all()  # reversed contigs are found too


def test_config1_simf(pb, oracle, ctx):
    """BASELINE.json configs[0]: 2 x 100 kbp, d = 0.01, the reference's own CPU-runnable case"""
    genomes = oracle_lib.port().simf_set(1, 100000, [0.01])
    want = oracle.process(genomes, 0, 0, threads=2)
    subst, homol = ctx.process(genomes, 0, 0)
    assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])
    mat = [pb.EvoModel(int(subst[i, j]), int(homol[i, j])) for i in range(2) for j in range(2)]
    assert pb.format_matrix(["g0", "g1"], mat) == "2\ng0  0.0000e+00  9.9410e-03\ng1  9.9410e-03  0.0000e+00\n"


def test_index_arrays_1mbp(pb, oracle, ctx):
    genomes = oracle_lib.port().simf_set(5, 1000000, [0.02])
    ctx.esa_build(genomes[0])
    got = ctx.esa_arrays()
    want = oracle.esa(genomes[0]).arrays()
    for k in ("SA", "LCP", "CLD", "FVC"):
        assert np.array_equal(got[k], want[k]), k


@pytest.mark.slow
def test_config2_full_size(pb, oracle, ctx):
    """BASELINE.json configs[1]: 8 x 5 Mbp (simf -s 2, d up to 0.05), full size, every count"""
    dists = [0.001, 0.002, 0.005, 0.01, 0.02, 0.03, 0.05]
    genomes = oracle_lib.port().simf_set(2, 5000000, dists)
    want = oracle.process(genomes, 0, 0, threads=16, timed=True)
    subst, homol = ctx.process(genomes, 0, 0)
    assert np.array_equal(homol, want["homologs"])
    assert np.array_equal(subst, want["subst"])
    assert np.array_equal(ctx.homology_counts(), want["hom_counts"].astype(np.uint64))


def _broadcast_stand_in(ctxs, ref):
    """index built on context 0, copied into the others (stands in for the NCCL broadcast)"""
    import torch

    from phylonium_b200 import sharding

    ctxs[0].esa_build(ref)
    a0 = ctxs[0].esa_device_arrays()
    for c in ctxs[1:]:
        c.esa_alloc(len(ref))
        a1 = c.esa_device_arrays()
        for name in a0:
            src = sharding.DeviceBuffer(a0[name][0], a0[name][1], 0).tensor()
            dst = sharding.DeviceBuffer(a1[name][0], a1[name][1], 0).tensor()
            dst.copy_(src)
        torch.cuda.synchronize()
        c.esa_finish_import()
        assert c.stat("esa.gc_count") == ctxs[0].stat("esa.gc_count")


def _sharded_family(seed=23, n=30000, count=7):
    rng = np.random.default_rng(seed)
    r = datasets.random_dna(rng, n)
    genomes = [r] + [datasets.mutate(rng, r, 0.004 * (k + 1)) for k in range(count - 1)]
    genomes[3] = datasets.revcomp(genomes[3])
    return genomes


@pytest.mark.parametrize("flags", [0, 4])
def test_sharded_plumbing_on_one_gpu(pb, oracle, flags):
    """the multi-GPU C ABI (index export/import, row store slices, matrix tiles) with two
    contexts on one device and plain copies in place of the NCCL collectives; 7 genomes on 2
    ranks leave a padding slot, which complete deletion (flags = 4) must not count as a genome"""
    import torch

    from phylonium_b200 import sharding

    genomes = _sharded_family()
    total, world = len(genomes), 2
    want = oracle.process(genomes, 0, flags, threads=4)
    thr = pb.threshold_for(genomes[0])
    ctxs = [pb.Context(), pb.Context()]
    try:
        plans = [sharding.make_plan(total, world, k) for k in range(world)]
        _broadcast_stand_in(ctxs, genomes[0])
        stores = []
        for k in range(world):
            p = plans[k]
            ctxs[k].rows_configure(p.padded_total, p.first)
            ctxs[k].map_queries(genomes[p.first : p.first + p.count], thr)
            ptr, bpg, tot = ctxs[k].rows_device()
            assert tot == p.padded_total
            stores.append((sharding.DeviceBuffer(ptr, bpg * tot, 0).tensor(), bpg))
        for k in range(world):  # stands in for the all-gather
            p, (st, bpg) = plans[k], stores[k]
            lo, hi = p.first * bpg, (p.first + p.per_rank) * bpg
            stores[1 - k][0][lo:hi].copy_(st[lo:hi])
        torch.cuda.synchronize()
        n = plans[0].padded_total
        acc = torch.zeros(2, n * n, dtype=torch.int64, device="cuda")
        for k in range(world):  # stands in for the all-reduce
            part = torch.zeros(2, n * n, dtype=torch.int64, device="cuda")
            torch.cuda.synchronize()  # torch's stream and the context's are not ordered
            ctxs[k].compare_tiles_dev(part[0].data_ptr(), part[1].data_ptr(), k, world, flags)
            torch.cuda.synchronize()  # the call only queues the work on the context's stream
            acc += part
        subst = acc[0].cpu().numpy().reshape(n, n)[:total, :total].astype(np.uint64)
        homol = acc[1].cpu().numpy().reshape(n, n)[:total, :total].astype(np.uint64)
        assert homol.any()
        assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])
        assert not acc[1].cpu().numpy().reshape(n, n)[total:, :].any()  # padding rows stay empty
        # a rank's share of the tiles is not a matrix to estimate distances from
        with pytest.raises(pb.PhyloError):
            ctxs[0].estimate(pb.DIST_JC, total=n)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("map_graph", [1, 2])
@pytest.mark.parametrize("flags", [0, 4])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_push_exchange_on_one_gpu(pb, oracle, world, flags, map_graph):
    """rows pushed into the peers' stores while the next batch is mapped (phylo_rows_set_peers,
    the in-process form of the IPC exchange), genomes dealt round-robin to the ranks, the
    slot-ordered matrix brought back into genome order: counts equal to the oracle's"""
    import torch

    from phylonium_b200 import sharding

    genomes = _sharded_family(count=11)
    total = len(genomes)
    want = oracle.process(genomes, 0, flags, threads=4)
    thr = pb.threshold_for(genomes[0])
    # several batches per rank; map_graph = 2: the last batch of each (mapping, rows, the kernel
    # that pushes them to the peers) goes out as one CUDA graph
    ctxs = [pb.Context(map_batch_bytes=50000, map_graph=map_graph) for _ in range(world)]
    try:
        plans = [sharding.make_plan(total, world, k, "interleaved") for k in range(world)]
        _broadcast_stand_in(ctxs, genomes[0])
        for k in range(world):
            ctxs[k].rows_configure(plans[k].padded_total, plans[k].first)
        ptrs = [c.rows_device()[0] for c in ctxs]
        for k in range(world):
            ctxs[k].rows_set_peers(ptrs, k)
        for k in range(world):
            ctxs[k].map_queries([genomes[g] for g in plans[k].genomes()], thr)
            assert ctxs[k].stat("map.batches") > 1
        torch.cuda.synchronize()  # stands in for the barrier across ranks
        n = plans[0].padded_total
        acc = torch.zeros(2, n * n, dtype=torch.int64, device="cuda")
        for k in range(world):
            part = torch.zeros(2, n * n, dtype=torch.int64, device="cuda")
            torch.cuda.synchronize()  # torch's stream and the context's are not ordered
            ctxs[k].compare_tiles_dev(part[0].data_ptr(), part[1].data_ptr(), k, world, flags)
            torch.cuda.synchronize()  # the call only queues the work on the context's stream
            acc += part
        got = sharding.genome_order(acc, plans[0]).cpu().numpy().astype(np.uint64)
        assert np.array_equal(got[0], want["subst"]) and np.array_equal(got[1], want["homologs"])
        # switching the push off again: the context keeps its rows to itself
        ctxs[0].rows_set_peers(None, 0)
    finally:
        for c in ctxs:
            c.close()


def _ipc_worker(rank, world, port, genomes, thr, flags, results):
    import os

    import torch
    import torch.distributed as dist

    import phylonium_b200 as pb
    from phylonium_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        total = len(genomes)
        plan = sharding.make_plan(total, world, rank, "interleaved")
        with pb.Context(0, map_batch_bytes=60000) as ctx:
            ctx.esa_build(genomes[0])  # index replicated
            ctx.rows_configure(plan.padded_total, plan.first)
            sharding.setup_push(ctx, rank, world)
            ctx.map_queries([genomes[g] for g in plan.genomes()], thr)
            torch.cuda.synchronize()
            dist.barrier()  # every rank's pushes have landed
            n = plan.padded_total
            part = torch.zeros(2, n * n, dtype=torch.int64, device="cuda")
            torch.cuda.synchronize()  # torch's stream and the context's are not ordered
            ctx.compare_tiles_dev(part[0].data_ptr(), part[1].data_ptr(), rank, world, flags)
            torch.cuda.synchronize()  # the call only queues the work on the context's stream
            part = part.cpu()
            dist.all_reduce(part)
            got = sharding.genome_order(part, plan).numpy().astype(np.uint64)
            dist.barrier()  # nobody closes its row store while a peer still has it mapped
            results[rank] = (got[0], got[1])
    finally:
        dist.destroy_process_group()


def test_row_exchange_across_processes(pb, oracle):
    """two processes (ranks) on this one GPU: CUDA IPC handles of the row stores are exchanged
    (sharding.setup_push) and every rank pushes its rows into the other's store"""
    import socket

    import torch.multiprocessing as mp

    genomes = _sharded_family(seed=5, count=9)
    want = oracle.process(genomes, 0, 0, threads=4)
    thr = pb.threshold_for(genomes[0])
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_ipc_worker, args=(2, port, genomes, thr, 0, results), nprocs=2, join=True)
        got = dict(results)
    for r in (0, 1):
        assert np.array_equal(got[r][0], want["subst"]) and np.array_equal(got[r][1], want["homologs"])


def test_second_pass_keeps_sequences_on_the_device(pb, oracle, ctx):
    """--2pass (src/phylonium.cxx:289-296): process() again with another reference; the
    sequences stay where the first pass put them"""
    genomes = _sharded_family(seed=9, count=6)
    first = ctx.process(genomes, 0, 0)
    want0 = oracle.process(genomes, 0, 0, threads=4)
    assert np.array_equal(first[0], want0["subst"]) and np.array_equal(first[1], want0["homologs"])
    for ref in (4, 3, 0):  # genome 3 is a reverse complement
        want = oracle.process(genomes, ref, 0, threads=4)
        subst, homol = ctx.process_again(ref)
        assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"]), ref
        assert ctx.stat("process.h2d_bytes") == 0
    with pytest.raises(pb.PhyloError):
        ctx.process_again(len(genomes))
    with pb.Context() as fresh, pytest.raises(pb.PhyloError):
        fresh.process_again(0)


@pytest.mark.parametrize("threads", [1, 3])
@pytest.mark.parametrize("flags", [0, 4])
def test_batched_process_from_ordinary_memory(pb, oracle, threads, flags):
    """phylo_process on plain (pageable) buffers with many batches: sequences staged through
    the pinned rings by worker threads, tile columns compared as the batches get mapped;
    40 genomes = three tile columns of 16"""
    rng = np.random.default_rng(77)
    r = datasets.random_dna(rng, 9000)
    genomes = [r] + [datasets.mutate(rng, r, 0.002 * (k + 1)) for k in range(39)]
    genomes[7] = datasets.revcomp(genomes[7])
    genomes[21] = genomes[21][:4000] + b"!" + genomes[21][4000:]
    want = oracle.process(genomes, 0, flags, threads=8)
    with pb.Context(map_batch_bytes=9000 * 16, stage_threads=threads) as c:
        subst, homol = c.process(genomes, 0, flags)
        assert c.stat("process.pageable") == 1 and c.stat("map.batches") >= 3
        assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])
        # same through pinned memory (the plain asynchronous copies)
        import torch

        pinned = [torch.frombuffer(bytearray(g), dtype=torch.uint8).pin_memory() for g in genomes]
        out = c.process_ptrs([t.data_ptr() for t in pinned], [len(g) for g in genomes], 0, flags)
        assert c.stat("process.pageable") == 0
        assert np.array_equal(out[0], want["subst"]) and np.array_equal(out[1], want["homologs"])


def test_ingest_pipelined_with_the_upload(pb, oracle):
    """phylo_ingest_*: sequences handed over one by one from several threads (as a FASTA
    parser would), each packed and uploaded at once; then process() on what is resident"""
    genomes = _sharded_family(seed=41, n=250000, count=9)  # several pieces of 2 MB? no: one each; plus a long one
    rng = np.random.default_rng(8)
    genomes.append(datasets.mutate(rng, genomes[0] * 10, 0.01))  # 2.5 Mbp: two pieces
    genomes[5] = genomes[5][:100000] + b"!" + genomes[5][100000:]
    want = oracle.process(genomes, 0, 0, threads=8)
    with pb.Context() as c:
        c.ingest(genomes, max_lens=[len(g) + 1000 for g in genomes], lanes=3, threads=4)
        subst, homol = c.process_again(0)
        assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"])
        want4 = oracle.process(genomes, 4, 4, threads=8)
        subst, homol = c.process_again(4, flags=4)
        assert np.array_equal(subst, want4["subst"]) and np.array_equal(homol, want4["homologs"])
        # a byte outside the alphabet, a sequence longer than announced, a missing sequence
        with pytest.raises(pb.PhyloError):
            c.ingest([genomes[0], genomes[1][:500] + b"n" + genomes[1][500:]])
        with pytest.raises(pb.PhyloError):
            c.ingest([genomes[0], genomes[1]], max_lens=[len(genomes[0]), 10])
        lens = np.array([len(genomes[0]), len(genomes[1])], dtype=np.uint64)
        assert c.lib.phylo_ingest_begin(c.h, 2, lens.ctypes.data, 2) == 0
        assert c.lib.phylo_ingest_put(c.h, 0, genomes[0], len(genomes[0])) == 0
        assert c.lib.phylo_ingest_end(c.h) != 0
        with pytest.raises(pb.PhyloError):
            c.process_again(0)  # nothing resident after a failed ingest
        subst, homol = c.process(genomes, 0, 0)  # the context is still usable
        assert np.array_equal(homol, want["homologs"])


def test_upload_paths_agree(pb, oracle):
    """sequences cross PCIe packed to 2 bits per base (default) or as bytes (upload_raw); a
    piece with more separators than the packed form lists (4096) goes over as it is"""
    rng = np.random.default_rng(31)
    r = datasets.random_dna(rng, 120000)
    shredded = b"!".join(r[k : k + 20] for k in range(0, 110000, 20))  # 5500 contigs of 20 bases
    genomes = [r, datasets.mutate(rng, r, 0.01), shredded, datasets.revcomp(datasets.mutate(rng, r, 0.02))]
    want = oracle.process(genomes, 0, 0, threads=4)
    for raw in (0, 1, -1):
        with pb.Context(upload_raw=raw) as c:
            subst, homol = c.process(genomes, 0, 0)
            assert c.stat("process.packed") == (0 if raw == 1 else 1)  # Python bytes are pageable memory
            assert np.array_equal(subst, want["subst"]) and np.array_equal(homol, want["homologs"]), raw
            thr = pb.threshold_for(r)
            c.esa_build(r)
            c.map_queries(genomes, thr)
            s2, h2 = c.compare_all()
            assert np.array_equal(s2, want["subst"]) and np.array_equal(h2, want["homologs"]), raw
            with pytest.raises(pb.PhyloError):
                c.process([r, r[:5000] + b"N" + r[5000:]], 0, 0)
            with pytest.raises(pb.PhyloError):
                c.process([r, r[:70000] + b"a"], 0, 0)
            subst, homol = c.process(genomes, 0, 0)  # still usable
            assert np.array_equal(homol, want["homologs"])


def test_properties_without_oracle(pb, ctx):
    """size-independent properties: symmetric matrix, zero diagonal, identical genomes have
    distance zero and full coverage, subst <= homologs <= min(lengths)"""
    rng = np.random.default_rng(3)
    r = datasets.random_dna(rng, 300000)
    genomes = [r, r, datasets.mutate(rng, r, 0.01), datasets.revcomp(datasets.mutate(rng, r, 0.03))]
    subst, homol = ctx.process(genomes, 0, 0)
    assert np.array_equal(subst, subst.T) and np.array_equal(homol, homol.T)
    assert not subst.diagonal().any() and not homol.diagonal().any()
    assert subst[0, 1] == 0 and homol[0, 1] == len(r)
    assert (subst <= homol).all() and (homol <= len(r)).all()
    # sortedness / disjointness of every filtered list
    for k in range(len(genomes)):
        h = ctx.homologies(k)
        ends = h["index_reference_projected"] + h["length"]
        assert (h["index_reference_projected"][1:] >= ends[:-1]).all()


def test_error_behaviour(pb):
    with pb.Context() as ctx:
        with pytest.raises(pb.PhyloError):
            ctx.map_queries([b"ACGT"], 5)  # no index yet
        with pytest.raises(pb.PhyloError):
            ctx.esa_build(b"")
        with pytest.raises(pb.PhyloError):
            ctx.esa_build(b"ACGTNNACGT")
        ctx.esa_build(b"ACGTACGTTTGACCA")
        with pytest.raises(pb.PhyloError):
            ctx.map_queries([b"ACGTXACGT"], 5)
        with pytest.raises(pb.PhyloError):
            ctx.map_queries([b"ACGTACGT"], 0)
        with pytest.raises(pb.PhyloError):
            ctx.set_option("nonsense", 1)
        # still usable afterwards
        ctx.map_queries([b"ACGTACGTTTGACCA", b""], 5)
        subst, homol = ctx.compare_all()
        assert homol[0, 1] == 0
