"""Committed reference outputs (tests/golden/*.npz, generated from the unmodified reference
by tests/golden/make_golden.py) against the CPU restatement — and, on a GPU, against the
CUDA path through the C ABI."""
import glob
import os

import numpy as np
import pytest

import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in FIXTURES]


def load(path):
    z = np.load(path)
    genomes = [z[f"genome_{i}"].tobytes() for i in range(int(z["n_genomes"]))]
    return z, genomes


def test_fixtures_exist():
    assert len(FIXTURES) >= 7


@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_port_reproduces_reference_outputs(path):
    z, genomes = load(path)
    port = oracle_lib.port()
    ref = genomes[0]
    thr = int(z["threshold"])
    assert port.threshold(ref) == thr
    esa = port.esa(ref)
    arr = esa.arrays()
    for k in ("S", "SA", "LCP", "CLD", "FVC"):
        assert np.array_equal(arr[k].astype(np.int64), z[k].astype(np.int64)), k
    text = z["match_text"].tobytes()
    for p, l, want in zip(z["match_pos"], z["match_len"], z["match_out"]):
        assert esa.get_match(text[p : p + l]) == tuple(int(x) for x in want)
    for i, g in enumerate(genomes):
        raw = esa.anchor_homologies(thr, g)
        assert np.array_equal(raw, z[f"raw_{i}"])
        assert np.array_equal(port.sort_filter(raw), z[f"filtered_{i}"])
    res = port.process(genomes, 0, 0, threads=2)
    assert np.array_equal(res["subst"], z["subst"]) and np.array_equal(res["homologs"], z["homologs"])
    if "subst_cd" in z:
        res = port.process(genomes, 0, 4, threads=2)
        assert np.array_equal(res["subst"], z["subst_cd"]) and np.array_equal(res["homologs"], z["homologs_cd"])
    names = [f"g{i}" for i in range(len(genomes))]
    for kind, tag in ((0, "raw"), (1, "jc"), (2, "ani")):
        assert port.format_matrix(names, z["subst"], z["homologs"], kind) == z[f"phylip_{tag}"].tobytes().decode()


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=IDS)
def test_cuda_path_reproduces_reference_outputs(path):
    import phylonium_b200 as pb

    z, genomes = load(path)
    ref = genomes[0]
    thr = int(z["threshold"])
    assert pb.threshold_for(ref) == thr
    with pb.Context(keep_raw=1, chunk=128) as ctx:
        ctx.esa_build(ref)
        arr = ctx.esa_arrays()
        for k in ("S", "SA", "LCP", "CLD", "FVC"):
            assert np.array_equal(arr[k].astype(np.int64), z[k].astype(np.int64)), k
        got = ctx.get_matches(z["match_text"].tobytes(), z["match_pos"], z["match_len"])
        assert np.array_equal(got, z["match_out"])
        ctx.map_queries(genomes, thr)
        for i in range(len(genomes)):
            assert np.array_equal(ctx.homologies(i, raw=True), z[f"raw_{i}"])
            assert np.array_equal(ctx.homologies(i), z[f"filtered_{i}"])
        subst, homol = ctx.compare_all()
        assert np.array_equal(subst, z["subst"]) and np.array_equal(homol, z["homologs"])
        if "subst_cd" in z:
            subst_cd, homol_cd = ctx.compare_all(pb.PHYLO_FLAG_COMPLETE_DELETION)
            assert np.array_equal(subst_cd, z["subst_cd"]) and np.array_equal(homol_cd, z["homologs_cd"])
        names = [f"g{i}" for i in range(len(genomes))]
        mat = [pb.EvoModel(int(subst[i, j]), int(homol[i, j])) for i in range(len(genomes)) for j in range(len(genomes))]
        for kind, tag in ((0, "raw"), (1, "jc"), (2, "ani")):
            assert pb.format_matrix(names, mat, kind) == z[f"phylip_{tag}"].tobytes().decode()
