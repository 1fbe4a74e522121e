#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libphylo_ref.so,
built from /root/reference by oracle/Makefile).  The reference's own tests hold no golden
vectors for the hot path, so these fixtures are reference OUTPUTS on small
deterministic inputs; they pin the CPU restatement and the CUDA path even where the
compiled reference is not available.

    python tests/golden/make_golden.py        (run in the build container)

Each fixture stores the input genomes and, from the reference: threshold, S/SA/LCP/CLD/FVC of
the index, longest matches at sampled positions, raw and filtered homology lists per genome,
the substitution/homology count matrices (plain and complete deletion) and the PHYLIP text.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import datasets  # noqa: E402
import oracle_lib  # noqa: E402

# small variants so that the fixtures stay a few hundred kB in total
FIXTURES = {
    "simf_pair": lambda: oracle_lib.reference().simf_set(1, 6000, [0.02]),
    "multi_contig": lambda: datasets.multi_contig_set(seed=3, n=4500),
    "bang_vs_base": lambda: datasets.bang_vs_base_set(seed=4, n=700),
    "repeats": lambda: datasets.repeat_set(seed=5, n=5000),
    "rearranged": lambda: datasets.rearranged_set(seed=7, n=8000, block=800),
    "identical_unrelated": lambda: datasets.identical_and_unrelated_set(seed=8, n=3000),
    "first_bases": lambda: datasets.substitution_in_first_bases(seed=10, n=1500),
}


def build(name):
    ref_lib = oracle_lib.reference()
    genomes = FIXTURES[name]()
    ref = genomes[0]
    thr = ref_lib.threshold(ref)
    esa = ref_lib.esa(ref)
    arr = esa.arrays()
    out = {"n_genomes": np.int64(len(genomes)), "threshold": np.int64(thr)}
    for i, g in enumerate(genomes):
        out[f"genome_{i}"] = np.frombuffer(g, np.uint8)
    for k in ("S", "SA", "LCP", "CLD", "FVC"):
        out[k] = arr[k].astype(np.int32) if arr[k].dtype == np.int64 else arr[k]
    rng = np.random.default_rng(99)
    text = b"".join(genomes[1:3])
    pos = np.sort(rng.integers(0, len(text), size=min(200, len(text))))
    lens = np.minimum(rng.integers(1, 800, size=len(pos)), len(text) - pos)
    out["match_text"] = np.frombuffer(text, np.uint8)
    out["match_pos"], out["match_len"] = pos.astype(np.int64), lens.astype(np.int64)
    out["match_out"] = np.array([esa.get_match(text[p : p + l]) for p, l in zip(pos, lens)], dtype=np.int64)
    lists = []
    for i, g in enumerate(genomes):
        raw = esa.anchor_homologies(thr, g)
        fil = ref_lib.sort_filter(raw)
        out[f"raw_{i}"], out[f"filtered_{i}"] = raw, fil
        lists.append(fil)
    plain = ref_lib.process(genomes, 0, 0, threads=1)
    out["subst"], out["homologs"] = plain["subst"], plain["homologs"]
    if all(len(l) for l in lists):
        cd = ref_lib.process(genomes, 0, 4, threads=1)
        out["subst_cd"], out["homologs_cd"] = cd["subst"], cd["homologs"]
    names = [f"g{i}" for i in range(len(genomes))]
    for kind, tag in ((0, "raw"), (1, "jc"), (2, "ani")):
        out[f"phylip_{tag}"] = np.frombuffer(ref_lib.format_matrix(names, plain["subst"], plain["homologs"], kind).encode(), np.uint8)
    return out


def main():
    assert oracle_lib.have_reference(), "build oracle/_ref first (make -C oracle ref)"
    for name in FIXTURES:
        data = build(name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: {os.path.getsize(path) / 1024:.0f} kB")


if __name__ == "__main__":
    main()
