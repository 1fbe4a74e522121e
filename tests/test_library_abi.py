"""The C-ABI library loads here (no GPU needed for dlopen) and exports every symbol
include/phylonium_b200.h declares; without a CUDA device it refuses to create a context."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "phylonium_b200.h")


@pytest.fixture(scope="module")
def lib():
    from phylonium_b200 import capi

    if not os.path.exists(capi.LIB_PATH):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "phylonium_b200", "csrc")], check=True)
    return capi.load_library()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(phylo_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    from phylonium_b200 import capi

    assert declared_symbols() == sorted(capi.SIGNATURES)


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_host_side_scalars(lib):
    """gc_content / min_anchor_length are host functions: check against the oracle here"""
    import oracle_lib
    from phylonium_b200 import capi

    o = oracle_lib.best()
    for s in (b"ACGT", b"GGGCC!AT", b"A" * 50 + b"C" * 13):
        assert capi.gc_content(s) == o.gc_content(s)
    for gc in (0.3, 0.5, 0.65):
        for l in (11, 2001, 200001, 10000001, 500000001):
            assert capi.min_anchor_length(0.025, gc, l) == o.min_anchor_length(0.025, gc, l)


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import phylonium_b200 as pb

    with pytest.raises(pb.PhyloError) as e:
        pb.Context()
    assert "no usable CUDA device" in str(e.value)
    with pytest.raises(pb.PhyloError):
        pb.process(0, [b"ACGTACGTAA", b"ACGTACGTAA"])


def test_host_packing_of_sequences():
    """the 2-bit packing applied before the upload (host_pack.cpp; AVX2/BMI2 with a scalar tail):
    codes, positions of '!', rejection of every byte outside the alphabet, odd lengths"""
    import numpy as np

    from phylonium_b200 import capi

    lib = capi.load_library()
    rng = np.random.default_rng(4)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    code_of = {ord("A"): 0, ord("C"): 1, ord("T"): 2, ord("G"): 3, ord("!"): 0}
    for n in list(range(0, 70)) + [255, 256, 257, 4099, 100003]:
        seq = alphabet[rng.integers(0, 4, size=n)].copy()
        bang_at = sorted(set(int(x) for x in rng.integers(0, max(1, n), size=min(n, 5)))) if n else []
        for p in bang_at:
            seq[p] = ord("!")
        packed = np.full((n + 3) // 4 + 8, 0xEE, np.uint8)
        bangs = np.zeros(16, np.uint32)
        nb = np.zeros(1, np.uint32)
        rc = lib.phylo_host_pack_2bit(seq.tobytes(), n, packed.ctypes.data, bangs.ctypes.data, 16, nb.ctypes.data)
        assert rc == 0 and int(nb[0]) == len(bang_at) and list(bangs[: len(bang_at)]) == bang_at
        assert (packed[(n + 3) // 4 :] == 0xEE).all()  # nothing written past the end
        codes = np.array([code_of[int(c)] for c in seq], dtype=np.uint8)
        got = (packed[np.arange(n) // 4] >> (2 * (np.arange(n) % 4))) & 3 if n else codes
        assert np.array_equal(got, codes)
        if n:  # every other byte value is an error, wherever it sits
            for bad in (0, ord("N"), ord("a"), 0xFF, 0x4F, ord("#"), ord(" ")):
                s2 = seq.copy()
                s2[int(rng.integers(0, n))] = bad
                assert lib.phylo_host_pack_2bit(s2.tobytes(), n, packed.ctypes.data, bangs.ctypes.data, 16, nb.ctypes.data) == 1
