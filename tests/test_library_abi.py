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
