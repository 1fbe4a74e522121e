"""The suffix sorter standing in for libdivsufsort64 (oracle/sa_standin.cxx) against a
definition-level sort: SA must order the suffixes of S = R#revcomp(R) by unsigned byte,
shorter-prefix first.  This is what pins SA parity."""
import numpy as np
import pytest

import datasets
import oracle_lib


def naive_sa(s: bytes):
    return sorted(range(len(s)), key=lambda i: s[i:])


@pytest.mark.parametrize("name", ["tiny", "bang_vs_base", "repeats", "multi_contig"])
def test_against_definition(name):
    lib = oracle_lib.port()
    ref = datasets.ALL_SETS[name]()[0][:6000]
    a = lib.esa(ref).arrays()
    S = a["S"].tobytes()
    assert S == ref + b"#" + datasets.revcomp(ref)
    assert a["SA"].tolist() == naive_sa(S)


def test_degenerate_texts():
    lib = oracle_lib.port()
    for ref in (b"A", b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", b"ACACACACACACACACACACACACACACACACACACACACACACACAC", b"AT" * 40, b"A!A!A!A"):
        a = lib.esa(ref).arrays()
        S = a["S"].tobytes()
        assert a["SA"].tolist() == naive_sa(S)
        # LCP by definition
        for r in range(1, len(S)):
            x, y = S[a["SA"][r - 1] :], S[a["SA"][r] :]
            l = 0
            while l < min(len(x), len(y)) and x[l] == y[l]:
                l += 1
            assert a["LCP"][r] == l
