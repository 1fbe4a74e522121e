"""Simulated benchmark genomes (same generator as the reference's test/simf, see
host/simgen.cxx).  simf_set(seed, length, dists) equals the sequences written by
`simf -s seed -l length -d d1 -d d2 …` (/root/reference/test/simf.cxx:70-90)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libphylo_simgen.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise ImportError(f"{_PATH} is missing: run `make -C phylonium_b200/host`")
        _lib = C.CDLL(_PATH)
        _lib.phylo_simgen.argtypes = [C.c_uint32, C.c_uint32, C.c_int64, C.c_double, C.c_int, C.c_void_p]
        _lib.phylo_simgen.restype = None
    return _lib


def simf(base_seed: int, mut_seed: int, length: int, divergence: float, raw: bool = False, out=None) -> bytes:
    """One genome; with `out` (a writable buffer address) nothing is returned."""
    lib = _load()
    if out is not None:
        lib.phylo_simgen(base_seed, mut_seed, length, divergence, int(raw), out)
        return b""
    buf = C.create_string_buffer(length)
    lib.phylo_simgen(base_seed, mut_seed, length, divergence, int(raw), buf)
    return buf.raw[:length]


def simf_set(seed: int, length: int, dists) -> list:
    return [simf(seed, seed + i, length, d) for i, d in enumerate([0.0] + list(dists))]
