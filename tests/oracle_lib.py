"""ctypes wrapper over the two CPU checkers (oracle/po_api.h).

TEST INFRASTRUCTURE: `port()` loads oracle/_build/libphylo_oracle.so (our CPU
restatement), `reference()` loads oracle/_ref/libphylo_ref.so (the unmodified
reference sources, built by oracle/Makefile when /root/reference is present).
Nothing under phylonium_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from functools import lru_cache

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "_build", "libphylo_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libphylo_ref.so")

HOM_DTYPE = np.dtype(
    [
        ("direction", "<i8"),
        ("index_reference", "<i8"),
        ("index_reference_projected", "<i8"),
        ("index_query", "<i8"),
        ("length", "<i8"),
    ]
)

_c_i64p = C.POINTER(C.c_int64)
_c_u64p = C.POINTER(C.c_uint64)


def build(target: str = "all") -> None:
    """Run oracle/Makefile (port always; ref only where /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, target], check=True)


def _as_bytes(s) -> bytes:
    if isinstance(s, bytes):
        return s
    if isinstance(s, str):
        return s.encode()
    return bytes(s)


class OracleLib:
    def __init__(self, path: str):
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.po_kind.restype = C.c_char_p
        L.po_revcomp.argtypes = [C.c_char_p, C.c_int64, C.c_char_p]
        L.po_filter_nucl.argtypes = [C.c_char_p, C.c_int64, C.c_char_p]
        L.po_filter_nucl.restype = C.c_int64
        L.po_gc_content.argtypes = [C.c_char_p, C.c_int64]
        L.po_gc_content.restype = C.c_double
        for f in (L.po_seqcmp, L.po_revseqcmp):
            f.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]
            f.restype = C.c_uint64
        L.po_min_anchor_length.argtypes = [C.c_double, C.c_double, C.c_int64]
        L.po_min_anchor_length.restype = C.c_int64
        L.po_esa_create.argtypes = [C.c_char_p, C.c_int64]
        L.po_esa_create.restype = C.c_void_p
        L.po_esa_destroy.argtypes = [C.c_void_p]
        L.po_esa_size.argtypes = [C.c_void_p]
        L.po_esa_size.restype = C.c_int64
        L.po_esa_arrays.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.po_get_match.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int, _c_i64p]
        L.po_anchor_homologies.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64]
        L.po_anchor_homologies.restype = C.c_int64
        L.po_sort_filter.argtypes = [C.c_void_p, C.c_int64, C.c_int]
        L.po_sort_filter.restype = C.c_int64
        L.po_compare.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_char_p, C.c_void_p, C.c_int64, _c_u64p]
        L.po_complete_delete.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
        L.po_complete_delete.restype = C.c_int64
        L.po_process.argtypes = [
            C.POINTER(C.c_char_p), _c_i64p, C.c_int64, C.c_int64, C.c_int, C.c_int,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
        ]
        L.po_process.restype = C.c_int
        L.po_process_rows.argtypes = [
            C.POINTER(C.c_char_p), _c_i64p, C.c_int64, C.c_int64, C.c_int, C.c_int,
            C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
        ]
        L.po_process_rows.restype = C.c_int
        L.po_estimate.argtypes = [C.c_uint64, C.c_uint64, C.c_int]
        L.po_estimate.restype = C.c_double
        L.po_format_matrix.argtypes = [C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_char_p, C.c_int64]
        L.po_format_matrix.restype = C.c_int64
        L.po_simf.argtypes = [C.c_uint32, C.c_uint32, C.c_int64, C.c_double, C.c_int, C.c_char_p]

    # ---- small helpers -------------------------------------------------
    @property
    def kind(self) -> str:
        return self.lib.po_kind().decode()

    def revcomp(self, s) -> bytes:
        s = _as_bytes(s)
        out = C.create_string_buffer(len(s) + 1)
        self.lib.po_revcomp(s, len(s), out)
        return out.raw[: len(s)]

    def filter_nucl(self, s) -> bytes:
        s = _as_bytes(s)
        out = C.create_string_buffer(len(s) + 1)
        n = self.lib.po_filter_nucl(s, len(s), out)
        return out.raw[:n]

    def gc_content(self, s) -> float:
        s = _as_bytes(s)
        return self.lib.po_gc_content(s, len(s))

    def seqcmp(self, a, b) -> int:
        a, b = _as_bytes(a), _as_bytes(b)
        return self.lib.po_seqcmp(a, b, min(len(a), len(b)))

    def revseqcmp(self, a, b) -> int:
        a, b = _as_bytes(a), _as_bytes(b)
        return self.lib.po_revseqcmp(a, b, min(len(a), len(b)))

    def min_anchor_length(self, p: float, gc: float, l: int) -> int:
        return self.lib.po_min_anchor_length(p, gc, l)

    def threshold(self, ref) -> int:
        """src/process.cxx:416-417 with ANCHOR_P_VALUE = 0.025."""
        ref = _as_bytes(ref)
        return self.min_anchor_length(0.025, self.gc_content(ref), 2 * len(ref) + 1)

    # ---- ESA -------------------------------------------------------------
    def esa(self, ref) -> "OracleEsa":
        return OracleEsa(self, _as_bytes(ref))

    def sort_filter(self, homs: np.ndarray, do_sort: bool = True) -> np.ndarray:
        h = np.ascontiguousarray(homs, dtype=HOM_DTYPE).copy()
        k = self.lib.po_sort_filter(h.ctypes.data, len(h), int(do_sort))
        return h[:k]

    def compare(self, qa, ha: np.ndarray, qb, hb: np.ndarray):
        qa, qb = _as_bytes(qa), _as_bytes(qb)
        ha = np.ascontiguousarray(ha, dtype=HOM_DTYPE)
        hb = np.ascontiguousarray(hb, dtype=HOM_DTYPE)
        out = (C.c_uint64 * 2)()
        self.lib.po_compare(qa, ha.ctypes.data, len(ha), qb, hb.ctypes.data, len(hb), out)
        return int(out[0]), int(out[1])

    def complete_delete(self, lists):
        offs = np.zeros(len(lists) + 1, dtype=np.int64)
        for g, l in enumerate(lists):
            offs[g + 1] = offs[g] + len(l)
        flat = np.concatenate([np.ascontiguousarray(l, dtype=HOM_DTYPE) for l in lists]) if offs[-1] else np.zeros(0, HOM_DTYPE)
        cap = int(sum(len(l) for l in lists)) * len(lists) + 16
        out = np.zeros(cap, dtype=HOM_DTYPE)
        out_offs = np.zeros(len(lists) + 1, dtype=np.int64)
        w = self.lib.po_complete_delete(flat.ctypes.data, offs.ctypes.data, len(lists), out.ctypes.data, out_offs.ctypes.data, cap)
        assert w <= cap
        return [out[out_offs[g] : out_offs[g + 1]].copy() for g in range(len(lists))]

    def process(self, seqs, ref_index: int = 0, flags: int = 0, threads: int = 1, timed: bool = False):
        seqs = [_as_bytes(s) for s in seqs]
        N = len(seqs)
        arr = (C.c_char_p * N)(*seqs)
        lens = np.array([len(s) for s in seqs], dtype=np.int64)
        subst = np.zeros(N * N, dtype=np.uint64)
        homol = np.zeros(N * N, dtype=np.uint64)
        timings = np.zeros(4, dtype=np.float64)
        hcount = np.zeros(N, dtype=np.int64)
        rc = self.lib.po_process(
            arr, lens.ctypes.data_as(_c_i64p), N, ref_index, flags, threads,
            subst.ctypes.data, homol.ctypes.data,
            timings.ctypes.data if timed else None, hcount.ctypes.data if timed else None,
        )
        assert rc == 0
        res = {"subst": subst.reshape(N, N), "homologs": homol.reshape(N, N)}
        if timed:
            res["timings"] = {"esa": timings[0], "anchor": timings[1], "compare": timings[2], "sa_sort": timings[3]}
            res["hom_counts"] = hcount
        return res

    def process_rows(self, seqs, rows, ref_index: int = 0, flags: int = 0, threads: int = 1, lens=None):
        """rows `rows` of process()'s matrix (every sequence is mapped, only those rows are
        compared).  seqs: byte strings, or raw host addresses together with `lens`."""
        if lens is None:
            seqs = [_as_bytes(s) for s in seqs]
            lens = [len(s) for s in seqs]
            arr = (C.c_char_p * len(seqs))(*seqs)
        else:
            arr = (C.c_char_p * len(seqs))(*[C.c_char_p(int(p)) for p in seqs])
        N = len(seqs)
        lens = np.ascontiguousarray(lens, dtype=np.int64)
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        subst = np.zeros(len(rows) * N, dtype=np.uint64)
        homol = np.zeros(len(rows) * N, dtype=np.uint64)
        timings = np.zeros(4, dtype=np.float64)
        rc = self.lib.po_process_rows(arr, lens.ctypes.data_as(_c_i64p), N, ref_index, flags, threads, rows.ctypes.data,
                                      len(rows), subst.ctypes.data, homol.ctypes.data, timings.ctypes.data)
        assert rc == 0
        return {"subst": subst.reshape(len(rows), N), "homologs": homol.reshape(len(rows), N),
                "timings": {"esa": timings[0], "anchor": timings[1], "compare": timings[2], "sa_sort": timings[3]}}

    def estimate(self, subst: int, homologs: int, kind: int = 1) -> float:
        return self.lib.po_estimate(int(subst), int(homologs), kind)

    def format_matrix(self, names, subst, homologs, kind: int = 1) -> str:
        N = len(names)
        arr = (C.c_char_p * N)(*[_as_bytes(n) for n in names])
        s = np.ascontiguousarray(subst, dtype=np.uint64).reshape(-1)
        h = np.ascontiguousarray(homologs, dtype=np.uint64).reshape(-1)
        cap = 64 + N * 64 + N * N * 16 + sum(len(n) for n in names)
        out = C.create_string_buffer(cap)
        k = self.lib.po_format_matrix(arr, s.ctypes.data, h.ctypes.data, N, kind, out, cap)
        assert k < cap
        return out.raw[:k].decode()

    def simf(self, base_seed: int, mut_seed: int, length: int, divergence: float, raw: bool = False) -> bytes:
        out = C.create_string_buffer(length + 1)
        self.lib.po_simf(base_seed, mut_seed, length, divergence, int(raw), out)
        return out.raw[:length]

    def simf_set(self, seed: int, length: int, dists) -> list:
        """Genomes as written by `simf -s seed -l length -d d1 -d d2 …` (test/simf.cxx:70-90):
        genome 0 is the undiverged base, genome i uses mutation seed seed+i."""
        ds = [0.0] + list(dists)
        return [self.simf(seed, seed + i, length, d) for i, d in enumerate(ds)]


class OracleEsa:
    def __init__(self, lib: OracleLib, ref: bytes):
        self.lib = lib
        self.ref = ref
        self.n = len(ref)
        self.handle = lib.lib.po_esa_create(ref, len(ref))
        self.m = lib.lib.po_esa_size(self.handle)

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.lib.po_esa_destroy(self.handle)
            self.handle = None

    def arrays(self):
        m = self.m
        SA = np.zeros(m, np.int64)
        LCP = np.zeros(m + 1, np.int64)
        CLD = np.zeros(m + 1, np.int64)
        FVC = np.zeros(m, np.uint8)
        S = np.zeros(m, np.uint8)
        self.lib.lib.po_esa_arrays(self.handle, SA.ctypes.data, LCP.ctypes.data, CLD.ctypes.data, FVC.ctypes.data, S.ctypes.data)
        return {"SA": SA, "LCP": LCP, "CLD": CLD, "FVC": FVC, "S": S}

    def get_match(self, q, cached: bool = True):
        q = _as_bytes(q)
        out = (C.c_int64 * 3)()
        self.lib.lib.po_get_match(self.handle, q, len(q), int(cached), out)
        return int(out[0]), int(out[1]), int(out[2])

    def anchor_homologies(self, thr: int, q) -> np.ndarray:
        q = _as_bytes(q)
        cap = len(q) // max(1, thr) + 16
        out = np.zeros(cap, dtype=HOM_DTYPE)
        k = self.lib.lib.po_anchor_homologies(self.handle, thr, q, len(q), out.ctypes.data, cap)
        assert k <= cap
        return out[:k].copy()


@lru_cache(maxsize=None)
def port() -> OracleLib:
    if not os.path.exists(PORT_SO):
        build("port")
    return OracleLib(PORT_SO)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


@lru_cache(maxsize=None)
def reference() -> OracleLib:
    if not have_reference():
        raise FileNotFoundError(REF_SO)
    return OracleLib(REF_SO)


def best() -> OracleLib:
    """The strongest checker available: the compiled reference if present, else the port."""
    return reference() if have_reference() else port()
