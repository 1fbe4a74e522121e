"""ctypes binding of include/phylonium_b200.h (libphylonium_b200.so, built in-tree by
phylonium_b200/csrc/Makefile).  There is no CPU path: loading fails loudly if the library
is missing, and creating a context fails loudly without a CUDA device."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PHYLONIUM_B200_LIB: another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("PHYLONIUM_B200_LIB") or os.path.join(_HERE, "libphylonium_b200.so")

PHYLO_FLAG_COMPLETE_DELETION = 4
DIST_RAW, DIST_JC, DIST_ANI = 0, 1, 2

HOM_DTYPE = np.dtype(
    [
        ("direction", "<i8"),
        ("index_reference", "<i8"),
        ("index_reference_projected", "<i8"),
        ("index_query", "<i8"),
        ("length", "<i8"),
    ]
)

_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_vpp = C.POINTER(C.c_void_p)

# every exported symbol of the header with (restype, argtypes); tests check the library
# exports exactly these
SIGNATURES = {
    "phylo_ctx_create": (C.c_int, [C.c_int, _vpp]),
    "phylo_ctx_destroy": (None, [C.c_void_p]),
    "phylo_last_error": (C.c_char_p, [C.c_void_p]),
    "phylo_version": (C.c_char_p, []),
    "phylo_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "phylo_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "phylo_get_stat": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)]),
    "phylo_gc_content": (C.c_double, [C.c_char_p, C.c_uint64]),
    "phylo_min_anchor_length": (C.c_uint64, [C.c_double, C.c_double, C.c_uint64]),
    "phylo_host_pack_2bit": (C.c_int, [C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "phylo_esa_build": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64]),
    "phylo_esa_size": (C.c_int, [C.c_void_p, _u64p]),
    "phylo_esa_get_arrays": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5),
    "phylo_esa_get_matches": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]),
    "phylo_map_queries": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_uint64, C.c_uint64]),
    "phylo_homology_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "phylo_get_homologies": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64, _u64p]),
    "phylo_compare_all": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "phylo_estimate": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "phylo_process": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]),
    "phylo_process_again": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]),
    "phylo_ingest_begin": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]),
    "phylo_ingest_put": (C.c_int, [C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint64]),
    "phylo_ingest_end": (C.c_int, [C.c_void_p]),
    "phylo_esa_build_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "phylo_map_queries_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "phylo_compare_all_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "phylo_esa_alloc": (C.c_int, [C.c_void_p, C.c_uint64]),
    "phylo_esa_device_arrays": (C.c_int, [C.c_void_p, _vpp, _u64p, _vpp, _vpp, _vpp, _vpp]),
    "phylo_esa_finish_import": (C.c_int, [C.c_void_p]),
    "phylo_rows_configure": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64]),
    "phylo_rows_device": (C.c_int, [C.c_void_p, _vpp, _u64p, _u64p]),
    "phylo_rows_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "phylo_rows_ipc_import": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "phylo_rows_set_peers": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "phylo_core_sites": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "phylo_compare_tiles_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}


class PhyloError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"phylonium_b200 error {code}: {message}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree shared library; no fallback of any kind."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C phylonium_b200/csrc` "
                "(or __graft_entry__.build()); phylonium_b200 has no CPU implementation"
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _b(s) -> bytes:
    if isinstance(s, bytes):
        return s
    if isinstance(s, str):
        return s.encode()
    return bytes(s)


def gc_content(seq) -> float:
    seq = _b(seq)
    return load_library().phylo_gc_content(seq, len(seq))


def min_anchor_length(p: float, gc: float, l: int) -> int:
    return int(load_library().phylo_min_anchor_length(p, gc, l))


def threshold_for(ref, p_value: float = 0.025) -> int:
    """process() derives the anchor threshold like this (/root/reference/src/process.cxx:416-417)."""
    ref = _b(ref)
    return min_anchor_length(p_value, gc_content(ref), 2 * len(ref) + 1)


class Context:
    """One phylo_ctx (one CUDA device, one stream)."""

    def __init__(self, device: int = -1, **options):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.phylo_ctx_create(device, C.byref(h))
        if rc != 0:
            raise PhyloError(rc, self.lib.phylo_last_error(None).decode())
        self.h = h
        self.N = 0
        for k, v in options.items():
            self.set_option(k, v)

    def close(self):
        if getattr(self, "h", None):
            self.lib.phylo_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise PhyloError(rc, self.lib.phylo_last_error(self.h).decode())

    def set_option(self, key: str, value: int):
        self._check(self.lib.phylo_set_option(self.h, key.encode(), int(value)))

    def set_stream(self, stream_ptr):
        self._check(self.lib.phylo_set_stream(self.h, stream_ptr))

    def stat(self, key: str) -> float:
        out = C.c_double()
        self._check(self.lib.phylo_get_stat(self.h, key.encode(), C.byref(out)))
        return out.value

    # ---- stage 1 ------------------------------------------------------------
    def esa_build(self, ref):
        ref = _b(ref)
        self._check(self.lib.phylo_esa_build(self.h, ref, len(ref)))

    def esa_build_ptr(self, host_ptr: int, n: int):
        """phylo_esa_build on caller-owned host memory (e.g. a pinned buffer)"""
        self._check(self.lib.phylo_esa_build(self.h, C.c_char_p(int(host_ptr)), n))

    def esa_build_dev(self, d_ptr: int, n: int):
        self._check(self.lib.phylo_esa_build_dev(self.h, d_ptr, n))

    def esa_size(self) -> int:
        m = C.c_uint64()
        self._check(self.lib.phylo_esa_size(self.h, C.byref(m)))
        return m.value

    def esa_arrays(self):
        m = self.esa_size()
        SA = np.zeros(m, np.int64)
        LCP = np.zeros(m + 1, np.int64)
        CLD = np.zeros(m + 1, np.int64)
        FVC = np.zeros(m, np.uint8)
        S = np.zeros(m, np.uint8)
        self._check(self.lib.phylo_esa_get_arrays(self.h, SA.ctypes.data, LCP.ctypes.data, CLD.ctypes.data, FVC.ctypes.data, S.ctypes.data))
        return {"SA": SA, "LCP": LCP, "CLD": CLD, "FVC": FVC, "S": S}

    def get_matches(self, text, offs, lens, use_table: bool = True) -> np.ndarray:
        text = _b(text)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint64)
        out = np.zeros((len(offs), 3), np.int64)
        self._check(self.lib.phylo_esa_get_matches(self.h, text, offs.ctypes.data, lens.ctypes.data, len(offs), int(use_table), out.ctypes.data))
        return out

    # ---- stage 2 ------------------------------------------------------------
    def map_queries(self, queries, threshold: int):
        qs = [_b(q) for q in queries]
        N = len(qs)
        arr = (C.c_char_p * max(N, 1))(*qs)
        lens = np.array([len(q) for q in qs], dtype=np.uint64)
        self._check(self.lib.phylo_map_queries(self.h, arr, lens.ctypes.data, N, threshold))
        self.N = N

    def map_queries_ptrs(self, ptrs, lens, threshold: int):
        """phylo_map_queries on caller-owned host memory: ptrs are addresses"""
        N = len(ptrs)
        arr = (C.c_char_p * max(N, 1))(*[C.c_char_p(int(p)) for p in ptrs])
        lens = np.ascontiguousarray(lens, dtype=np.uint64)
        self._check(self.lib.phylo_map_queries(self.h, arr, lens.ctypes.data, N, threshold))
        self.N = N

    def map_queries_dev(self, d_ptr: int, offs, lens, threshold: int):
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint64)
        self._check(self.lib.phylo_map_queries_dev(self.h, d_ptr, offs.ctypes.data, lens.ctypes.data, len(offs), threshold))
        self.N = len(offs)

    def homology_counts(self, raw: bool = False) -> np.ndarray:
        out = np.zeros(max(self.N, 1), np.uint64)
        self._check(self.lib.phylo_homology_counts(self.h, out.ctypes.data, int(raw)))
        return out[: self.N]

    def homologies(self, index: int, raw: bool = False) -> np.ndarray:
        n = C.c_uint64()
        self._check(self.lib.phylo_get_homologies(self.h, index, int(raw), None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=HOM_DTYPE)
        if n.value:
            self._check(self.lib.phylo_get_homologies(self.h, index, int(raw), out.ctypes.data, n.value, C.byref(n)))
        return out

    # ---- stage 3 ------------------------------------------------------------
    def compare_all(self, flags: int = 0, total: int | None = None):
        N = total if total is not None else self.N
        subst = np.zeros((N, N), np.uint64)
        homol = np.zeros((N, N), np.uint64)
        self._check(self.lib.phylo_compare_all(self.h, flags, subst.ctypes.data, homol.ctypes.data))
        return subst, homol

    def compare_all_dev(self, d_subst: int, d_homologs: int, flags: int = 0):
        self._check(self.lib.phylo_compare_all_dev(self.h, flags, d_subst, d_homologs))

    def compare_tiles_dev(self, d_subst: int, d_homologs: int, rank: int, world: int, flags: int = 0):
        self._check(self.lib.phylo_compare_tiles_dev(self.h, flags, rank, world, d_subst, d_homologs))

    def core_sites(self):
        """(core, border, seg) bitmaps as uint32 arrays (bit b of word w = reference column 32 w + b)"""
        words = C.c_uint64(0)
        self._check(self.lib.phylo_core_sites(self.h, None, None, None, C.byref(words)))
        core, border, seg = (np.zeros(words.value, np.uint32) for _ in range(3))
        self._check(self.lib.phylo_core_sites(self.h, core.ctypes.data, border.ctypes.data, seg.ctypes.data, C.byref(words)))
        return core, border, seg

    def estimate(self, kind: int = DIST_JC, total: int | None = None) -> np.ndarray:
        N = total if total is not None else self.N
        out = np.zeros((N, N), np.float64)
        self._check(self.lib.phylo_estimate(self.h, kind, out.ctypes.data))
        return out

    # ---- the process() seam -------------------------------------------------
    def process(self, seqs, ref_index: int = 0, flags: int = 0):
        qs = [_b(q) for q in seqs]
        N = len(qs)
        arr = (C.c_char_p * N)(*qs)
        lens = np.array([len(q) for q in qs], dtype=np.uint64)
        subst = np.zeros((N, N), np.uint64)
        homol = np.zeros((N, N), np.uint64)
        self._check(self.lib.phylo_process(self.h, arr, lens.ctypes.data, N, ref_index, flags, subst.ctypes.data, homol.ctypes.data))
        self.N = N
        return subst, homol

    def process_ptrs(self, ptrs, lens, ref_index: int = 0, flags: int = 0, out=None):
        """phylo_process on caller-owned host memory (e.g. pinned buffers): ptrs are
        addresses, out = (subst, homologs) numpy uint64 N x N arrays to fill."""
        N = len(ptrs)
        arr = (C.c_char_p * N)(*[C.c_char_p(int(p)) for p in ptrs])
        lens = np.ascontiguousarray(lens, dtype=np.uint64)
        if out is None:
            out = (np.zeros((N, N), np.uint64), np.zeros((N, N), np.uint64))
        self._check(self.lib.phylo_process(self.h, arr, lens.ctypes.data, N, ref_index, flags, out[0].ctypes.data, out[1].ctypes.data))
        self.N = N
        return out

    def ingest(self, seqs, max_lens=None, lanes: int = 4, threads: int = 4):
        """phylo_ingest_begin / _put (from `threads` Python threads, ctypes drops the GIL) / _end;
        afterwards process_again(ref) is process() on the sequences"""
        from concurrent.futures import ThreadPoolExecutor

        qs = [_b(q) for q in seqs]
        caps = np.array(max_lens if max_lens is not None else [len(q) for q in qs], dtype=np.uint64)
        self._check(self.lib.phylo_ingest_begin(self.h, len(qs), caps.ctypes.data, lanes))
        with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
            rcs = list(ex.map(lambda k: self.lib.phylo_ingest_put(self.h, k, qs[k], len(qs[k])), range(len(qs))))
        self._check(self.lib.phylo_ingest_end(self.h))
        assert all(rc == 0 for rc in rcs)
        self.N = len(qs)

    def process_again(self, ref_index: int, flags: int = 0, out=None):
        """second pass of --2pass: same sequences (still on the device), another reference"""
        N = self.N
        if out is None:
            out = (np.zeros((N, N), np.uint64), np.zeros((N, N), np.uint64))
        self._check(self.lib.phylo_process_again(self.h, ref_index, flags, out[0].ctypes.data, out[1].ctypes.data))
        return out

    # ---- multi-GPU plumbing ---------------------------------------------------
    def esa_alloc(self, n: int):
        self._check(self.lib.phylo_esa_alloc(self.h, n))

    def esa_device_arrays(self):
        S, SA, LCP, CLD, FVC = (C.c_void_p() for _ in range(5))
        sb = C.c_uint64()
        self._check(self.lib.phylo_esa_device_arrays(self.h, C.byref(S), C.byref(sb), C.byref(SA), C.byref(LCP), C.byref(CLD), C.byref(FVC)))
        m = self.esa_size()
        return {
            "S": (S.value, sb.value),
            "SA": (SA.value, 4 * m),
            "LCP": (LCP.value, 4 * (m + 1)),
            "CLD": (CLD.value, 4 * (m + 1)),
            "FVC": (FVC.value, m),
        }

    def esa_finish_import(self):
        self._check(self.lib.phylo_esa_finish_import(self.h))

    def rows_configure(self, total: int, first: int):
        self._check(self.lib.phylo_rows_configure(self.h, total, first))

    def rows_device(self):
        p = C.c_void_p()
        b, t = C.c_uint64(), C.c_uint64()
        self._check(self.lib.phylo_rows_device(self.h, C.byref(p), C.byref(b), C.byref(t)))
        return p.value, b.value, t.value

    IPC_HANDLE_BYTES = 64

    def rows_ipc_export(self) -> bytes:
        buf = C.create_string_buffer(self.IPC_HANDLE_BYTES)
        self._check(self.lib.phylo_rows_ipc_export(self.h, buf))
        return buf.raw

    def rows_ipc_import(self, handles, rank: int):
        """handles: one 64-byte handle per rank, in rank order"""
        blob = b"".join(handles)
        assert len(blob) == self.IPC_HANDLE_BYTES * len(handles)
        self._check(self.lib.phylo_rows_ipc_import(self.h, blob, len(handles), rank))

    def rows_set_peers(self, ptrs, rank: int):
        """same-process peers: the row store addresses of all ranks (None switches the push off)"""
        if not ptrs:
            self._check(self.lib.phylo_rows_set_peers(self.h, None, 0, 0))
            return
        arr = (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])
        self._check(self.lib.phylo_rows_set_peers(self.h, arr, len(ptrs), rank))
