"""ctypes binding of oracle/libgpe_oracle.so -- the CPU restatement of the reference.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (gnn_pe_b200) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile the restatement (and, when /root/reference is present, oracle/_ref)."""
    so = os.path.join(_HERE, "libgpe_oracle.so")
    src = os.path.join(_HERE, "gpe_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libgpe_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_graph_load.restype = C.c_void_p
        L.orc_graph_load.argtypes = [C.c_char_p]
        L.orc_graph_from_csr.restype = C.c_void_p
        L.orc_graph_from_csr.argtypes = [C.c_uint32, _u32p, _u32p, _u32p]
        L.orc_graph_free.argtypes = [C.c_void_p]
        L.orc_graph_meta.argtypes = [C.c_void_p, _u32p]
        L.orc_graph_csr.argtypes = [C.c_void_p, _u32p, _u32p, _u32p]
        L.orc_label_embedding.argtypes = [C.c_uint32, C.c_uint32, _f64p]
        L.orc_vertex_embeddings.argtypes = [C.c_void_p, C.c_uint32, _f64p, _f64p]
        L.orc_degree_order.argtypes = [C.c_void_p, _u32p]
        L.orc_enumerate.restype = C.c_uint64
        L.orc_enumerate.argtypes = [C.c_void_p, C.c_uint32, _u32p, C.c_int]
        L.orc_paths_copy.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, _u32p]
        L.orc_rows_per_partition.argtypes = [C.c_void_p, _u32p, C.c_uint32, _u64p]
        L.orc_query_plan.restype = C.c_uint32
        L.orc_query_plan.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, _u32p, _u32p,
                                     _u32p, _f64p, _u32p, C.POINTER(C.c_uint32)]
        L.orc_filter.restype = C.c_uint64
        L.orc_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, _u64p, C.c_void_p, C.c_uint64,
                                 C.c_void_p, C.c_uint32]
        L.orc_query_connected.restype = C.c_int
        L.orc_query_connected.argtypes = [C.c_void_p]
        L.orc_matching_order.argtypes = [C.c_void_p, C.c_void_p, _u32p, _u32p, _u32p]
        L.orc_refine.restype = C.c_uint64
        L.orc_refine.argtypes = [C.c_void_p, C.c_void_p, _u64p, _u32p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.orc_pge_groups.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, _f64p, _f64p, C.c_void_p]
        L.orc_pge_filter.restype = C.c_uint64
        L.orc_pge_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, _u64p, C.c_void_p, C.c_uint64]
        L.orc_pge_online.restype = C.c_uint64
        L.orc_pge_online.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64]
        L.orc_online.restype = C.c_uint64
        L.orc_online.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64]
        L.orc_online_streaming.restype = C.c_uint64
        L.orc_online_streaming.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, _u32p, _f64p,
                                           C.c_uint64, _f64p, C.c_int]
        _LIB = L
    return _LIB


UINT_MAX = 0xFFFFFFFF


class OracleGraph:
    """A graph held by the oracle; also carries the enumerated path table once built."""

    def __init__(self, handle):
        if not handle:
            raise OSError("oracle: cannot open graph")
        self._h = handle
        meta = np.zeros(5, dtype=np.uint32)
        lib().orc_graph_meta(self._h, meta)
        self.V, self.E, self.labels_count, self.max_degree, self.max_label_freq = (int(x) for x in meta)
        self.L = 0
        self.n_rows = 0

    @classmethod
    def load(cls, path: str) -> "OracleGraph":
        return cls(lib().orc_graph_load(path.encode()))

    @classmethod
    def from_csr(cls, offsets, nbrs, labels) -> "OracleGraph":
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        nbrs = np.ascontiguousarray(nbrs, dtype=np.uint32)
        labels = np.ascontiguousarray(labels, dtype=np.uint32)
        if nbrs.size == 0:
            nbrs = np.zeros(1, dtype=np.uint32)
        return cls(lib().orc_graph_from_csr(len(labels), offsets, nbrs, labels))

    def __del__(self):
        try:
            lib().orc_graph_free(self._h)
        except Exception:
            pass

    def csr(self):
        off = np.zeros(self.V + 1, dtype=np.uint32)
        nbr = np.zeros(max(2 * self.E, 1), dtype=np.uint32)
        lab = np.zeros(max(self.V, 1), dtype=np.uint32)
        lib().orc_graph_csr(self._h, off, nbr, lab)
        return off, nbr[: 2 * self.E], lab[: self.V]

    def embeddings(self, e: int):
        x = np.zeros((max(self.V, 1), e), dtype=np.float64)
        vde = np.zeros((max(self.V, 1), e), dtype=np.float64)
        lib().orc_vertex_embeddings(self._h, e, x, vde)
        return x[: self.V], vde[: self.V]

    def degree_order(self) -> np.ndarray:
        out = np.zeros(max(self.V, 1), dtype=np.uint32)
        lib().orc_degree_order(self._h, out)
        return out[: self.V]

    def enumerate(self, L: int, sorted_nodes, literal: bool = False) -> int:
        sn = np.ascontiguousarray(sorted_nodes, dtype=np.uint32)
        self.n_rows = int(lib().orc_enumerate(self._h, L, sn, 1 if literal else 0))
        self.L = L
        return self.n_rows

    def paths(self, first: int = 0, n: int | None = None) -> np.ndarray:
        n = self.n_rows - first if n is None else n
        out = np.zeros((max(n, 1), self.L), dtype=np.uint32)
        if n:
            lib().orc_paths_copy(self._h, first, n, out)
        return out[:n]

    def rows_per_partition(self, membership, p: int) -> np.ndarray:
        out = np.zeros(p, dtype=np.uint64)
        lib().orc_rows_per_partition(self._h, np.ascontiguousarray(membership, dtype=np.uint32), p, out)
        return out

    def all_paths_text(self) -> str:
        """The bytes main.cpp:110-119 writes to all_paths.txt."""
        rows = self.paths()
        lines = [str(self.n_rows)]
        lines += [" ".join(map(str, r)) + " " for r in rows.tolist()]
        return "\n".join(lines) + "\n"


def label_embedding(label: int, e: int) -> np.ndarray:
    out = np.zeros(e, dtype=np.float64)
    lib().orc_label_embedding(label, e, out)
    return out


def query_plan(q: OracleGraph, L: int, e: int, literal: bool = False):
    cap = 4096
    vids = np.zeros((cap, L), dtype=np.uint32)
    labels = np.zeros((cap, L), dtype=np.uint32)
    degs = np.zeros((cap, L), dtype=np.uint32)
    pde = np.zeros((cap, L * e), dtype=np.float64)
    weight = np.zeros(cap, dtype=np.uint32)
    nqp = C.c_uint32(0)
    n = lib().orc_query_plan(q._h, L, e, 1 if literal else 0, cap, vids, labels, degs, pde, weight, C.byref(nqp))
    return dict(vids=vids[:n], labels=labels[:n], degrees=degs[:n], pde=pde[:n], weight=weight[:n],
                n_query_paths=int(nqp.value))


def filter_candidates(g: OracleGraph, q: OracleGraph, e: int, literal_plan: bool = False,
                      both_orientations: bool = False):
    """Brute-force dominance filter over g's enumerated table.  Returns (list of sorted
    candidate arrays per query vertex, survivors per plan path).  both_orientations: the exact mode of
    SURVEY.md 8f-4 (every plan path against both orientations of every stored row) -- not the reference's rule."""
    mode = int(bool(literal_plan)) | (2 if both_orientations else 0)
    off = np.zeros(q.V + 1, dtype=np.uint64)
    total = int(lib().orc_filter(g._h, q._h, e, mode, off, None, 0, None, 0))
    cand = np.zeros(max(total, 1), dtype=np.uint32)
    surv = np.zeros(4096, dtype=np.uint64)
    lib().orc_filter(g._h, q._h, e, mode, off, cand.ctypes.data_as(C.c_void_p), total,
                     surv.ctypes.data_as(C.c_void_p), len(surv))
    n_plan = len(query_plan(q, g.L, e, literal_plan)["weight"])
    sets = [cand[int(off[u]): int(off[u + 1])].copy() for u in range(q.V)]
    return sets, surv[:n_plan]


def matching_order(g: OracleGraph, q: OracleGraph, counts):
    order = np.zeros(max(q.V, 1), dtype=np.uint32)
    pivot = np.zeros(max(q.V, 1), dtype=np.uint32)
    lib().orc_matching_order(g._h, q._h, np.ascontiguousarray(counts, dtype=np.uint32), order, pivot)
    return order[: q.V], pivot[: q.V]


def refine(g: OracleGraph, q: OracleGraph, cand_sets, limit: int = UINT_MAX, want_matches: int = 0):
    off = np.zeros(q.V + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(c) for c in cand_sets])
    flat = np.concatenate([np.asarray(c, dtype=np.uint32) for c in cand_sets] + [np.zeros(1, dtype=np.uint32)])
    flat = np.ascontiguousarray(flat, dtype=np.uint32)
    if want_matches:
        m = np.zeros((want_matches, q.V), dtype=np.uint32)
        n = int(lib().orc_refine(g._h, q._h, off, flat, limit, m.ctypes.data_as(C.c_void_p), want_matches))
        return n, m[: min(n, want_matches)]
    return int(lib().orc_refine(g._h, q._h, off, flat, limit, None, 0))


def online(g: OracleGraph, q: OracleGraph, e: int, limit: int = UINT_MAX) -> int:
    return int(lib().orc_online(g._h, q._h, e, limit))


def online_streaming(g: OracleGraph, q: OracleGraph, L: int, e: int, sorted_nodes, vde, limit: int = UINT_MAX,
                     threads: int = 0):
    t3 = np.zeros(3, dtype=np.float64)
    n = int(lib().orc_online_streaming(g._h, q._h, L, e, np.ascontiguousarray(sorted_nodes, dtype=np.uint32),
                                       np.ascontiguousarray(vde, dtype=np.float64), limit, t3, threads))
    return n, t3


# ---- GNN-PGE (the reference's per-vertex variant) ----------------------------------------
def pge_groups(g: OracleGraph, pl: int, e: int):
    """Path groups of every vertex: (pg, plg) as V x 2*pl*e arrays [lo0, hi0, lo1, hi1, ...] and has[V]."""
    V = g.V
    pg = np.zeros((max(V, 1), 2 * pl * e), dtype=np.float64)
    plg = np.zeros((max(V, 1), 2 * pl * e), dtype=np.float64)
    has = np.zeros(max(V, 1), dtype=np.uint8)
    lib().orc_pge_groups(g._h, pl, e, pg, plg, has.ctypes.data_as(C.c_void_p))
    return pg[:V], plg[:V], has[:V]


def pge_filter(g: OracleGraph, q: OracleGraph, pl: int, e: int):
    """Candidate set of every query vertex under GNN-PGE's leaf test (sorted ids)."""
    off = np.zeros(q.V + 1, dtype=np.uint64)
    total = int(lib().orc_pge_filter(g._h, q._h, pl, e, off, None, 0))
    cand = np.zeros(max(total, 1), dtype=np.uint32)
    lib().orc_pge_filter(g._h, q._h, pl, e, off, cand.ctypes.data_as(C.c_void_p), total)
    return [cand[int(off[u]): int(off[u + 1])].copy() for u in range(q.V)]


def pge_online(g: OracleGraph, q: OracleGraph, pl: int, e: int, limit: int = UINT_MAX) -> int:
    return int(lib().orc_pge_online(g._h, q._h, pl, e, limit))
