"""ctypes binding of libgpe.so (include/gpe.h) -- the only way Python reaches the GPU path.

There is no CPU fallback: if the shared library is missing or no B200 is visible, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpe.so")
LIMIT_MAX = 0xFFFFFFFF
FILTER_NO_PRUNE = 1
FILTER_BOTH_ORIENTATIONS = 2  # exact mode: the true number of embeddings, not the reference's under-count
COMM_ID_BYTES = 128

_LIB = None


class GpeError(RuntimeError):
    pass


def build(force: bool = False) -> str:
    """Compile libgpe.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cpp", ".h"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "gpe.h"))
    synth = os.path.join(_HERE, "libgpe_synth.so")  # the seeded generator of the large synthetic graphs (same Makefile)
    stale = (not os.path.exists(LIB_PATH)) or (not os.path.exists(synth)) or \
        any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        if not os.path.exists("/usr/local/cuda/bin/nvcc") and os.path.exists(LIB_PATH) and not force:
            return LIB_PATH
        subprocess.check_call(["make", "-C", csrc, "-j8"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class Batch(C.Structure):
    _fields_ = [("n_queries", C.c_uint32), ("q_vbase", C.c_void_p), ("q_ebase", C.c_void_p),
                ("q_offsets", C.c_void_p), ("q_nbrs", C.c_void_p), ("q_labels", C.c_void_p),
                ("limits", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("table_rows", "table_tiles", "tile_rows", "row_bytes", "scan_items", "scan_items_unpruned",
                 "scan_rows", "scan_launches", "select_launches", "compact_launches", "join_launches",
                 "build_launches")] + \
               [(n, C.c_float) for n in
                ("last_scan_ms", "last_select_ms", "last_compact_ms", "last_join_ms", "last_build_ms",
                 "last_enumerate_ms")] + \
               [(n, C.c_uint64) for n in ("n_qpaths", "n_qblocks", "n_slots", "n_candidates", "join_items",
                                        "kernel_launches", "h2d_bytes", "d2h_bytes", "join_exports", "join_donations", "join_steps", "join_warp_iters", "join_idle_polls", "join_bfs", "join_fallbacks", "join_reruns", "table_ids_only", "stored_row_bytes", "exchange_bytes", "exchange_redos")] + \
               [(n, C.c_float) for n in ("host_upload_ms", "host_enqueue_ms", "host_plan_ms", "host_finish_ms")]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/gpe.h declares (tests check the library exports all of them)
SYMBOLS = [
    "gpe_create", "gpe_destroy", "gpe_last_error", "gpe_abi_version", "gpe_host_load_graph", "gpe_host_gen_vde",
    "gpe_host_query_plan", "gpe_set_graph", "gpe_set_embeddings", "gpe_enumerate", "gpe_dump_paths",
    "gpe_start_rows", "gpe_build_table", "gpe_dump_table", "gpe_filter", "gpe_get_candidates", "gpe_refine",
    "gpe_batch_upload", "gpe_batch_filter", "gpe_batch_join", "gpe_batch_download", "gpe_clamp_answer",
    "gpe_query_batch", "gpe_batch_cand_info", "gpe_batch_cand_export", "gpe_batch_cand_merge",
    "gpe_batch_scan", "gpe_batch_bitmap", "gpe_batch_bitmap_merge",
    "gpe_batch_get_candidates", "gpe_batch_get_plan", "gpe_pge_build", "gpe_host_pge_groups", "gpe_pge_dump_groups", "gpe_pge_batch_upload",
    "gpe_pge_batch_filter", "gpe_pge_query_batch", "gpe_get_stats", "gpe_stream", "gpe_sync", "gpe_set_timing",
    "gpe_collect_timings", "gpe_comm_unique_id", "gpe_comm_init", "gpe_comm_init_all", "gpe_comm_destroy", "gpe_comm_info",
    "gpe_build_table_shard", "gpe_batch_step", "gpe_batch_finish", "gpe_multi_batch_upload", "gpe_multi_batch_step",
    "gpe_multi_batch_finish", "gpe_multi_query_batch", "gpe_query_batches", "gpe_set_table_layout",
]


def lib():
    """Load libgpe.so; raises (loudly) when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise GpeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        L.gpe_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.gpe_destroy.argtypes = [vp]
        L.gpe_destroy.restype = None
        L.gpe_last_error.argtypes = [vp]
        L.gpe_last_error.restype = C.c_char_p
        L.gpe_host_load_graph.argtypes = [C.c_char_p, C.POINTER(u32), C.POINTER(u32), vp, vp, vp]
        L.gpe_host_gen_vde.argtypes = [u32, vp, vp, vp, u32, vp, vp]
        L.gpe_host_query_plan.argtypes = [u32, vp, vp, vp, u32, u32, u32, vp, vp, vp, vp, C.POINTER(u32)]
        L.gpe_set_graph.argtypes = [vp, u32, vp, vp, vp]
        L.gpe_set_embeddings.argtypes = [vp, u32, vp]
        L.gpe_enumerate.argtypes = [vp, u32, vp, vp, u32, vp, C.POINTER(u64)]
        L.gpe_dump_paths.argtypes = [vp, u64, u64, vp]
        L.gpe_start_rows.argtypes = [vp, vp]
        L.gpe_build_table.argtypes = [vp, vp, C.POINTER(u64)]
        L.gpe_dump_table.argtypes = [vp, u64, u64, vp, vp, vp, vp]
        L.gpe_filter.argtypes = [vp, u32, vp, vp, vp, vp, u32, u32, vp, vp]
        L.gpe_get_candidates.argtypes = [vp, vp]
        L.gpe_refine.argtypes = [vp, u32, vp, vp, vp, vp, vp, u64, C.POINTER(u64), vp, vp, vp, u64]
        L.gpe_batch_upload.argtypes = [vp, C.POINTER(Batch), u32]
        L.gpe_batch_filter.argtypes = [vp]
        L.gpe_batch_join.argtypes = [vp, u32, u32]
        L.gpe_batch_download.argtypes = [vp, vp]
        L.gpe_clamp_answer.argtypes = [u64, u64]
        L.gpe_clamp_answer.restype = u64
        L.gpe_query_batch.argtypes = [vp, C.POINTER(Batch), u32, vp]
        L.gpe_batch_cand_info.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
        L.gpe_batch_cand_export.argtypes = [vp, vp, vp]
        L.gpe_batch_cand_merge.argtypes = [vp, u32, vp, vp, u64]
        L.gpe_batch_scan.argtypes = [vp]
        L.gpe_batch_bitmap.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
        L.gpe_batch_bitmap_merge.argtypes = [vp, u32, vp]
        L.gpe_pge_build.argtypes = [vp, u32, vp]
        L.gpe_host_pge_groups.argtypes = [u32, vp, vp, vp, u32, u32, vp, vp, vp]
        L.gpe_pge_dump_groups.argtypes = [vp, vp, vp, vp]
        L.gpe_pge_batch_upload.argtypes = [vp, C.POINTER(Batch)]
        L.gpe_pge_batch_filter.argtypes = [vp]
        L.gpe_pge_query_batch.argtypes = [vp, C.POINTER(Batch), vp]
        L.gpe_query_batches.argtypes = [vp, u32, vp, u32, vp]
        L.gpe_set_table_layout.argtypes = [vp, C.c_int]
        L.gpe_comm_unique_id.argtypes = [vp]
        L.gpe_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
        L.gpe_comm_init_all.argtypes = [vp, C.c_int]
        L.gpe_comm_destroy.argtypes = [vp]
        L.gpe_comm_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.gpe_build_table_shard.argtypes = [vp, C.POINTER(u64)]
        L.gpe_batch_step.argtypes = [vp]
        L.gpe_batch_finish.argtypes = [vp, vp]
        L.gpe_multi_batch_upload.argtypes = [vp, C.c_int, C.POINTER(Batch), u32]
        L.gpe_multi_batch_step.argtypes = [vp, C.c_int]
        L.gpe_multi_batch_finish.argtypes = [vp, C.c_int, vp]
        L.gpe_multi_query_batch.argtypes = [vp, C.c_int, C.POINTER(Batch), u32, vp]
        L.gpe_batch_get_candidates.argtypes = [vp, vp, vp]
        L.gpe_batch_get_plan.argtypes = [vp, vp, vp]
        L.gpe_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.gpe_stream.argtypes = [vp]
        L.gpe_stream.restype = vp
        L.gpe_sync.argtypes = [vp]
        L.gpe_set_timing.argtypes = [vp, C.c_int]
        L.gpe_collect_timings.argtypes = [vp, vp, vp]
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


# ---- host-side mirror (no GPU needed) -----------------------------------------------------------------------
def host_load_graph(path: str):
    V, E = C.c_uint32(0), C.c_uint32(0)
    rc = lib().gpe_host_load_graph(path.encode(), C.byref(V), C.byref(E), None, None, None)
    if rc:
        raise GpeError(f"gpe_host_load_graph({path}) failed with {rc}")
    off = np.zeros(V.value + 1, dtype=np.uint32)
    nbr = np.zeros(max(2 * E.value, 1), dtype=np.uint32)
    lab = np.zeros(max(V.value, 1), dtype=np.uint32)
    lib().gpe_host_load_graph(path.encode(), C.byref(V), C.byref(E), _ptr(off), _ptr(nbr), _ptr(lab))
    return off, nbr[: 2 * E.value], lab[: V.value]


def host_gen_vde(offsets, nbrs, labels, e: int):
    offsets, nbrs, labels = _u32(offsets), _u32(nbrs), _u32(labels)
    V = len(labels)
    x = np.zeros((max(V, 1), e), dtype=np.float64)
    vde = np.zeros((max(V, 1), e), dtype=np.float64)
    rc = lib().gpe_host_gen_vde(V, _ptr(offsets), _ptr(nbrs), _ptr(labels), e, _ptr(x), _ptr(vde))
    if rc:
        raise GpeError(f"gpe_host_gen_vde failed with {rc}")
    return x[:V], vde[:V]


def host_query_plan(q_offsets, q_nbrs, q_labels, L: int, e: int):
    q_offsets, q_nbrs, q_labels = _u32(q_offsets), _u32(q_nbrs), _u32(q_labels)
    nq = len(q_labels)
    cap = 4096
    vids = np.zeros((cap, L), dtype=np.uint32)
    labels = np.zeros((cap, L), dtype=np.uint32)
    degs = np.zeros((cap, L), dtype=np.uint32)
    pde = np.zeros((cap, L * e), dtype=np.float64)
    n = C.c_uint32(0)
    rc = lib().gpe_host_query_plan(nq, _ptr(q_offsets), _ptr(q_nbrs), _ptr(q_labels), L, e, cap, _ptr(vids),
                                   _ptr(labels), _ptr(degs), _ptr(pde), C.byref(n))
    if rc:
        raise GpeError(f"gpe_host_query_plan failed with {rc}")
    k = n.value
    return dict(vids=vids[:k], labels=labels[:k], degrees=degs[:k], pde=pde[:k])


def host_pge_groups(offsets, nbrs, labels, pl: int, e: int):
    """GNN-PGE path groups on the host: (pg, plg) as V x 2*pl*e [lo, hi, ...] and has[V]."""
    offsets, nbrs, labels = _u32(offsets), _u32(nbrs), _u32(labels)
    V = len(labels)
    pg = np.zeros((max(V, 1), 2 * pl * e), dtype=np.float64)
    plg = np.zeros((max(V, 1), 2 * pl * e), dtype=np.float64)
    has = np.zeros(max(V, 1), dtype=np.uint8)
    rc = lib().gpe_host_pge_groups(V, _ptr(offsets), _ptr(nbrs), _ptr(labels), pl, e, _ptr(pg), _ptr(plg), _ptr(has))
    if rc:
        raise GpeError(f"gpe_host_pge_groups failed with {rc}")
    return pg[:V], plg[:V], has[:V]


def pack_queries(queries):
    """Concatenate query graphs (objects with .offsets/.nbrs/.labels) into gpe_batch arrays."""
    vbase = np.zeros(len(queries) + 1, dtype=np.uint32)
    ebase = np.zeros(len(queries) + 1, dtype=np.uint32)
    offs, nbrs, labs = [], [], []
    for i, q in enumerate(queries):
        vbase[i + 1] = vbase[i] + len(q.labels)
        ebase[i + 1] = ebase[i] + len(q.nbrs)
        offs.append(_u32(q.offsets))
        nbrs.append(_u32(q.nbrs))
        labs.append(_u32(q.labels))
    cat = lambda xs: np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros(0, np.uint32), dtype=np.uint32)
    return vbase, ebase, cat(offs), cat(nbrs + [np.zeros(1, np.uint32)]), cat(labs)


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through libgpe (rank 0 calls it and hands the bytes to the other ranks)."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    if lib().gpe_comm_unique_id(buf):
        raise GpeError(lib().gpe_last_error(None).decode())
    return buf.raw


class MultiGpu:
    """One process, several GPUs: contexts in rank order sharing one NCCL communicator (gpe_comm_init_all)."""

    def __init__(self, devices):
        self.ctxs = [GpeContext(d) for d in devices]
        self._L = lib()
        self._arr = (C.c_void_p * len(self.ctxs))(*[c._h for c in self.ctxs])
        if len(self.ctxs) > 1 and self._L.gpe_comm_init_all(self._arr, len(self.ctxs)):
            raise GpeError(self._L.gpe_last_error(self.ctxs[0]._h).decode())

    def _ck(self, rc):
        if rc:
            msgs = [self._L.gpe_last_error(c._h).decode() for c in self.ctxs]
            raise GpeError("; ".join(m for m in msgs if m))

    def build(self, g, L, e, p, sorted_nodes, membership, vde):
        rows = []
        for c in self.ctxs:
            c.set_graph(g.offsets, g.nbrs, g.labels)
            c.set_embeddings(vde)
            n_rows, _ = c.enumerate(L, sorted_nodes, membership, p)
            rows.append(c.build_table_shard())
        return n_rows, rows

    def query_batch(self, queries, limits=None, flags: int = 0) -> np.ndarray:
        b = self.ctxs[0]._batch_struct(queries, limits)
        ans = np.zeros(max(len(queries), 1), dtype=np.uint64)
        self._ck(self._L.gpe_multi_query_batch(self._arr, len(self.ctxs), C.byref(b), flags, _ptr(ans)))
        return ans[: len(queries)]

    def close(self):
        for c in self.ctxs:
            c.close()


class GpeContext:
    """One GPU, one stream.  Mirrors the reference's offline/online split:
    set_graph -> set_embeddings -> enumerate (offline) -> build_table -> filter/refine/query_batch (online)."""

    def __init__(self, device: int = 0):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.gpe_create(device, C.byref(h))
        if rc:
            raise GpeError(f"gpe_create({device}) failed: {self._L.gpe_last_error(None).decode()}")
        self._h = h
        self.V = 0
        self.L = 0
        self.e = 0
        self.n_rows = 0
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.gpe_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise GpeError(self._L.gpe_last_error(self._h).decode())

    # ---- data ----
    def set_graph(self, offsets, nbrs, labels):
        offsets, nbrs, labels = _u32(offsets), _u32(nbrs), _u32(labels)
        self.V = len(labels)
        self._ck(self._L.gpe_set_graph(self._h, self.V, _ptr(offsets), _ptr(nbrs), _ptr(labels)))

    def set_embeddings(self, vde):
        vde = np.ascontiguousarray(vde, dtype=np.float64)
        self.e = vde.shape[1]
        self._ck(self._L.gpe_set_embeddings(self._h, self.e, _ptr(vde)))

    # ---- S1 ----
    def enumerate(self, L: int, sorted_nodes, membership, p: int):
        sorted_nodes, membership = _u32(sorted_nodes), _u32(membership)
        rows = np.zeros(p, dtype=np.uint64)
        n = C.c_uint64(0)
        self._ck(self._L.gpe_enumerate(self._h, L, _ptr(sorted_nodes), _ptr(membership), p, _ptr(rows), C.byref(n)))
        self.L, self.n_rows = L, int(n.value)
        return self.n_rows, rows

    def dump_paths(self, first: int = 0, n: int | None = None) -> np.ndarray:
        n = self.n_rows - first if n is None else n
        out = np.zeros((max(n, 1), self.L), dtype=np.uint32)
        self._ck(self._L.gpe_dump_paths(self._h, first, n, _ptr(out)))
        return out[:n]

    def start_rows(self) -> np.ndarray:
        out = np.zeros(self.V + 1, dtype=np.uint64)
        self._ck(self._L.gpe_start_rows(self._h, _ptr(out)))
        return out

    # ---- S2 ----
    def set_table_layout(self, layout: int):
        """0 auto, 1 materialised rows, 2 vertex ids only (rows gathered by the scan)."""
        self._ck(self._L.gpe_set_table_layout(self._h, layout))

    def build_table(self, part_select=None) -> int:
        sel = None if part_select is None else np.ascontiguousarray(part_select, dtype=np.uint8)
        n = C.c_uint64(0)
        self._ck(self._L.gpe_build_table(self._h, _ptr(sel), C.byref(n)))
        self.table_rows = int(n.value)
        return self.table_rows

    def dump_table(self, first: int = 0, n: int | None = None):
        n = self.table_rows - first if n is None else n
        D = self.L * self.e
        vids = np.zeros((max(n, 1), self.L), dtype=np.uint32)
        labels = np.zeros((max(n, 1), self.L), dtype=np.uint32)
        degs = np.zeros((max(n, 1), self.L), dtype=np.uint32)
        pde = np.zeros((max(n, 1), D), dtype=np.float64)
        self._ck(self._L.gpe_dump_table(self._h, first, n, _ptr(vids), _ptr(labels), _ptr(degs), _ptr(pde)))
        return vids[:n], labels[:n], degs[:n], pde[:n]

    def filter(self, plan: dict, nq: int, flags: int = 0):
        """plan: dict(vids, labels, degrees, pde) as host_query_plan returns.  Returns (candidate
        lists per query vertex, survivors per plan path)."""
        vids, labels, degs = _u32(plan["vids"]), _u32(plan["labels"]), _u32(plan["degrees"])
        pde = np.ascontiguousarray(plan["pde"], dtype=np.float64)
        n = len(vids)
        off = np.zeros(nq + 1, dtype=np.uint64)
        surv = np.zeros(max(n, 1), dtype=np.uint64)
        self._ck(self._L.gpe_filter(self._h, n, _ptr(vids), _ptr(labels), _ptr(degs), _ptr(pde), nq, flags,
                                    _ptr(off), _ptr(surv)))
        cand = np.zeros(max(int(off[nq]), 1), dtype=np.uint32)
        self._ck(self._L.gpe_get_candidates(self._h, _ptr(cand)))
        return [cand[int(off[u]): int(off[u + 1])].copy() for u in range(nq)], surv[:n]

    # ---- S3 ----
    def refine(self, q_offsets, q_nbrs, q_labels, cand_sets, limit: int = LIMIT_MAX, want_matches: int = 0):
        q_offsets, q_nbrs, q_labels = _u32(q_offsets), _u32(q_nbrs), _u32(q_labels)
        nq = len(q_labels)
        off = np.zeros(nq + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(c) for c in cand_sets])
        flat = _u32(np.concatenate([np.asarray(c, dtype=np.uint32) for c in cand_sets] + [np.zeros(1, np.uint32)]))
        n = C.c_uint64(0)
        order = np.zeros(nq, dtype=np.uint32)
        pivot = np.zeros(nq, dtype=np.uint32)
        matches = np.zeros((max(want_matches, 1), nq), dtype=np.uint32) if want_matches else None
        self._ck(self._L.gpe_refine(self._h, nq, _ptr(q_offsets), _ptr(q_nbrs if len(q_nbrs) else np.zeros(1, np.uint32)),
                                    _ptr(q_labels), _ptr(off), _ptr(flat), limit, C.byref(n), _ptr(order), _ptr(pivot),
                                    _ptr(matches), want_matches))
        res = dict(n_matches=int(n.value), order=order, pivot=pivot)
        if want_matches:
            res["matches"] = matches[: min(int(n.value), want_matches)]
        return res

    # ---- batch ----
    def _batch_struct(self, queries, limits=None):
        vbase, ebase, offs, nbrs, labs = pack_queries(queries)
        lim = None if limits is None else np.ascontiguousarray(limits, dtype=np.uint64)
        b = Batch(len(queries), vbase.ctypes.data, ebase.ctypes.data, offs.ctypes.data, nbrs.ctypes.data,
                  labs.ctypes.data, None if lim is None else lim.ctypes.data)
        self._keep = (vbase, ebase, offs, nbrs, labs, lim)  # keep the host arrays alive
        return b

    def query_batch(self, queries, limits=None, flags: int = 0) -> np.ndarray:
        """End-to-end: host query graphs in, answer counts out (one call, H2D and D2H inside)."""
        b = self._batch_struct(queries, limits)
        ans = np.zeros(max(len(queries), 1), dtype=np.uint64)
        self._ck(self._L.gpe_query_batch(self._h, C.byref(b), flags, _ptr(ans)))
        return ans[: len(queries)]

    def prepare_batches(self, batches, limits=None):
        """Pack several batches (lists of query graphs) into gpe_batch structs: the host buffers gpe_query_batches takes."""
        n = len(batches)
        structs = (Batch * max(n, 1))()
        keep, outs = [], []
        for i, qs in enumerate(batches):
            structs[i] = self._batch_struct(qs, None if limits is None else limits[i])
            keep.append(self._keep)
            outs.append(np.zeros(max(len(qs), 1), dtype=np.uint64))
        ptrs = (C.c_void_p * max(n, 1))(*[o.ctypes.data for o in outs])
        return dict(n=n, structs=structs, keep=keep, outs=outs, ptrs=ptrs, sizes=[len(qs) for qs in batches])

    def run_batches(self, prepared, flags: int = 0):
        """gpe_query_batches: all the batches in ONE call, host planning of batch i+1 overlapped with the GPU work of batch i."""
        self._ck(self._L.gpe_query_batches(self._h, prepared["n"], prepared["structs"], flags, prepared["ptrs"]))
        self._n_queries = prepared["sizes"][-1] if prepared["n"] else 0
        return [o[:k] for o, k in zip(prepared["outs"], prepared["sizes"])]

    def query_batches(self, batches, limits=None, flags: int = 0):
        return self.run_batches(self.prepare_batches(batches, limits), flags)

    def batch_upload(self, queries, limits=None, flags: int = 0):
        b = self._batch_struct(queries, limits)
        self._n_queries = len(queries)
        self._ck(self._L.gpe_batch_upload(self._h, C.byref(b), flags))

    def batch_filter(self):
        self._ck(self._L.gpe_batch_filter(self._h))

    def batch_join(self, rank: int = 0, world: int = 1):
        self._ck(self._L.gpe_batch_join(self._h, rank, world))

    def batch_download(self) -> np.ndarray:
        out = np.zeros(max(self._n_queries, 1), dtype=np.uint64)
        self._ck(self._L.gpe_batch_download(self._h, _ptr(out)))
        return out[: self._n_queries]

    # ---- multi-GPU with NCCL inside the library (one process per GPU) ----
    def comm_init(self, rank: int, world: int, unique_id: bytes):
        buf = C.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        self._ck(self._L.gpe_comm_init(self._h, rank, world, buf))

    def comm_info(self):
        r, w, v = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self._L.gpe_comm_info(self._h, C.byref(r), C.byref(w), C.byref(v)))
        return r.value, w.value, v.value

    def build_table_shard(self) -> int:
        n = C.c_uint64(0)
        self._ck(self._L.gpe_build_table_shard(self._h, C.byref(n)))
        return int(n.value)

    def batch_step(self):
        self._ck(self._L.gpe_batch_step(self._h))

    def batch_finish(self) -> np.ndarray:
        out = np.zeros(max(self._n_queries, 1), dtype=np.uint64)
        self._ck(self._L.gpe_batch_finish(self._h, _ptr(out)))
        return out[: self._n_queries]

    def batch_cand_info(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._ck(self._L.gpe_batch_cand_info(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def batch_cand_export(self, d_counts_ptr: int, d_cand_ptr: int):
        self._ck(self._L.gpe_batch_cand_export(self._h, d_counts_ptr, d_cand_ptr))

    def batch_cand_merge(self, world: int, d_counts_ptr: int, d_cand_ptr: int, stride: int):
        self._ck(self._L.gpe_batch_cand_merge(self._h, world, d_counts_ptr, d_cand_ptr, stride))

    def batch_scan(self):
        self._ck(self._L.gpe_batch_scan(self._h))

    def batch_bitmap(self):
        """(device pointer, bytes) of the batch's candidate bitmaps."""
        p, n = C.c_void_p(), C.c_uint64(0)
        self._ck(self._L.gpe_batch_bitmap(self._h, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def batch_bitmap_merge(self, world: int, d_all_ptr: int):
        self._ck(self._L.gpe_batch_bitmap_merge(self._h, world, d_all_ptr))

    def batch_get_candidates(self):
        n_slots, total = self.batch_cand_info()
        off = np.zeros(n_slots + 1, dtype=np.uint64)
        cand = np.zeros(max(total, 1), dtype=np.uint32)
        self._ck(self._L.gpe_batch_get_candidates(self._h, _ptr(off), _ptr(cand)))
        return off, cand[:total]

    def batch_get_plan(self, n_slots: int):
        order = np.zeros(max(n_slots, 1), dtype=np.uint32)
        pivot = np.zeros(max(n_slots, 1), dtype=np.uint32)
        self._ck(self._L.gpe_batch_get_plan(self._h, _ptr(order), _ptr(pivot)))
        return order[:n_slots], pivot[:n_slots]

    # ---- GNN-PGE variant (see include/gpe.h for its status) ----
    def pge_build(self, pl: int, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.pge_pl = pl
        self._ck(self._L.gpe_pge_build(self._h, pl, _ptr(x)))

    def pge_dump_groups(self):
        pde = self.pge_pl * self.e
        pg = np.zeros((max(self.V, 1), 2 * pde), dtype=np.float64)
        plg = np.zeros((max(self.V, 1), 2 * pde), dtype=np.float64)
        has = np.zeros(max(self.V, 1), dtype=np.uint8)
        self._ck(self._L.gpe_pge_dump_groups(self._h, _ptr(pg), _ptr(plg), _ptr(has)))
        return pg[: self.V], plg[: self.V], has[: self.V]

    def pge_batch_upload(self, queries, limits=None):
        b = self._batch_struct(queries, limits)
        self._n_queries = len(queries)
        self._ck(self._L.gpe_pge_batch_upload(self._h, C.byref(b)))

    def pge_batch_filter(self):
        self._ck(self._L.gpe_pge_batch_filter(self._h))

    def pge_query_batch(self, queries, limits=None) -> np.ndarray:
        b = self._batch_struct(queries, limits)
        ans = np.zeros(max(len(queries), 1), dtype=np.uint64)
        self._ck(self._L.gpe_pge_query_batch(self._h, C.byref(b), _ptr(ans)))
        return ans[: len(queries)]

    def clamp(self, raw: int, limit: int) -> int:
        return int(self._L.gpe_clamp_answer(int(raw), int(limit)))

    # ---- measurement ----
    def stats(self) -> dict:
        s = Stats()
        self._ck(self._L.gpe_get_stats(self._h, C.byref(s)))
        return s.asdict()

    def set_timing(self, mode: int):
        """0 off, 1 synchronous per stage (stats()['last_*_ms']), 2 deferred (collect_timings())."""
        self._ck(self._L.gpe_set_timing(self._h, int(mode)))

    def collect_timings(self) -> dict:
        ms = np.zeros(5, dtype=np.float64)
        cnt = np.zeros(5, dtype=np.uint64)
        self._ck(self._L.gpe_collect_timings(self._h, _ptr(ms), _ptr(cnt)))
        names = ["select", "scan", "compact", "join", "enumerate"]
        return {n: dict(ms=float(ms[i]), launches=int(cnt[i])) for i, n in enumerate(names)}

    def sync(self):
        self._ck(self._L.gpe_sync(self._h))

    @property
    def stream(self) -> int:
        return int(self._L.gpe_stream(self._h) or 0)
