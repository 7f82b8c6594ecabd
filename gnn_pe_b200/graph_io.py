"""`.graph` text format and CSR helpers (numpy, host side only).

Format (reference: libsrc/graph/graph.cpp:163-242):
``t V E`` / ``v id label degree`` for ids 0..V-1 in order / ``e u v`` once per undirected edge.
Adjacency lists are sorted ascending after load (graph.cpp:231-233).
"""
from __future__ import annotations

import dataclasses
import numpy as np


@dataclasses.dataclass
class CSRGraph:
    """Undirected labelled simple graph in CSR form; ``nbrs`` sorted ascending per vertex."""
    offsets: np.ndarray  # uint32 [V+1]
    nbrs: np.ndarray     # uint32 [2E]
    labels: np.ndarray   # uint32 [V]

    @property
    def V(self) -> int:
        return int(self.labels.shape[0])

    @property
    def E(self) -> int:
        return int(self.nbrs.shape[0] // 2)

    @property
    def degrees(self) -> np.ndarray:
        return np.diff(self.offsets.astype(np.int64)).astype(np.uint32)

    @property
    def labels_count(self) -> int:
        """graph.cpp:223: max(#distinct labels, max label + 1)."""
        if self.V == 0:
            return 0
        return int(max(len(np.unique(self.labels)), int(self.labels.max()) + 1))

    def edge_list(self) -> np.ndarray:
        """Each undirected edge once as (u, v) with u < v, sorted."""
        src = np.repeat(np.arange(self.V, dtype=np.uint32), self.degrees.astype(np.int64))
        keep = src < self.nbrs
        return np.stack([src[keep], self.nbrs[keep]], axis=1)


def csr_from_edges(V: int, edges: np.ndarray, labels: np.ndarray) -> CSRGraph:
    """Build a CSR from an (E, 2) array of undirected edges (each once, no loops, no duplicates)."""
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    if edges.size:
        if (edges[:, 0] == edges[:, 1]).any():
            raise ValueError("self loops are not supported (the reference's path DFS ignores them)")
        lo = np.minimum(edges[:, 0], edges[:, 1])
        hi = np.maximum(edges[:, 0], edges[:, 1])
        key = lo * V + hi
        if len(np.unique(key)) != len(key):
            raise ValueError("duplicate edges are not supported (simple graphs only)")
    src = np.concatenate([edges[:, 0], edges[:, 1]])
    dst = np.concatenate([edges[:, 1], edges[:, 0]])
    order = np.lexsort((dst, src))
    src, dst = src[order], dst[order]
    offsets = np.zeros(V + 1, dtype=np.int64)
    if src.size:
        offsets[1:] = np.bincount(src, minlength=V)
    offsets = np.cumsum(offsets)
    if offsets[-1] >= 1 << 32:
        raise ValueError("2E must fit in 32 bits (reference uses ui offsets, graph.h:61)")
    return CSRGraph(offsets.astype(np.uint32), dst.astype(np.uint32), np.asarray(labels, dtype=np.uint32))


def read_graph(path: str) -> CSRGraph:
    V = E = 0
    labels = None
    edges = []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "t":
                V, E = int(t[1]), int(t[2])
                labels = np.zeros(V, dtype=np.uint32)
            elif t[0] == "v":
                labels[int(t[1])] = int(t[2])
            elif t[0] == "e":
                edges.append((int(t[1]), int(t[2])))
    return csr_from_edges(V, np.array(edges, dtype=np.int64).reshape(-1, 2), labels)


def write_graph(path: str, g: CSRGraph) -> None:
    deg = g.degrees
    el = g.edge_list()
    with open(path, "w") as f:
        f.write(f"t {g.V} {len(el)}\n")
        for v in range(g.V):
            f.write(f"v {v} {int(g.labels[v])} {int(deg[v])}\n")
        for u, v in el:
            f.write(f"e {int(u)} {int(v)}\n")


def degree_order(g: CSRGraph) -> np.ndarray:
    """Stable ascending-degree vertex order = line order of membership.txt (gnnpe.py:71-76)."""
    return np.argsort(g.degrees, kind="stable").astype(np.uint32)


def write_membership(path: str, sorted_nodes: np.ndarray, membership: np.ndarray) -> None:
    with open(path, "w") as f:
        for v in sorted_nodes:
            f.write(f"{int(v)} {int(membership[int(v)])}\n")


def read_membership(path: str, V: int):
    data = np.loadtxt(path, dtype=np.int64).reshape(-1, 2)
    sorted_nodes = data[:, 0].astype(np.uint32)
    membership = np.zeros(V, dtype=np.uint32)
    membership[data[:, 0]] = data[:, 1]
    return sorted_nodes, membership


def block_membership(V: int, p: int) -> np.ndarray:
    """Stand-in for the METIS assignment (pymetis is absent): contiguous id blocks."""
    return (np.arange(V, dtype=np.int64) * p // max(V, 1)).astype(np.uint32)
