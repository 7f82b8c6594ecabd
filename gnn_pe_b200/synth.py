"""Seeded synthetic inputs of the shapes BASELINE.json names.

The reference ships no generator; these write/return graphs in its formats.
"""
from __future__ import annotations

import numpy as np

from .graph_io import CSRGraph, csr_from_edges


def _dedup_edges(V: int, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    keep = u != v
    u, v = u[keep], v[keep]
    lo = np.minimum(u, v).astype(np.int64)
    hi = np.maximum(u, v).astype(np.int64)
    key = np.unique(lo * V + hi)
    return key


def chung_lu_graph(V: int, E: int, n_labels: int, gamma: float = 3.0, degree_cap: int = 1000,
                   seed: int = 2022) -> CSRGraph:
    """Power-law (Chung-Lu) simple graph with exactly E edges and uniform labels.

    Expected degree of vertex i is proportional to (i + i0)^(-1/(gamma-1)), capped at
    ``degree_cap``.  SURVEY.md section 8(d): gamma=3.0 with cap 1000 keeps the l=2 table of
    the 1M/10M config at ~0.5 B rows so it fits one B200.
    """
    rng = np.random.default_rng(seed)
    mean_deg = 2.0 * E / V
    w = (np.arange(V, dtype=np.float64) + 1.0) ** (-1.0 / (gamma - 1.0))
    for _ in range(8):  # rescale under the cap
        w *= mean_deg / w.mean()
        w = np.minimum(w, float(degree_cap))
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    keys = np.zeros(0, dtype=np.int64)
    while len(keys) < E:
        need = int((E - len(keys)) * 1.1) + 16
        u = np.searchsorted(cdf, rng.random(need)).astype(np.int64)
        v = np.searchsorted(cdf, rng.random(need)).astype(np.int64)
        keys = np.union1d(keys, _dedup_edges(V, u, v))
    if len(keys) > E:
        keys = rng.permutation(keys)[:E]
    edges = np.stack([keys // V, keys % V], axis=1)
    perm = rng.permutation(V)  # so that vertex id carries no degree information
    edges = perm[edges]
    labels = rng.integers(0, n_labels, size=V, dtype=np.int64)
    return csr_from_edges(V, edges, labels)


def uniform_graph(V: int, E: int, n_labels: int, seed: int = 2022) -> CSRGraph:
    """Erdos-Renyi style simple graph (Poisson degrees) with exactly E edges."""
    rng = np.random.default_rng(seed)
    keys = np.zeros(0, dtype=np.int64)
    while len(keys) < E:
        need = int((E - len(keys)) * 1.1) + 16
        u = rng.integers(0, V, size=need, dtype=np.int64)
        v = rng.integers(0, V, size=need, dtype=np.int64)
        keys = np.union1d(keys, _dedup_edges(V, u, v))
    if len(keys) > E:
        keys = rng.permutation(keys)[:E]
    edges = np.stack([keys // V, keys % V], axis=1)
    labels = rng.integers(0, n_labels, size=V, dtype=np.int64)
    return csr_from_edges(V, edges, labels)


def table_rows_l2(g: CSRGraph) -> int:
    """Row count of the l=2 path table: sum over vertices of C(deg, 2)."""
    d = g.degrees.astype(np.int64)
    return int((d * (d - 1) // 2).sum())


def random_walk_query(g: CSRGraph, n_vertices: int, rng: np.random.Generator, induced: bool = True,
                      max_steps: int = 100000) -> CSRGraph:
    """Query graph from a random walk on the data graph.

    Walk until ``n_vertices`` distinct vertices are seen; vertices are renumbered in order of
    discovery.  ``induced=True`` keeps every data edge among them (dense query), otherwise
    only the walk's tree edges (sparse query).  Labels are the data labels; degrees are those
    inside the query graph.  Connected by construction.
    """
    off, nbr = g.offsets, g.nbrs
    while True:
        cur = int(rng.integers(0, g.V))
        if off[cur + 1] - off[cur] == 0:
            continue
        seen = {cur: 0}
        tree = []
        for _ in range(max_steps):
            if len(seen) == n_vertices:
                break
            d = int(off[cur + 1] - off[cur])
            nxt = int(nbr[int(off[cur]) + int(rng.integers(0, d))])
            if nxt not in seen:
                seen[nxt] = len(seen)
                tree.append((seen[cur], seen[nxt]))
            cur = nxt
        if len(seen) == n_vertices:
            break
    verts = np.array(sorted(seen, key=seen.get), dtype=np.int64)
    if induced:
        edges = []
        for a in verts:
            row = nbr[int(off[a]):int(off[a + 1])]
            for b in verts:
                if a < b and np.searchsorted(row, b) < len(row) and row[np.searchsorted(row, b)] == b:
                    edges.append((seen[int(a)], seen[int(b)]))
    else:
        edges = tree
    return csr_from_edges(n_vertices, np.array(edges, dtype=np.int64).reshape(-1, 2), g.labels[verts])


def query_batch(g: CSRGraph, n_queries: int, n_vertices, seed: int = 2023, mixed: bool = False):
    """``n_queries`` random-walk queries.  ``n_vertices`` is an int or an inclusive (lo, hi) range;
    ``mixed`` alternates induced (dense) and tree (sparse) queries (config 5)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_queries):
        nv = n_vertices if isinstance(n_vertices, int) else int(rng.integers(n_vertices[0], n_vertices[1] + 1))
        out.append(random_walk_query(g, nv, rng, induced=(not mixed) or (i % 2 == 0)))
    return out
