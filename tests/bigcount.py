"""Exact (arbitrary precision) embedding counts for small data graphs -- test infrastructure.

The reference's refinement (custom.h:757-888) counts one embedding at a time, so neither it nor the oracle
can reach counts beyond ~1e9 in a test.  What it counts is well defined though: injective maps f of the query
vertices with f(start) in C(start) (taken as it is, :827-830) and, for every other query vertex u,
label(f(u)) == label(u), deg(f(u)) >= deg(u) (:757-797), every query edge on a data edge.  This module counts
exactly that with Python integers: pendant subtrees whose labels are unique in the query are folded into
per-vertex weights by a tree DP (they cannot collide with anything), the remaining core is enumerated.
"""
from __future__ import annotations

import numpy as np


def exact_count(g, q, start: int, cand_start) -> int:
    """g, q: graph_io.CSRGraph; start: the reference's start vertex (order[0]); cand_start: C(start)."""
    goff, gn, gl = g.offsets.astype(np.int64), g.nbrs, g.labels
    gdeg = np.diff(goff)
    qoff, qn, ql = q.offsets.astype(np.int64), q.nbrs, q.labels
    nq = q.V
    qadj = [list(map(int, qn[qoff[u]:qoff[u + 1]])) for u in range(nq)]
    qdeg = [len(a) for a in qadj]
    nbrs_of = lambda x: gn[goff[x]:goff[x + 1]]
    uniq = [sum(1 for v in range(nq) if ql[v] == ql[u]) == 1 for u in range(nq)]
    alive = [True] * nq
    rem = qdeg[:]
    children = [[] for _ in range(nq)]
    changed = True
    while changed:
        changed = False
        for u in range(nq):
            if alive[u] and rem[u] == 1 and uniq[u] and u != start and sum(alive) > 1:
                p = next(v for v in qadj[u] if alive[v])
                alive[u] = False
                rem[p] -= 1
                children[p].append(u)
                changed = True
    memo = {}

    def weight(u, x):  # ways to map everything peeled below u, given u -> x
        key = (u, x)
        if key in memo:
            return memo[key]
        w = 1
        for c in children[u]:
            s = 0
            for y in nbrs_of(x):
                y = int(y)
                if gl[y] == ql[c] and gdeg[y] >= qdeg[c]:
                    s += weight(c, y)
            w *= s
            if w == 0:
                break
        memo[key] = w
        return w

    core = [u for u in range(nq) if alive[u]]
    order = [start]
    while len(order) < len(core):
        for u in core:
            if u not in order and any(v in order for v in qadj[u]):
                order.append(u)
                break
    pos = {u: i for i, u in enumerate(order)}
    total = 0
    emb = {}

    def rec(i, prod):
        nonlocal total
        if i == len(order):
            total += prod
            return
        u = order[i]
        back = [v for v in qadj[u] if alive[v] and pos[v] < i]
        pool = nbrs_of(emb[back[0]]) if back else []
        for y in pool:
            y = int(y)
            if gl[y] != ql[u] or gdeg[y] < qdeg[u] or y in emb.values():
                continue
            if any(not _edge(goff, gn, emb[b], y) for b in back[1:]):
                continue
            w = weight(u, y)
            if w:
                emb[u] = y
                rec(i + 1, prod * w)
                del emb[u]

    for x in cand_start:
        x = int(x)
        w = weight(start, x)
        if w:
            emb[start] = x
            rec(1, w)
            del emb[start]
    return total


def _edge(goff, gn, a, b):
    row = gn[goff[a]:goff[a + 1]]
    i = int(np.searchsorted(row, b))
    return i < len(row) and int(row[i]) == b


def reference_answer(g, q, l: int, e: int, limit: int) -> int:
    """min(exact count, limit) as the reference reports it (custom.h:846-855), with the candidate set of the start
    vertex taken from the oracle's filter (the reference's one-orientation quirk shapes it)."""
    from gnn_pe_b200 import graph_io
    from oracle import oracle
    og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
    oq = oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels)
    og.enumerate(l + 1, graph_io.degree_order(g))
    sets, _ = oracle.filter_candidates(og, oq, e)
    order, _ = oracle.matching_order(og, oq, [len(s) for s in sets])
    total = exact_count(g, q, int(order[0]), sets[int(order[0])])
    return min(total, max(int(limit), 1)), total
