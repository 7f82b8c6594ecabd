"""Small data/query graphs whose embedding counts exceed 64 bits (or whose factorised count has 64-bit
intermediates while the answer is small).  Shared by the CPU check of tests/bigcount.py and the GPU tests."""
import numpy as np

from gnn_pe_b200 import graph_io


def _graph(labels, edges):
    return graph_io.csr_from_edges(len(labels), np.array(sorted(set(edges)), dtype=np.int64).reshape(-1, 2),
                                   np.array(labels, dtype=np.uint32))


def _aligned_query(g, qlabels, qedges, image):
    """Number the query vertices by the data rank (degree, id) of their intended images.  The reference stores one
    orientation per path -- the one that starts at the endpoint enumerated first, i.e. of lower (degree, id) in the data
    graph and of lower id in the query (custom.h:68-79, :94-119) -- and never compares the reverse, so a
    query numbered any other way loses candidates to that quirk and the answer collapses to 0."""
    deg = g.degrees
    order = sorted(range(len(qlabels)), key=lambda u: (int(deg[image[u]]), image[u]))
    new = {u: i for i, u in enumerate(order)}
    return _graph([qlabels[u] for u in order], [(min(new[a], new[b]), max(new[a], new[b])) for a, b in qedges])


def star(n_leaf_labels=12, per_label=64):
    """Hub of label 0 with `per_label` neighbours of each of `n_leaf_labels` labels; the query is the star with one
    leaf per label: per_label ** n_leaf_labels embeddings (64^12 = 2^72 == 0 mod 2^64)."""
    labels, edges = [0], []
    for l in range(1, n_leaf_labels + 1):
        for _ in range(per_label):
            labels.append(l)
            edges.append((0, len(labels) - 1))
    g = _graph(labels, edges)
    image = [0] + [1 + i * per_label for i in range(n_leaf_labels)]
    q = _aligned_query(g, [0] + list(range(1, n_leaf_labels + 1)), [(0, i) for i in range(1, n_leaf_labels + 1)], image)
    return g, q


def weighted_pair(children=4, per_label=256, extra_b=True):
    """Core p1 - m - p2 with two same-label (X) leaves u1 on p1, u2 on p2, each carrying `children` pendant vertices
    of query-unique labels.  Data: x1 adjacent to p1 and p2, x2 to p1 only, x3 (extra_b) to p2 only; every x is adjacent
    to `per_label` vertices of every child label.  Pairs (u1, u2) of distinct x weigh (per_label^children)^2 each:
    2^64 per pair at 4 x 256."""
    P1, M, P2, X = 0, 1, 2, 3
    labels = [P1, M, P2, X, X] + ([X] if extra_b else [])
    p1, m, p2, x1, x2 = 0, 1, 2, 3, 4
    xs = [x1, x2] + ([5] if extra_b else [])
    edges = [(p1, m), (m, p2), (p1, x1), (p2, x1), (p1, x2)]
    if extra_b:
        edges.append((p2, 5))
    first_of = []
    for cl in range(2 * children):
        first_of.append(len(labels))
        for _ in range(per_label):
            labels.append(4 + cl)
            v = len(labels) - 1
            edges += [(x, v) for x in xs]
    # query: 0 p1, 1 m, 2 p2, 3 u1, 4 u2, then u1's children (labels 4..), u2's children
    ql = [P1, M, P2, X, X] + [4 + i for i in range(2 * children)]
    qe = [(0, 1), (1, 2), (0, 3), (2, 4)]
    qe += [(3, 5 + i) for i in range(children)] + [(4, 5 + children + i) for i in range(children)]
    g = _graph(labels, edges)
    return g, _aligned_query(g, ql, qe, [p1, m, p2, x2, xs[-1] if extra_b else x1] + first_of)


def saturated_minuend(children=11, per_label=64, light=1):
    """Triangle p - t - s (labels 0, X, 2) plus a leaf u of label X on p that carries `children` pendant vertices of
    query-unique labels.  Data: x1 (adjacent to p0, s0 and per_label vertices of every child label) and x2 (adjacent to
    p0 and `light` vertices of every child label).  t can only be x1, so u must be x2: light^children embeddings, but
    the table sum S_u[p0] = per_label^children + light^children is beyond 2^62 at 11 x 64."""
    P, X, S = 0, 1, 2
    labels = [P, S, X, X]
    p0, s0, x1, x2 = 0, 1, 2, 3
    edges = [(p0, x1), (p0, x2), (p0, s0), (s0, x1)]
    first_of = []
    for cl in range(children):
        first = len(labels)
        first_of.append(first)
        for _ in range(per_label):
            labels.append(3 + cl)
            edges.append((x1, len(labels) - 1))
        edges += [(x2, first + i) for i in range(light)]
    ql = [P, X, S, X] + [3 + i for i in range(children)]
    qe = [(0, 1), (1, 2), (0, 2), (0, 3)] + [(3, 4 + i) for i in range(children)]
    g = _graph(labels, edges)
    return g, _aligned_query(g, ql, qe, [p0, x1, s0, x2] + first_of)


CASES = {
    "star_2^72": lambda: star(12, 64),
    "star_small": lambda: star(5, 3),
    "pair_3x2^64": lambda: weighted_pair(4, 256, True),
    "pair_1x2^64": lambda: weighted_pair(4, 256, False),
    "pair_small": lambda: weighted_pair(2, 3, True),
    "minuend_saturated": lambda: saturated_minuend(11, 64, 1),
    "minuend_saturated_3": lambda: saturated_minuend(11, 64, 3),
    "minuend_small": lambda: saturated_minuend(3, 4, 2),
}
