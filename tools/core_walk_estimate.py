"""Analysis only (CPU, numpy/scipy; not product code).  Round-2 groundwork: for config 2's queries with exactly ONE repeated label pair, compare
  walk  = partial label-paths a depth-first walk of the core path enumerates (both orientations, the cheaper one), with
  mitm  = half-length label-walks from both same-label endpoints (what meeting in the middle would enumerate to count
          the CLOSED walks that inclusion-exclusion subtracts from the table product).
Degree filters and subtree-table pruning are ignored on both sides (upper bounds on both)."""
import sys, pickle, re
import numpy as np, scipy.sparse as sp
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from collections import Counter

w, g, queries = bench.load_workload("config2")
V = g.V
lab = g.labels.astype(np.int64)
deg = np.diff(g.offsets.astype(np.int64))
rows = np.repeat(np.arange(V), deg); cols = g.nbrs.astype(np.int64)
A = sp.csr_matrix((np.ones(len(cols), dtype=np.float64), (rows, cols)), shape=(V, V))
ind = [sp.diags((lab == l).astype(np.float64)) for l in range(20)]

def walks(labels_seq):
    """number of label-constrained walks of every prefix length starting from all vertices of labels_seq[0]"""
    v = (lab == labels_seq[0]).astype(np.float64)
    out = []
    for l in labels_seq[1:]:
        v = (A.T @ v) * (lab == l)
        out.append(float(v.sum()))
    return out

def tree_path(q, a, b):
    off, nbr = q.offsets, q.nbrs
    prev = {a: None}; st = [a]
    while st:
        u = st.pop()
        for v in nbr[off[u]:off[u+1]]:
            v = int(v)
            if v not in prev: prev[v] = u; st.append(v)
    p = [b]
    while prev[p[-1]] is not None: p.append(prev[p[-1]])
    return p[::-1]

steps = {}
for line in open(os.path.join(ROOT, 'profiles', 'r01j_joinstats_config2.log')):
    m = re.match(r'q(\d+) nq=(\d+) ne=(\d+): .*steps=(\d+) ', line)
    if m: steps[int(m[1])] = int(m[4])
tot_walk = tot_mitm = tot_meas = 0
print("query dist measured_steps walk_est mitm_est")
for i, q in enumerate(queries):
    c = Counter(map(int, q.labels))
    rep = [l for l, n in c.items() if n > 1]
    if len(rep) != 1 or c[rep[0]] != 2 or len(q.nbrs)//2 != q.V - 1: continue
    a, b = [u for u in range(q.V) if int(q.labels[u]) == rep[0]]
    p = tree_path(q, a, b)
    seq = [int(q.labels[u]) for u in p]
    k = len(p) - 1  # edges between the pair
    # depth-first walk of the inner path (the two endpoints are counted leaves): start at an inner end, k-2 steps
    inner = seq[1:-1]
    if len(inner) >= 1:
        w1 = sum(walks(inner)) + 50000 if len(inner) > 1 else 50000
        w2 = sum(walks(inner[::-1])) + 50000 if len(inner) > 1 else 50000
        walk = min(w1, w2)
    else:
        walk = 0
    # closed walks through a1 = a2 = x: forward ceil(k/2) steps along seq, backward floor(k/2) steps along reversed seq
    h1, h2 = (k + 1) // 2, k // 2
    f = walks(seq[:h1 + 1]); bwd = walks(seq[::-1][:h2 + 1])
    mitm = sum(f) + sum(bwd)
    tot_walk += walk; tot_mitm += mitm; tot_meas += steps.get(i, 0)
    print(i, k, steps.get(i), int(walk), int(mitm))
print("totals: measured", tot_meas, "walk_est", int(tot_walk), "mitm_est", int(tot_mitm))
