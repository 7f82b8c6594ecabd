"""gnn-pge/data_vertices.bin (GNN-PGE/src/main.cpp:179-194) rebuilt on the CPU from the host mirror -- label embeddings,
their neighbour sums, dominance embeddings, path groups -- against the md5 of the file the unmodified GNN-PGE binary wrote
(tests/golden/<case>/golden_pge.json, `bin_md5`).  Pins the record layout `host/main --filter pge -m offline` writes."""
import hashlib
import json
import os

import numpy as np
import pytest

from gnn_pe_b200 import gpe, graph_io
from tests.golden_util import CASES, load_case


@pytest.mark.parametrize("name", CASES)
def test_vertices_bin_bytes(name):
    gold = load_case(name)
    pge = json.load(open(os.path.join(gold["dir"], "golden_pge.json")))
    g = graph_io.read_graph(gold["data_path"])
    e, pl = pge["e"], pge["pl"]
    x, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, e)
    pg, plg, _ = gpe.host_pge_groups(g.offsets, g.nbrs, g.labels, pl, e)
    off = g.offsets.astype(np.int64)
    deg = np.diff(off)
    nx = np.zeros_like(x)
    for j in range(int(deg.max())):          # neighbour by neighbour, in CSR order: the reference's summation order
        m = deg > j
        nx[m] += x[g.nbrs[off[:-1][m] + j]]
    assert (x + nx).tobytes() == vde.tobytes()
    rec = np.dtype([("vid", "<u4"), ("label", "<u4"), ("degree", "<u4"), ("key", "<f8"), ("x", "<f8", e), ("nx", "<f8", e),
                    ("vde", "<f8", e), ("pg", "<f8", 2 * pl * e), ("plg", "<f8", 2 * pl * e)])
    assert rec.itemsize == 20 + 3 * e * 8 + 4 * pl * e * 8   # packed, no padding
    out = np.zeros(g.V, dtype=rec)
    out["vid"], out["label"], out["degree"] = np.arange(g.V), g.labels, deg
    out["x"], out["nx"], out["vde"], out["pg"], out["plg"] = x, nx, vde, pg, plg
    data = np.uint32(g.V).tobytes() + out.tobytes()
    assert hashlib.md5(data).hexdigest() == pge["bin_md5"]
