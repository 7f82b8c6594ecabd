"""The oracle's restatement of the reference's sibling variant GNN-PGE (per-vertex path groups)
against golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden_pge.py): path groups as written
by `main -m offline` (md5 over the raw doubles + samples as bit patterns), candidate sets from the reference's own
R*-tree traversal, answers from the unmodified binary.  CPU only: this is the groundwork (oracle first) for the
GPU version of that filter."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle
from tests.golden_util import CASES, load_case


@pytest.fixture(scope="module", params=CASES)
def case(request):
    gold = load_case(request.param)
    with open(os.path.join(gold["dir"], "golden_pge.json")) as f:
        pge = json.load(f)
    return gold, pge, oracle.OracleGraph.load(gold["data_path"])


def test_path_groups_bit_exact(case):
    gold, pge, g = case
    pg, plg, has = oracle.pge_groups(g, pge["pl"], pge["e"])
    assert pg.shape == (pge["V"], 2 * pge["pl"] * pge["e"])
    assert hashlib.md5(pg.tobytes()).hexdigest() == pge["pg_md5"]
    assert hashlib.md5(plg.tobytes()).hexdigest() == pge["plg_md5"]
    for v, s in pge["sample"].items():
        assert [float(x).hex() for x in pg[int(v)]] == s["pg"]
        assert [float(x).hex() for x in plg[int(v)]] == s["plg"]
    # boxes are boxes; a vertex without any path of pl vertices stores [vde, vde | 0 ...] (main.cpp:103-121)
    assert np.all(pg[:, 0::2] <= pg[:, 1::2]) and np.all(plg[:, 0::2] <= plg[:, 1::2])
    e = pge["e"]
    for v in np.flatnonzero(has == 0):
        assert np.all(pg[v, 2 * e:] == 0.0) and np.all(pg[v, 0:2 * e:2] == pg[v, 1:2 * e:2])


def test_candidates_equal_the_reference_traversal(case):
    gold, pge, g = case
    for qrec, qf in zip(pge["queries"], gold["query_paths_files"]):
        assert os.path.basename(qf) == qrec["file"]
        q = oracle.OracleGraph.load(qf)
        got = oracle.pge_filter(g, q, pge["pl"], pge["e"])
        assert [c.tolist() for c in got] == qrec["candidates"], qrec["file"]


def test_answers(case):
    gold, pge, g = case
    g.enumerate(3, g.degree_order())  # (the shared refinement needs nothing from it; keeps the handle in its usual state)
    for qrec, qf in zip(pge["queries"], gold["query_paths_files"]):
        assert qrec["answer"] == qrec["main_answer"]
        q = oracle.OracleGraph.load(qf)
        assert oracle.pge_online(g, q, pge["pl"], pge["e"]) == qrec["answer"], qrec["file"]
