"""GNN-PGE on the GPU (k4_pge.cu through the C ABI) against the golden vectors of the unmodified reference
(tests/golden/<case>/golden_pge.json): path groups bit for bit, candidate sets, answers."""
import hashlib
import json
import os

import pytest

from gnn_pe_b200 import gpe, graph_io
from tests.golden_util import CASES, load_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_pge_matches_reference_golden(name):
    gold = load_case(name)
    pge = json.load(open(os.path.join(gold["dir"], "golden_pge.json")))
    g = graph_io.read_graph(gold["data_path"])
    ctx = gpe.GpeContext(0)
    try:
        ctx.set_graph(g.offsets, g.nbrs, g.labels)
        x, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, pge["e"])
        ctx.set_embeddings(vde)
        ctx.pge_build(pge["pl"], x)
        pg, plg, has = ctx.pge_dump_groups()
        assert hashlib.md5(pg.tobytes()).hexdigest() == pge["pg_md5"], "pg"
        assert hashlib.md5(plg.tobytes()).hexdigest() == pge["plg_md5"], "plg"
        queries = [graph_io.read_graph(qf) for qf in gold["query_paths_files"]]
        ctx.pge_batch_upload(queries)
        ctx.pge_batch_filter()
        off, cand = ctx.batch_get_candidates()
        slot = 0
        for qrec, q in zip(pge["queries"], queries):
            for cset in qrec["candidates"]:
                assert cand[int(off[slot]):int(off[slot + 1])].tolist() == cset, (qrec["file"], slot)
                slot += 1
        ctx.batch_join()
        raw = ctx.batch_download()
        got = [ctx.clamp(int(r), gpe.LIMIT_MAX) for r in raw]
        assert got == [qrec["answer"] for qrec in pge["queries"]], got
        assert ctx.pge_query_batch(queries).tolist() == got
    finally:
        ctx.close()
