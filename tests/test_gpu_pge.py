"""GNN-PGE on the GPU (k4_pge.cu through the C ABI) against the golden vectors of the unmodified reference
(tests/golden/<case>/golden_pge.json): path groups bit for bit, candidate sets, answers.

The kernels were written at the very end of round 1, after the GPU budget was spent: they compile for sm_100a but
have not been seen on a GPU yet.  Until they have, every case runs in its OWN process (a fault cannot touch the CUDA
context of the parity suite) and is marked xfail(strict=False): a pass shows up as XPASS, a failure does not turn
the suite red.  Remove the marker once green."""
import os
import subprocess
import sys

import pytest

from tests.golden_util import CASES, ROOT

pytestmark = pytest.mark.gpu

_SCRIPT = r'''
import hashlib, json, os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from gnn_pe_b200 import gpe, graph_io
from tests.golden_util import load_case

gold = load_case(sys.argv[1])
pge = json.load(open(os.path.join(gold["dir"], "golden_pge.json")))
g = graph_io.read_graph(gold["data_path"])
ctx = gpe.GpeContext(0)
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
ctx.close()
print("ok", sys.argv[1], got)
'''


@pytest.mark.xfail(strict=False, reason="GNN-PGE kernels: written after round 1's GPU budget was spent, not yet seen on a GPU")
@pytest.mark.parametrize("name", CASES)
def test_pge_matches_reference_golden(name):
    r = subprocess.run([sys.executable, "-c", _SCRIPT, name], capture_output=True, timeout=90, cwd=ROOT)
    assert r.returncode == 0, (r.stdout.decode()[-1500:], r.stderr.decode()[-3000:])
