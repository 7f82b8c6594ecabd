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


@pytest.mark.parametrize("name", CASES)
def test_host_cli_pge(name, tmp_path):
    """host/main --filter pge keeps GNN-PGE's command line (GNN-PGE/src/main.cpp): data_vertices.bin byte for byte as the
    unmodified binary writes it, its `Answer Num:` line online, a directory of queries as one batch."""
    import shutil
    import subprocess
    from tests.golden_util import ROOT
    main = os.path.join(ROOT, "host", "main")
    if not os.path.exists(main):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "host")])
    gold = load_case(name)
    pge = json.load(open(os.path.join(gold["dir"], "golden_pge.json")))
    d = str(tmp_path) + "/"
    os.makedirs(d + "gnn-pge")
    common = [main, "--filter", "pge", "-f", d, "-d", gold["data_path"], "-l", str(pge["pl"]), "-e", str(pge["e"])]
    subprocess.check_call(common + ["-m", "offline"])
    assert hashlib.md5(open(d + "gnn-pge/data_vertices.bin", "rb").read()).hexdigest() == pge["bin_md5"]
    qdir = d + "queries"
    os.makedirs(qdir)
    for i, (qf, qrec) in enumerate(zip(gold["query_paths_files"], pge["queries"])):
        shutil.copy(qf, f"{qdir}/q{i:03d}.graph")
        if i < 2:
            out = subprocess.check_output(common + ["-m", "online", "-q", qf]).decode()
            assert out.startswith(f"Answer Num: {qrec['main_answer']} Query Time (ms): "), out
    out = subprocess.check_output(common + ["-m", "online", "-q", qdir]).decode().splitlines()
    assert [int(line.split("Answer Num: ")[1]) for line in out[:-1]] == [q["main_answer"] for q in pge["queries"]]
    out = subprocess.check_output(common + ["-m", "online", "-q", gold["query_paths_files"][0], "-n", "3"]).decode()
    assert out.startswith(f"Answer Num: {min(3, pge['queries'][0]['main_answer'])} ")
    r = subprocess.run(common + ["-m", "online", "-q", qdir, "-g", "2"], capture_output=True)
    assert r.returncode == 1 and b"one GPU" in r.stderr
