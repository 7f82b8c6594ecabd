"""CPU-side checks of the product library: it loads, exports every symbol include/gpe.h declares, fails
loudly without a GPU, and its host mirror (loader, embeddings, query plan) agrees with the golden vectors
of the unmodified reference.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

from gnn_pe_b200 import gpe, graph_io
from tests.golden_util import CASES, ROOT, hex_to_f64, load_case


@pytest.fixture(scope="module")
def lib():
    gpe.build()
    return gpe.lib()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "gpe.h")).read()
    declared = set(re.findall(r"\b(gpe_[a-z_0-9]+)\s*\(", header))
    assert declared == set(gpe.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.gpe_abi_version() == 1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = lib.gpe_create(0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.gpe_last_error(None)
    with pytest.raises(gpe.GpeError):
        gpe.GpeContext(0)


def test_product_never_imports_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "gnn_pe_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(d, f)).read()
                for line in src.splitlines():
                    s = line.strip()
                    if s.startswith(("import ", "from ", "#include")):
                        assert "oracle" not in s, (f, line)


def test_clamp_answer(lib):
    assert lib.gpe_clamp_answer(45426, gpe.LIMIT_MAX) == 45426
    assert lib.gpe_clamp_answer(45426, 100) == 100
    assert lib.gpe_clamp_answer(5, 100) == 5
    assert lib.gpe_clamp_answer(5, 0) == 1   # custom.h:851: the limit is tested after counting a match
    assert lib.gpe_clamp_answer(0, 0) == 0


@pytest.mark.parametrize("name", CASES)
def test_host_mirror_against_golden(lib, name):
    gold = load_case(name)
    off, nbr, lab = gpe.host_load_graph(gold["data_path"])
    g = graph_io.read_graph(gold["data_path"])
    assert np.array_equal(off, g.offsets) and np.array_equal(nbr, g.nbrs) and np.array_equal(lab, g.labels)
    e, L = gold["e"], gold["l"] + 1
    x, vde = gpe.host_gen_vde(off, nbr, lab, e)
    q0 = gold["queries"][0]
    for v, hexes in q0["data_vde_sample"].items():
        assert vde[int(v)].tobytes() == hex_to_f64(hexes).tobytes()
    for qf, rec in zip(gold["query_paths_files"], gold["queries"]):
        qo, qn, ql = gpe.host_load_graph(qf)
        plan = gpe.host_query_plan(qo, qn, ql, L, e)
        assert plan["vids"].tolist() == [p["vids"] for p in rec["plan"]]
        for j, p in enumerate(rec["plan"]):
            assert plan["pde"][j].tobytes() == hex_to_f64(p["pde"]).tobytes()
        qx, _ = gpe.host_gen_vde(qo, qn, ql, e)
        for labn, hexes in rec["label_x"].items():
            for u in np.nonzero(ql == int(labn))[0]:
                assert qx[u].tobytes() == hex_to_f64(hexes).tobytes()


def test_host_loader_errors(lib, tmp_path):
    with pytest.raises(gpe.GpeError):
        gpe.host_load_graph(str(tmp_path / "missing.graph"))


@pytest.mark.parametrize("name", CASES)
def test_host_pge_groups_against_reference_golden(lib, name):
    """The host mirror of GNN-PGE's path groups (what the batch upload computes per query) against data_vertices.bin of
    the unmodified reference (tests/golden/<case>/golden_pge.json)."""
    import hashlib
    import json
    gold = load_case(name)
    pge = json.load(open(os.path.join(gold["dir"], "golden_pge.json")))
    off, nbr, lab = gpe.host_load_graph(gold["data_path"])
    pg, plg, has = gpe.host_pge_groups(off, nbr, lab, pge["pl"], pge["e"])
    assert hashlib.md5(pg.tobytes()).hexdigest() == pge["pg_md5"]
    assert hashlib.md5(plg.tobytes()).hexdigest() == pge["plg_md5"]


def test_host_mirror_accepts_any_32_bit_label():
    """gen_vde_x only seeds a generator with the label (custom.h:492-511), so any 32-bit label is legal: the host mirror
    must not size a table by the largest label id (an allocation of tens of GB, or std::bad_alloc through the C ABI)."""
    import numpy as np
    from gnn_pe_b200 import gpe
    from oracle import oracle
    off = np.array([0, 1, 3, 4], dtype=np.uint32)
    nbr = np.array([1, 0, 2, 1], dtype=np.uint32)
    lab = np.array([0x80000000, 7, 0xFFFFFFF0], dtype=np.uint32)
    x, vde = gpe.host_gen_vde(off, nbr, lab, 3)
    for v in range(3):
        assert x[v].tobytes() == oracle.label_embedding(int(lab[v]), 3).tobytes()
    assert vde[1].tobytes() == (x[1] + (x[0] + x[2])).tobytes()
    plan = gpe.host_query_plan(off, nbr, lab, 3, 3)
    assert len(plan["vids"]) >= 1 and set(plan["labels"].ravel().tolist()) <= set(lab.tolist())


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/gpe.h must compile as C99 (no C++ in the signatures) and a C program must link
    against libgpe.so with nothing but -lgpe (the cgo / JNI / ctypes situation of INTEGRATION.md)."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "gpe.h"\n#include <stdio.h>\n'
                   'int main(void) {\n'
                   '    gpe_stats st; gpe_batch b; (void)st; (void)b;\n'
                   '    printf("%d %llu\\n", gpe_abi_version(), (unsigned long long)gpe_clamp_answer(7, 3));\n'
                   '    return 0;\n}\n')
    exe = tmp_path / "t"
    libdir = os.path.join(ROOT, "gnn_pe_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-lgpe", f"-Wl,-rpath,{libdir}"])
    assert subprocess.check_output([str(exe)]).decode().split() == ["1", "3"]


def test_null_arguments_never_crash():
    """Every entry point of include/gpe.h called with all-NULL / zero arguments (no context can exist without a GPU, so
    this is also what a caller that ignored a failed gpe_create would do): an error code or a harmless value, never a
    crash.  Runs in a child process so that a crash is a test failure with the function's name, not a dead test run."""
    import subprocess
    import sys
    header = open(os.path.join(ROOT, "include", "gpe.h")).read()
    protos = re.findall(r"\n(int|uint64_t|void \*|const char \*|void)\s*(gpe_[a-z_0-9]+)\s*\(([^;]*?)\);", header)
    assert len(protos) == len(gpe.SYMBOLS)
    calls = [(ret, name, 0 if args.strip() in ("", "void") else args.count(",") + 1) for ret, name, args in protos]
    child = (
        "import ctypes as C, sys\n"
        f"L = C.CDLL({os.path.join(ROOT, 'gnn_pe_b200', 'libgpe.so')!r})\n"
        f"for ret, name, n in {calls!r}:\n"
        "    print(name, flush=True)\n"
        "    f = getattr(L, name); f.restype = C.c_uint64\n"
        "    r = f(*([C.c_void_p(0)] * n))\n"
        "    if ret == 'int' and name != 'gpe_abi_version' and (r & 0xffffffff) == 0: sys.exit('accepted NULLs: ' + name)\n"
        "print('done')\n")
    r = subprocess.run([sys.executable, "-c", child], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("done"), (r.returncode, r.stdout.strip().splitlines()[-1:], r.stderr[-300:])
