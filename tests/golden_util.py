"""Loading of the committed golden cases (tests/golden/<case>/)."""
import glob
import json
import os
import re
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["quickstart", "uniform300", "powerlaw500_e3", "uniform200_l3e4"]


def load_case(name):
    d = os.path.join(GOLDEN, name)
    with open(os.path.join(d, "golden.json")) as f:
        gold = json.load(f)
    qfiles = [p for p in glob.glob(os.path.join(d, "q*.graph")) if re.fullmatch(r"q\d+\.graph", os.path.basename(p))]
    qfiles.sort(key=lambda p: int(os.path.basename(p)[1:-6]))
    gold["dir"] = d
    gold["data_path"] = os.path.join(d, "data.graph")
    gold["membership_path"] = os.path.join(d, "membership.txt")
    gold["query_paths_files"] = qfiles
    assert len(qfiles) == len(gold["queries"])
    return gold


def hex_to_f64(hexes):
    return np.array([struct.unpack(">d", bytes.fromhex(h))[0] for h in hexes], dtype=np.float64)
