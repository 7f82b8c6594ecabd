#!/usr/bin/env python
"""Static evidence from the built library (no GPU needed): registers / stack / local memory per kernel of libgpe.so
(cuobjdump --dump-resource-usage) and the SASS mnemonics that show the bulk-copy engine, mbarriers, FP64 compares and
reductions (cuobjdump -sass).  Usage: python tools/resource_usage.py > profiles/<round>_resource_usage.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gnn_pe_b200", "libgpe.so")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    res = []
    for o in out:
        o = re.sub(r"^void ", "", o.replace("(anonymous namespace)::", "").replace("gpe::", ""))
        res.append(o[:o.index("(")] if "(" in o else o)
    return res


def main():
    txt = subprocess.check_output(["cuobjdump", "--dump-resource-usage", LIB], text=True)
    rows = re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", txt)
    own = [r for r in rows if re.search(r"k\d_|scan|gpe", r[0]) and "cub" not in r[0]]
    names = demangle([r[0] for r in own])
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(rows)} device functions, {len(own)} of this repo's (the rest are CUB sort/scan instances)")
    print(f"{'kernel':70s} {'regs':>5s} {'stack':>6s} {'static smem':>12s} {'local':>6s}")
    for n, r in sorted(zip(names, own)):
        print(f"{n[:70]:70s} {r[1]:>5s} {r[2]:>6s} {r[3]:>12s} {r[4]:>6s}")
    sass = subprocess.check_output(["cuobjdump", "-sass", LIB], text=True)
    cnt = collections.Counter()
    fn = None
    per_fn = collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for key in ("UBLKCP", "SYNCS", "DSETP", "DADD", "ATOMS", "ATOMG", "RED", "REDUX", "MATCH", "VOTE", "SHFL", "LDL", "STL"):
                if op.startswith(key):
                    cnt[key] += 1
                    per_fn[fn][key] += 1
    print("\n# SASS mnemonic counts over all sm_100a cubins")
    for k, v in sorted(cnt.items()):
        print(f"{k:8s} {v}")
    print("\n# kernels with bulk copies (TMA engine) / local-memory traffic")
    fns = list(per_fn)
    for n, f in sorted(zip(demangle(fns), fns)):
        c = per_fn[f]
        if c["UBLKCP"] or c["LDL"] or c["STL"]:
            print(f"{n[:70]:70s} UBLKCP={c['UBLKCP']} SYNCS={c['SYNCS']} LDL={c['LDL']} STL={c['STL']}")


if __name__ == "__main__":
    main()
