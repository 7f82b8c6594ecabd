"""Summarise an ncu `--page source --csv --print-source cuda,sass` dump: hot CUDA lines by samples/instructions.
   usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K [--launch-skip n --launch-count 1] | python tools/ncu_hotlines.py [min_pct]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
minp = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[h]
ie, te, ss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
stall = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
num = lambda x: int(x) if x.isdigit() else 0
lines = [r for r in rows[h + 1:] if r and r[0].isdigit() and r[ie].isdigit()]
ti = sum(int(r[ie]) for r in lines) or 1
ts = sum(num(r[ss]) for r in lines) or 1
tt = sum(num(r[te]) for r in lines)
print(f"kernel lines={len(lines)} inst={ti} thread_inst={tt} avg_active={tt/ti:.2f} samples={ts}")
for r in lines:
    n, t, s = num(r[ie]), num(r[te]), num(r[ss])
    if n == 0 or (100 * n / ti < minp and 100 * s / ts < minp):
        continue
    top = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i][6:]) for i in stall), reverse=True)[:2]
    print(f"{r[0]:>5} inst%={100*n/ti:5.1f} smp%={100*s/ts:5.1f} act={t/n:5.1f} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]} | {r[1].strip()[:100]}")
