"""Markdown table of kept bench lines.  usage: python tools/results_table.py profiles/r02*_bench_*.json"""
import json, sys
print("| file | workload | GPUs | queries/s (resident) | e2e queries/s | ms/step | select / scan / compact / join ms | scan frac (in-step) | oracle parity |")
print("|---|---|---|---|---|---|---|---|---|")
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f"| `{f}` | unreadable: {e} |")
        continue
    if "roofline" not in d or "stage_ms_per_step" not in d["roofline"]:
        continue
    st = d["roofline"]["stage_ms_per_step"]
    op = d.get("oracle_parity") or {}
    print(f"| `{f.split('/')[-1]}` | {d['config'].get('name', '?')}{' (pge)' if d.get('filter') else ''} | {d['n_gpus']} | {d['value']:,.0f} | {d['e2e']['value']:,.0f} | {d['ms_per_step']:.2f} | "
          f"{st.get('select', 0):.2f} / {st['scan']:.2f} / {st['compact']:.2f} / {st['join']:.2f} | {d['roofline']['frac']:.2f} | "
          f"{'%d ok' % op.get('checked', 0) if op.get('ok') else ('—' if not op else 'MISMATCH')} |")
