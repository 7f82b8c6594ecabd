"""profiles/join_ncu.json -- what the kept ncu capture of the join's persistent kernel says, in the form bench.py embeds in
its `join` object.  usage: python tools/make_join_ncu_json.py profiles/<capture>_ncu_summary.json <tests per launch> <source tag>"""
import json, sys
rec = json.load(open(sys.argv[1]))[0]
tests = float(sys.argv[2])
g = lambda k: rec.get(k)
ms = g("gpu__time_duration.sum")
ms = ms if rec.get("gpu__time_duration.sum__unit", "ms").startswith("ms") else ms / 1e3
dram_r = g("dram__bytes_read.sum") * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[rec["dram__bytes_read.sum__unit"]]
out = dict(source=sys.argv[3], kernel=g("Kernel Name")[:60], ms_under_ncu=ms, candidate_tests=tests, tests_per_s=tests / ms * 1e3,
           active_threads_per_instruction=g("smsp__thread_inst_executed_per_inst_executed.ratio"),
           issue_slots_busy_pct=g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           warps_active_per_sm=g("sm__warps_active.avg.per_cycle_active"),
           l1_hit_pct=g("l1tex__t_sector_hit_rate.pct"), l2_hit_pct=g("lts__t_sector_hit_rate.pct"),
           dram_read_bytes=dram_r, dram_sectors_per_test=dram_r / 32 / tests,
           l2_sectors_per_test=g("lts__t_sectors.sum") / tests,
           stall_long_scoreboard_per_issue=g("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
           registers_per_thread=g("launch__registers_per_thread"),
           note="a capture under ncu (cold caches, serialised): shares and ratios, not a bench time")
json.dump(out, open("profiles/join_ncu.json", "w"), indent=1)
print(json.dumps(out, indent=1))
