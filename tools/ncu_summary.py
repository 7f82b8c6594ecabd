"""Selected metrics of one kernel from an .ncu-rep (ncu --page raw --csv) as a small CSV + JSON.
   usage: python tools/ncu_summary.py X.ncu-rep out_prefix [kernel-regex]"""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"]
if len(sys.argv) > 3:
    cmd += ["--kernel-name", "regex:" + sys.argv[3]]
rows = list(csv.reader(subprocess.check_output(cmd, stderr=subprocess.DEVNULL).decode().splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_lookup_miss.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum"]
res = []
with open(out + ".csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["launch", "metric", "unit", "value"])
    for li, r in enumerate(rows[2:]):
        rec = {}
        for name in WANT:
            if name in hdr:
                i = hdr.index(name)
                w.writerow([li, name, units[i], r[i]])
                rec[name] = r[i] if name == "Kernel Name" else (float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else None)
                if name != "Kernel Name":
                    rec[name + "__unit"] = units[i]
        res.append(rec)
json.dump(res, open(out + ".json", "w"), indent=1)
print(open(out + ".csv").read())
