"""Selected metrics of an `ncu --page raw --csv` export as a markdown summary (one table per captured launch).

    ncu -i gpurun_out/prof_acc_ntt.ncu-rep --page raw --csv > gpurun_out/prof_acc_ntt_raw.csv
    python scripts/ncu_raw_to_md.py gpurun_out/prof_acc_ntt_raw.csv > profiles/rNN_ncu_accumulate_ntt_full.md
"""
import csv
import sys

METRICS = [
    "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("## `%s`\n" % r[col["Kernel Name"]][:70])
    print("| metric | value | unit |\n|---|---|---|")
    for m in METRICS:
        if m in col:
            print("| %s | %s | %s |" % (m, r[col[m]], units[col[m]]))
    print()
