"""Summarises .ncu-rep captures (ncu --set full) as text: python tools/ncu_summary.py report.ncu-rep > profiles/....txt"""
import csv
import subprocess
import sys

WANT = [
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit: registers (CTAs/SM)"), ("launch__occupancy_limit_shared_mem", "occupancy limit: shared memory (CTAs/SM)"),
    ("gpu__time_duration.sum", "duration"), ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC per SM (of 4)"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per instruction"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active, % of peak"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput, % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch resolving / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not selected / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no instruction / issue"),
]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    print(f"# {rep}: ncu --set full --clock-control none (cold, serialised launches: shares and ratios, not bench values)")
    for r in rows[2:]:
        print("kernel:", r[h.index("Kernel Name")].split("(")[0])
        for key, label in WANT:
            if key in h:
                i = h.index(key)
                print(f"  {label:46s} {r[i]} {units[i]}")
        print()
