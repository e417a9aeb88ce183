"""Distil an .ncu-rep (one kernel launch, --set full) into the small metric CSV kept under profiles/.
    python scripts/distil_ncu.py gpurun_out/r01e/rk45_full.ncu-rep profiles/r01e_rk45_ncu_full_metrics.csv"""
import csv, io, re, subprocess, sys

KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum$|gpu__time_duration\.sum$|launch__(block_size|grid_size|registers_per_thread|"
                  r"shared_mem_per_block_(dynamic|static)|occupancy_limit_.*|waves_per_multiprocessor)$|"
                  r"sm__inst_executed_pipe_(fp64|lsu|alu|fma|xu|uniform)\.avg\.pct_of_peak_sustained_active$|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active$|sm__throughput\.avg\.pct_of_peak_sustained_elapsed$|"
                  r"sm__inst_executed\.avg\.per_cycle_(active|elapsed)$|sm__issue_active\.avg\.pct.*|smsp__issue_active\.avg\.pct.*|"
                  r"smsp__inst_executed\.sum$|smsp__warps_eligible\.avg\.per_cycle_active$|smsp__issue_active\.avg\.per_cycle_active$|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$|smsp__cycles_active\.avg$|"
                  r"l1tex__t_sector_hit_rate\.pct$|lts__t_sector_hit_rate\.pct$|lts__t_bytes\.sum$|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|smsp__inst_executed_op_(local|shared|global)_(ld|st)\.sum$|"
                  r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$|dram__throughput\.avg\.pct_of_peak_sustained_elapsed$|"
                  r"sm__sass_thread_inst_executed_op_d(fma|mul|add)_pred_on\.sum$|local_(load|store)_.*)")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
names, units, vals = rows[0], rows[1], rows[2]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f); w.writerow(["metric", "unit", "value"])
    for n, u, v in sorted(zip(names, units, vals)):
        if KEEP.match(n):
            w.writerow([n, u, v])
print(open(sys.argv[2]).read())
