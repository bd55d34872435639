"""Reduce an `ncu --page raw --csv` export to the columns the roofline discussion uses (one line per launch)."""
import csv
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = [hdr.index(k) for k in KEEP if k in hdr]
w = csv.writer(sys.stdout)
w.writerow([hdr[i] for i in idx])
w.writerow([units[i] for i in idx])
for r in rows[2:]:
    r2 = [r[i] for i in idx]
    r2[0] = r2[0][:110]
    w.writerow(r2)
