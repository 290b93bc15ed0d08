"""One line per profiled launch of an `ncu --set full` report: duration, DRAM bytes, throughput percentages,
occupancy, shared-memory bank conflicts, stall reasons.   python tools/ncu_full_summary.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = [("gpu__time_duration.sum", "gpu.time_duration"), ("dram__bytes_read.sum", "dram.bytes_read"),
        ("dram__bytes_write.sum", "dram.bytes_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu.dram_throughput"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm.throughput"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex.throughput"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts.throughput"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "sm.warps_active"),
        ("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "smem_dyn"), ("launch__shared_mem_per_block_static", "smem_static"),
        ("smsp__inst_executed.sum", "warp_inst"), ("lts__t_sector_hit_rate.pct", "l2_hit"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard")]
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    parts = ["kernel=" + r[ik].split("(")[0]]
    for m, short in want:
        if m in hdr:
            i = hdr.index(m)
            parts.append(f"{short}={r[i]}{units[i] if units[i] not in ('', '%') else ('%' if units[i] == '%' else '')}")
    print("; ".join(parts))
