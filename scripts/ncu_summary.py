"""Summaries of `ncu --set full` captures for profiles/: python scripts/ncu_summary.py <report.ncu-rep> <out.txt> [title].
Reads the raw page (ncu -i ... --page raw --csv) and keeps the metrics the roofline discussion uses."""
import csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_sample_buffer_full"]
with open(out, "w") as f:
    f.write(f"{title}\n(ncu --set full --clock-control none; one launch; raw page)\n\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        for k in KEEP:
            if k in d and d[k] != "":
                f.write(f"{k:100s} {d[k]} {u.get(k, '')}\n")
        rd, wr = float(d.get("dram__bytes_read.sum", 0) or 0), float(d.get("dram__bytes_write.sum", 0) or 0)
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        tot = rd * scale.get(u.get("dram__bytes_read.sum", "byte"), 1.0) + wr * scale.get(u.get("dram__bytes_write.sum", "byte"), 1.0)
        f.write(f"{'dram bytes read + written per launch':100s} {tot:.6g} byte\n\n")
        print(d.get("Kernel Name", "?")[:60], "dram_total_bytes", tot)
