import csv, subprocess, sys, io
rep = sys.argv[1]
want = ["gpu__time_duration.sum","launch__grid_size","launch__block_size","launch__registers_per_thread","launch__waves_per_multiprocessor",
"sm__warps_active.avg.pct_of_peak_sustained_active","sm__inst_executed.sum","smsp__inst_executed.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed",
"dram__throughput.avg.pct_of_peak_sustained_elapsed","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_sector_hit_rate.pct",
"smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__average_warp_latency_issue_stalled_long_scoreboard.pct","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
"smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
"smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
"smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","sm__maximum_warps_per_active_cycle_pct"]
out = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]; units = rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("##", d["Kernel Name"][:80], "id", d["ID"])
    for w in want:
        if w in d: print("  %-90s %s %s" % (w, d[w], units[hdr.index(w)]))
