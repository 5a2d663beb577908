"""Print the headline metrics of every kernel in an .ncu-rep (via `ncu --page raw --csv`)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
for r in rows[2:]:
    print("-----", r[hdr.index("Kernel Name")][:80])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:62s} {r[i]:>16s} {units[i]}")
