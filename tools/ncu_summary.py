"""Summarise an .ncu-rep (read on the CPU box): per-kernel duration, tensor / DRAM utilisation, stalls."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled",
        "sm__inst_executed_pipe_xu", "smsp__issue_active.avg.pct", "sm__pipe_fma", "sm__pipe_alu", "pipe_tmem", "tensor"]
extra = sys.argv[2:]
for r in rows[2:]:
    print("=" * 100)
    for i, h in enumerate(hdr):
        if any(k in h for k in keys + extra):
            print(f"  {h:95s} {r[i][:70]}  [{rows[1][i]}]")
