"""Extract the metrics the notes cite from an .ncu-rep (read here, without a GPU):  python tools/ncu_summary.py in.ncu-rep out.csv"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_uniform.sum"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel"] + [r[hdr.index("Kernel Name")][:90] for r in data])
    for i, name in enumerate(hdr):
        if name in KEYS or "issue_stalled" in name and name.endswith("per_issue_active.ratio"):
            w.writerow([name + " [" + units[i] + "]"] + [r[i] for r in data])
print("wrote", sys.argv[2])
