"""Print the metrics that matter from an .ncu-rep (one column per captured kernel).
python scripts/ncu_keys.py gpurun_out/x.ncu-rep [extra_substring ...]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__cycles_active.avg",
        "gpc__cycles_elapsed.max.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tc", "sm__inst_executed_pipe_uniform",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_bytes.sum.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
extra = sys.argv[2:]
ki = hdr.index("Kernel Name")
print("kernels:", [d[ki].split("(")[0][-40:] for d in data])
for i, h in enumerate(hdr):
    if h in KEYS or any(e in h for e in extra) or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        vals = [d[i] for d in data]
        if all(v in ("0", "") for v in vals):
            continue
        print(f"{h} [{units[i]}] = " + " | ".join(vals))
