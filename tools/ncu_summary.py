"""Summarise an Nsight Compute report: `python tools/ncu_summary.py report.ncu-rep [out.json]`.

Reads `ncu -i report --page raw --csv` and keeps, per profiled launch, the metrics the DESIGN/roofline discussion
uses (duration, DRAM bytes, registers, occupancy limiters, pipe utilisation, issue rate, stall reasons)."""
import csv, io, json, subprocess, sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
out = []
for row in data:
    rec = {"Kernel Name": row[hdr.index("Kernel Name")]}
    stalls = {}
    for name, unit, val in zip(hdr, units, row):
        if name in KEEP:
            rec[name] = {"value": val, "unit": unit}
        elif name.startswith(STALL_PREFIX) and name.endswith("_per_issue_active.ratio"):
            try:
                stalls[name[len(STALL_PREFIX):-len("_per_issue_active.ratio")]] = float(val.replace(",", ""))
            except ValueError:
                pass
    rec["stalls_per_issue (top 6)"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
    out.append(rec)
text = json.dumps(out if len(out) > 1 else out[0], indent=1)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
else:
    print(text)
