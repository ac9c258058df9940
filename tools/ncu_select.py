#!/usr/bin/env python
"""Keep the columns of an `ncu -i X.ncu-rep --page raw --csv` dump that profiles/README.md cites.

    ncu -i gpurun_out/r02_decode.ncu-rep --page raw --csv | python tools/ncu_select.py > profiles/r02_ncu_decode_raw_selected.csv
"""
import csv
import sys

KEEP_EXACT = ("ID", "Kernel Name", "Block Size", "Grid Size")
KEEP_PREFIX = (
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.sum", "dram__cycles_elapsed.sum",
    "dram__sectors_read.sum", "dram__sectors_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__registers_per_thread",
    "smsp__cycles_active.avg.per_second", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__inst_executed_op_", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_", "smsp__average_warp_latency_issue_stalled_", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__inst_executed_pipe_tensor",
)


def main():
    rows = list(csv.reader(sys.stdin))
    head = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names = rows[head]
    keep = [i for i, n in enumerate(names) if n in KEEP_EXACT or any(n.startswith(p) for p in KEEP_PREFIX)]
    w = csv.writer(sys.stdout)
    for r in rows[head:]:
        if len(r) == len(names):
            w.writerow([r[i] for i in keep])


if __name__ == "__main__":
    main()
