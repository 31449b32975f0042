"""Prints the handful of ncu metrics the profiles/ summaries quote.

    ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tests/ncu_summary.py raw.csv
"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct",
        "sm__inst_executed_pipe_alu.sum ", "sm__inst_executed_pipe_fma.sum.pct", "sm__inst_executed_pipe_lsu.sum.pct",
        "sm__inst_executed_pipe_lsu.sum ", "sm__pipe_alu_cycles_active.avg.pct", "smsp__inst_executed.sum ",
        "sm__inst_executed.sum ", "sm__inst_executed_pipe_fmaheavy", "sm__inst_executed_pipe_fmalite",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued", "smsp__average_warps_issue_stalled",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", "sm__cycles_elapsed.max", "sm__cycles_active.avg ",
        "smsp__warps_eligible.avg.per_cycle_active", "lts__t_sectors_op_read.sum ", "lts__t_sectors_op_write.sum ",
        "lts__t_bytes.sum ", "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_adu", "sm__inst_executed_pipe_cbu",
        "sm__inst_executed_pipe_uniform", "smsp__cycles_active.avg ", "sm__pipe_shared_cycles_active",
        "l1tex__lsu_writeback_active", "smsp__inst_executed_op_shared", "smsp__inst_issued.avg.per_cycle_active"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    for row in rows[2:]:
        print("==", row[hdr.index("Kernel Name")][:90])
        for h, u, v in zip(hdr, rows[1], row):
            name = h.split(".", 2)[-1] if h.count(".") >= 3 and h.split(".")[1].startswith("Triage") else h
            if any((name + " ").startswith(w) or name.startswith(w) for w in WANT):
                print(f"{name} [{u}] = {v}")


if __name__ == "__main__":
    main(sys.argv[1])
