#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py quote.
Usage: python tools/ncu_summary.py gpurun_out/prof_trace.ncu-rep > profiles/r01_trace_full.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum.per_second",
    "lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum", "lts__t_requests_srcunit_tex_op_red.sum",
    "sm__cycles_elapsed.avg.per_second",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print(f"kernel: {d.get('Kernel Name')}   grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"  {k:85s} {d[k]:>18s} {u[k]}")
        stalls = sorted(((float(v), k[len(STALL):].replace("_per_issue_active.ratio", "")) for k, v in d.items()
                         if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v), reverse=True)
        print("  warp stall reasons (warps per issue-active cycle): " +
              ", ".join(f"{n} {v:.2f}" for v, n in stalls[:8]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
