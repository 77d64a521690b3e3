"""Summarise an .ncu-rep (one row per profiled launch) into a small text table.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt

Reads `ncu -i <rep> --page raw --csv`; prints for every launch the metrics the roofline
accounting in DESIGN.md uses (duration, FP64 pipe %, lanes active per instruction, issue slots,
stall on instruction fetch, executed FP64 thread instructions, DRAM bytes, registers, occupancy).
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("duration_ms", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"),
    ("warps_active_per_sm", "sm__warps_active.avg.per_cycle_active"),
    ("fp64_pipe_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("lanes_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("stall_no_instruction", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall_math_pipe", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall_short_scoreboard", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall_long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
    ("stall_dispatch", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"),
    ("stall_not_selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("inst_executed", "smsp__inst_executed.sum"),
    ("dfma_per_cycle", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed"),
    ("dmul_per_cycle", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"),
    ("dadd_per_cycle", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"),
    ("cycles", "sm__cycles_elapsed.avg"),
    ("dram_read", "dram__bytes_read.sum"),
    ("dram_write", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("smem_wavefronts_pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("smem_ld_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
    ("smem_st_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_mio_throttle", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("stall_lg_throttle", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: {len(rows) - 2} profiled launch(es); ncu --set full --clock-control none")
    for r in rows[2:]:
        print(f"\nkernel: {r[col['Kernel Name']]}")
        vals = {}
        for name, key in KEYS:
            if key in col:
                vals[name] = (r[col[key]], units[col[key]])
                print(f"  {name:24s} {r[col[key]]:>18s} {units[col[key]]}")
        try:
            cyc = float(vals["cycles"][0].replace(",", ""))
            f = lambda k: float(vals[k][0].replace(",", ""))
            flops = (2 * f("dfma_per_cycle") + f("dmul_per_cycle") + f("dadd_per_cycle")) * cyc
            ms = f("duration_ms") * (1e-3 if vals["duration_ms"][1] == "us" else 1.0)
            print(f"  {'fp64_flops_executed':24s} {flops:18.4e} (2*DFMA + DMUL + DADD thread instructions)")
            print(f"  {'fp64_tflops':24s} {flops / (ms * 1e-3) / 1e12:18.3f} TFLOP/s")
        except Exception as exc:  # pragma: no cover
            print("  (flops summary unavailable:", exc, ")")


if __name__ == "__main__":
    main(sys.argv[1])
