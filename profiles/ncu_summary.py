#!/usr/bin/env python
"""Prints the metrics we track from an .ncu-rep (run where ncu is installed; no GPU needed):
    python profiles/ncu_summary.py gpurun_out/<name>.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum",
    "sm__inst_executed_pipe_uniform.sum",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum",
    "smsp__inst_executed_op_global_atom.sum", "smsp__inst_executed_op_shared_atom.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        print("==", name)
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f"  {h:75s} {vals[i]:>20s} {units[i]}")
        stalls = [(float(vals[i]), h[len(STALL):-len('_per_issue_active.ratio')]) for i, h in enumerate(hdr)
                  if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and vals[i]]
        print("  stall reasons (warps per issue-active cycle):",
              ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
    if False and "--source" in sys.argv:
        top = int(sys.argv[sys.argv.index("--source") + 1])
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(src)))
        hdr = rows[0]
        try:
            ci = hdr.index("# Warp Stall Sampling (All Samples)")
        except ValueError:
            ci = next(i for i, h in enumerate(hdr) if "Sampling" in h)
        si = hdr.index("Source")
        ei = next((i for i, h in enumerate(hdr) if h.startswith("# Instructions Executed") or h == "Instructions Executed"), None)
        body = [r for r in rows[1:] if len(r) > ci and r[ci].replace('.', '', 1).isdigit()]
        tot = sum(float(r[ci]) for r in body) or 1.0
        print(f"== top {top} source lines by stall samples (total {tot:.0f})")
        for r in sorted(body, key=lambda r: -float(r[ci]))[:top]:
            ex = r[ei] if ei is not None else ""
            print(f"  {100*float(r[ci])/tot:5.1f}%  exec={ex:>12s}  {r[si].strip()[:110]}")


if __name__ == "__main__":
    main()
