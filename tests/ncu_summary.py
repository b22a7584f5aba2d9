"""Condenses an .ncu-rep into the handful of numbers DESIGN.md / profiles/ quote (run in the authoring container)."""
import csv
import subprocess
import sys


def main(path, kernel_filter=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed",
            "lts__t_sectors_srcunit_tex_op_red.sum", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
    for r in data:
        name = r[col["Kernel Name"]]
        if kernel_filter and kernel_filter not in name:
            continue
        print("kernel:", name[:100], "grid", r[col["Grid Size"]], "block", r[col["Block Size"]])
        for w in want:
            if w in col:
                print("  %-82s %-10s %s" % (w, units[col[w]], r[col[w]]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
