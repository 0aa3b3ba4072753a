#!/usr/bin/env python
"""Summarise .ncu-rep files (ncu --set full) into a small tracked text file: tools/summarize_ncu.py OUT.md rep1 rep2 ..."""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    lines = ["# ncu --set full summaries (clock-control none)", ""]
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            lines.append(f"## {rep}: unreadable")
            continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            lines.append(f"## {rep} — {d.get('Kernel Name', '?')}")
            lines.append("")
            lines.append("| metric | value | unit |")
            lines.append("|---|---|---|")
            for w in WANT:
                if w in d:
                    lines.append(f"| {w} | {d[w]} | {u.get(w, '')} |")
            try:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
                tr = sum(float(d[k].replace(",", "")) * scale[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                lines.append(f"| dram traffic (read+write) | {tr / 1e9:.3f} | Gbyte |")
            except Exception:
                pass
            lines.append("")
    open(out, "w").write("\n".join(lines))


if __name__ == "__main__":
    main()
