#!/usr/bin/env python3
"""Digest of `ncu --page raw --csv` captures: one line per profiled launch with the columns DESIGN.md quotes.
  python tools/ncu_rows.py profiles/r02_full_ntt_pw_raw.csv profiles/r02_full_small_raw.csv > profiles/r02_ncu_summary.txt"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "ms"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy%"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "alu%"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue%"),
        ("smsp__inst_executed.sum", "warp_inst"),
        ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_disp"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("sass__inst_executed_local_loads", "LDL"), ("sass__inst_executed_local_stores", "STL"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conf")]


def main():
    print("# one line per launch captured with `ncu --set full --clock-control none` (tools/profile_round.sh); stall columns = warps stalled per issue-active cycle")
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        print("## " + path)
        print("kernel | " + " | ".join(c[1] for c in COLS))
        ki = hdr.index("Kernel Name")
        for r in rows[2:]:
            name = r[ki].split("(")[0].replace("void ", "")[:34]
            vals = []
            for col, _ in COLS:
                if col in hdr:
                    v = r[hdr.index(col)].replace(",", "")
                    u = units[hdr.index(col)]
                    try:
                        f = float(v)
                        if col.startswith("dram__bytes"):
                            f *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3}.get(u, 1.0)
                            vals.append("%.2fGB" % (f / 1e9))
                        elif col == "gpu__time_duration.sum":
                            f *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(u, 1.0)   # -> ms
                            vals.append("%.3f" % f)
                        elif f > 1e6:
                            vals.append("%.0fM" % (f / 1e6))
                        else:
                            vals.append(("%.2f" % f).rstrip("0").rstrip("."))
                    except ValueError:
                        vals.append(v[:10])
                else:
                    vals.append("-")
            print(name + " | " + " | ".join(vals))


if __name__ == "__main__":
    main()
