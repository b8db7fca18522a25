"""Per-launch summary of an `ncu --set full` capture exported with `ncu -i X.ncu-rep --page raw --csv > X.csv`.

    python profiles/summarize_ncu_raw.py X.csv > profiles/ncu_X_summary.md
"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem -> tensor core wavefronts %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.per_cycle_active", "warps active / SM"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary: {path}\n")
    print("| # | kernel | " + " | ".join(lbl for _, lbl in KEYS) + " |")
    print("|---|---|" + "---:|" * len(KEYS))
    for n, r in enumerate(rows[2:]):
        name = r[col["Kernel Name"]].replace("void ", "").split("(")[0]
        cells = []
        for k, _ in KEYS:
            if k in col and r[col[k]] != "":
                v = r[col[k]]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[col[k]]}".strip())
            else:
                cells.append("-")
        print(f"| {n} | `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
