#!/usr/bin/env python
"""Per-launch table of the metrics DESIGN.md / bench.py quote, from an `ncu --set full` report (ncu -i REP --page raw --csv).
  python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, body = rows[0], rows[1], rows[2:]
        idx = [hdr.index(c) for c in COLS if c in hdr]
        print(f"\n## {rep}\n")
        print("| " + " | ".join(hdr[i] for i in idx) + " |")
        print("|" + "---|" * len(idx))
        print("| " + " | ".join(units[i] for i in idx) + " |")
        for r in body:
            cells = [r[i] for i in idx]
            cells[0] = cells[0].split("(")[0].replace("void ", "").replace("wb200::", "").replace("<unnamed>::", "")
            print("| " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
