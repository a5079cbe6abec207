"""Compact table of the metrics DESIGN.md / profiles/README.md quote from an `ncu --csv --page raw` file."""
import csv
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    cols = [(w, names.index(w)) for w in WANT if w in names]
    out = csv.writer(sys.stdout)
    out.writerow([f"{w} [{units[i]}]" if units[i] else w for w, i in cols])
    for r in data:
        out.writerow([r[i][:70] for _, i in cols])


if __name__ == "__main__":
    main(sys.argv[1])
