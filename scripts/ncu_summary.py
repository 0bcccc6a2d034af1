"""Export selected raw metrics of an `ncu --set full` report as a small CSV (the .ncu-rep itself is not committed).
usage: python scripts/ncu_summary.py gpurun_out/<name>.ncu-rep profiles/<name>_summary.csv"""
import csv
import io
import subprocess
import sys

COLS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "launch__cluster_size"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head = rows[0]
    idx = [head.index(c) if c in head else None for c in COLS]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(COLS)
        for r in rows[1:]:
            w.writerow([r[i] if i is not None and i < len(r) else "" for i in idx])
    print(f"{len(rows) - 2} launches -> {out}")


if __name__ == "__main__":
    main()
