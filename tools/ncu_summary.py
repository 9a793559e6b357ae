"""Summarise an .ncu-rep (raw page) into a small CSV + per-kernel dram traffic: python tools/ncu_summary.py rep out.csv"""
import csv
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size']


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([rows[1][i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        rd, wr = r[hdr.index('dram__bytes_read.sum')], r[hdr.index('dram__bytes_write.sum')]
        ur, uw = rows[1][hdr.index('dram__bytes_read.sum')], rows[1][hdr.index('dram__bytes_write.sum')]
        print(f"{name[:95]:95s} t={r[hdr.index('gpu__time_duration.sum')]:>10s}us rd={rd}{ur} wr={wr}{uw} tensor%={r[hdr.index('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')][:5]} dram%={r[hdr.index('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')][:5]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
