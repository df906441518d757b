"""Columns of interest of an `ncu --set full` report as a small CSV.

  python profiles/summarize_ncu.py gpurun_out/prof_v4_b4096.ncu-rep profiles/r1/ncu_full_summary_b4096_v4.csv
"""
import csv
import io
import subprocess
import sys

COLUMNS = [
    'Kernel Name', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
    'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    # stalls per issued instruction (the latency-bound kernels of the chain)
    'sm__cycles_elapsed.max',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']


def main():
  report, out = sys.argv[1], sys.argv[2]
  raw = subprocess.check_output(['ncu', '-i', report, '--page', 'raw', '--csv'],
                                stderr=subprocess.DEVNULL).decode()
  rows = list(csv.reader(io.StringIO(raw)))
  header = rows[0]
  keep = [header.index(c) for c in COLUMNS if c in header]
  with open(out, 'w', newline='') as f:
    w = csv.writer(f)
    for r in rows:
      w.writerow([r[i] for i in keep])


if __name__ == '__main__':
  main()
