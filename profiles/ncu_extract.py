#!/usr/bin/env python3
"""Print the judged metrics from `ncu -i X.ncu-rep --page raw --csv` output (stdin or file)."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max']


def main():
    f = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
    r = list(csv.reader(f))
    hdr, units, rows = r[0], r[1], r[2:]
    names = [row[hdr.index('Kernel Name')][:60] for row in rows]
    print('kernels:', names)
    for i, h in enumerate(hdr):
        if h in WANT or 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
            vals = [row[i] for row in rows]
            try:
                if all(float(v.replace(',', '')) < 0.5 for v in vals) and 'stalled' in h:
                    continue
            except ValueError:
                pass
            print(f'{h:90s} {units[i]:10s} {vals}')


if __name__ == '__main__':
    main()
