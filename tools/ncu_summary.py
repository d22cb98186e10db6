"""Condense `ncu -i report.ncu-rep --page raw --csv` into the side-by-side table format used under profiles/."""
import csv
import sys

KEEP = ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second')


def main(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith('==')))
    header, units, data = rows[0], rows[1], rows[2:]
    name_i = header.index('Kernel Name')
    print('%-90s %-16s %s' % ('Kernel Name', '', ' | '.join(r[name_i][:70] for r in data)))
    for key in ('Block Size', 'Grid Size'):
        i = header.index(key)
        print('%-90s %-16s %s' % (key, '', ' | '.join(r[i] for r in data)))
    for i, h in enumerate(header):
        if h in KEEP:
            print('%-90s %-16s %s' % (h, units[i], ' | '.join(r[i] for r in data)))


if __name__ == '__main__':
    main(sys.argv[1])
