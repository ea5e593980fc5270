"""Summarise an `ncu --page raw --csv` dump: python profiles/ncu_summary.py raw.csv [pattern ...]"""
import csv
import sys

DEFAULT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__t_sectors.sum', 'lts__t_sector_hit_rate.pct',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_sector_hit_rate.pct',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size',
    'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
    'sm__inst_executed_pipe_fp64.sum',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
    'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum',
    'lts__t_sectors_srcunit_tex_op_read.sum',
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    pats = sys.argv[2:]
    names = DEFAULT if not pats else [h for h in hdr if any(p in h for p in pats)]
    kcol = hdr.index('Kernel Name')
    print('kernels:', [r[kcol][:60] for r in rows[2:]])
    for w in names:
        if w in hdr:
            i = hdr.index(w)
            print('%-72s %-10s %s' % (w, units[i], [r[i] for r in rows[2:]]))
        else:
            print('%-72s (absent)' % w)


if __name__ == '__main__':
    main()
