"""Key metrics per kernel from an ncu report exported with `ncu -i X.ncu-rep --page raw --csv`."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sector_hit_rate.pct']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    want = WANT + [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and
                   h.endswith('per_issue_active.ratio')]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('kernel: %s' % d['Kernel Name'])
        for k in want:
            if k in d and d[k] not in ('', '0'):
                print('  %-88s %s %s' % (k, d[k], units[hdr.index(k)]))


if __name__ == '__main__':
    main(sys.argv[1])
