"""Key metrics per kernel from an ncu report exported with `ncu -i X.ncu-rep --page raw --csv`.

    python tools/ncu_summary.py raw.csv                          # text summary of every kernel in the export
    python tools/ncu_summary.py raw.csv --json profiles/hot_kernel_metrics.json --kernel group_rows \
        --source profiles/r2x_ncu_summary.txt --psfs 2240        # the file bench.py reads for roofline.*

The JSON holds the figures of the dominant kernel that cannot be measured live by bench.py (DRAM
bytes per launch, shared-memory wavefront / FP64 pipe / issue utilisation) and the bound they imply.
"""
import argparse
import csv
import json
import re

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

UNIT_SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'usecond': 1e-6,
              'ms': 1e-3, 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1.0, 's': 1.0}


def rows_of(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    return hdr, units, [dict(zip(hdr, r)) for r in rows[2:]]


def num(d, units, hdr, key):
    """value of metric `key` in SI base units (bytes, seconds) or as the plain number"""
    v = d.get(key, '')
    if v == '':
        return None
    x = float(v.replace(',', ''))
    return x * UNIT_SCALE.get(units[hdr.index(key)], 1.0)


def summary(path):
    hdr, units, rows = rows_of(path)
    want = WANT + [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and
                   h.endswith('per_issue_active.ratio')]
    for d in rows:
        print('kernel: %s' % d['Kernel Name'])
        for k in want:
            if k in d and d[k] not in ('', '0'):
                print('  %-88s %s %s' % (k, d[k], units[hdr.index(k)]))


def to_json(path, out, kernel, source, psfs):
    hdr, units, rows = rows_of(path)
    pick = [d for d in rows if re.search(kernel, d['Kernel Name'])]
    if not pick:
        raise SystemExit('no kernel matches %r' % kernel)
    d = pick[-1]
    get = lambda k: num(d, units, hdr, k)   # noqa: E731
    rd, wr, dur = get('dram__bytes_read.sum'), get('dram__bytes_write.sum'), get('gpu__time_duration.sum')
    fracs = {'smem_wavefront_frac': get('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
             'fp64_pipe_frac': get('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'),
             'issue_frac': get('smsp__issue_active.avg.pct_of_peak_sustained_active'),
             'dram_frac_under_ncu': get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')}
    fracs = {k: (v / 100.0 if v is not None else None) for k, v in fracs.items()}
    names = {'smem_wavefront_frac': 'shared-memory wavefronts', 'fp64_pipe_frac': 'FP64 pipe', 'issue_frac': 'issue slots',
             'dram_frac_under_ncu': 'hbm'}
    top = max((k for k in fracs if fracs[k] is not None), key=lambda k: fracs[k])
    res = {'kernel': d['Kernel Name'].split('(')[0], 'bound': names[top],
           'dram_bytes_per_launch': (rd or 0) + (wr or 0), 'dram_read_bytes': rd, 'dram_write_bytes': wr,
           'duration_under_ncu_ms': dur * 1e3 if dur else None, 'psfs_per_launch': psfs,
           'smem_wavefronts': get('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'),
           'smem_bank_conflicts': get('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
           'l2_hit_rate': (get('lts__t_sector_hit_rate.pct') or 0) / 100.0,
           'registers_per_thread': get('launch__registers_per_thread'),
           'warps_per_scheduler': get('smsp__warps_active.avg.per_cycle_active'),
           'source': source}
    res.update(fracs)
    with open(out, 'w') as f:
        json.dump(res, f, indent=1)
        f.write('\n')
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--json')
    ap.add_argument('--kernel', default='group_rows')
    ap.add_argument('--source', default='')
    ap.add_argument('--psfs', type=int, default=2240)
    a = ap.parse_args()
    if a.json:
        to_json(a.csv, a.json, a.kernel, a.source, a.psfs)
    else:
        summary(a.csv)
