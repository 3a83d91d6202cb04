"""Per-source-line hot spots of one kernel from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv
--kernel-name regex:NAME` (a report captured with `--import-source on`, library built with -lineinfo).

    python tools/ncu_source_summary.py src.csv [--top 40] > profiles/r2x_row_kernel_source_hotspots.txt

Prints, per CUDA source line, the warp-stall samples (all / not issued), the share of the kernel's samples, the
instructions executed and the dominant stall reasons - the evidence behind "where does a transform's time go".
"""
import argparse
import csv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--top', type=int, default=40)
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    path, hdr, lines = '', None, []
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            path = r[1].split('/')[-1]
        elif r and r[0] == 'Line No':
            hdr = r
        elif hdr and r and r[0].strip().isdigit() and len(r) >= len(hdr):
            d = dict(zip(hdr[4:], r[4:]))      # columns 0-3 are (line, source, address, sass)
            lines.append((path, int(r[0]), r[1].strip(), d))

    def num(d, k):
        try:
            return float(d.get(k, '0').replace(',', ''))
        except ValueError:
            return 0.0
    tot = sum(num(d, 'Warp Stall Sampling (All Samples)') for *_, d in lines) or 1.0
    tot_inst = sum(num(d, 'Instructions Executed') for *_, d in lines) or 1.0
    stall_keys = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
    print('kernel samples %d, warp instructions %d (source lines with samples: %d)' % (tot, tot_inst, len(lines)))
    print('%6s %6s %7s %7s  %-22s %-34s %s' % ('share', 'cum', 'samples', 'inst %', 'file:line', 'top stall reasons', 'source'))
    cum = 0.0
    for path, no, src, d in sorted(lines, key=lambda t: -num(t[3], 'Warp Stall Sampling (All Samples)'))[:a.top]:
        s = num(d, 'Warp Stall Sampling (All Samples)')
        cum += s
        top = sorted(((num(d, k), k[6:]) for k in stall_keys), reverse=True)[:3]
        why = ' '.join('%s %.0f%%' % (k, 100 * v / s) for v, k in top if v > 0 and s > 0)
        print('%5.1f%% %5.1f%% %7d %6.1f%%  %-22s %-34s %s' % (100 * s / tot, 100 * cum / tot, s,
              100 * num(d, 'Instructions Executed') / tot_inst, '%s:%d' % (path, no), why, src[:90]))


if __name__ == '__main__':
    main()
