"""Summarise an ncu launch list (gpu__time_duration.sum per launch) down to the LAST train step of the run:
from the last conv3x3_first_kernel launch to the last sgd_momentum_kernel launch."""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    out = []
    for row in r:
        v = float(row[vi].replace(',', ''))
        v = v / 1000 if row[ui] == 'ns' else (v * 1000 if row[ui] == 'ms' else v)
        out.append((row[ki], v))
    return out


def main(path):
    seq = load(path)
    starts = [i for i, (n, _) in enumerate(seq) if 'im2col_first_kernel' in n or 'conv3x3_first_kernel' in n]
    ends = [i for i, (n, _) in enumerate(seq) if 'sgd_momentum_kernel' in n or 'adam_kernel' in n]
    if starts and ends and ends[-1] > starts[-1]:
        seq = seq[starts[-1]:ends[-1] + 1]
    agg = collections.OrderedDict()
    for n, v in seq:
        k = n.split('(')[0].replace('void ', '').replace('hk::', '')[-70:]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in seq)
    print(f'last step: {len(seq)} launches, {tot / 1000:.3f} ms summed kernel time (cold-cache, serialised)')
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{v:10.1f} us {100 * v / tot:5.1f}%  n={n:3d}  {k}')


if __name__ == '__main__':
    main(sys.argv[1])
