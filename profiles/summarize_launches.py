"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total time, share).
    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.txt"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith('==')))
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3}.get(r[ui], 1.0)
        name = r[ki].split('(')[0].replace('void ', '').replace('vqs::<unnamed>::', 'vqs::')
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print('# %s: %d launches, %.1f us total (cold-cache, serialised: compare SHARES)' % (path, sum(v[0] for v in agg.values()), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-72s n=%4d %11.1f us %6.2f%%' % (k[:72], v[0], v[1], 100 * v[1] / tot))


if __name__ == '__main__':
    main(sys.argv[1])
