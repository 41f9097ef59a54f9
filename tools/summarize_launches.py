"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
usage: python tools/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/rNN_launches_summary.txt"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r'^void ', '', name)
    m = re.match(r'(paacb::)?([A-Za-z0-9_:]+)(<.*?>)?\(', name)
    base = name.split('(')[0]
    return base[:110]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            v = float(r['Metric Value'].replace(',', ''))
            unit = r.get('Metric Unit', 'ns')
            ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(unit, 1)
            rows.append((int(r['ID']), short(r['Kernel Name']), r['Grid Size'], r['Block Size'], ns))
    rows = rows[skip:]
    agg = OrderedDict()
    for _, k, grid, blk, ns in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print('# %s: %d launches (first %d skipped), total %.3f ms (cold-cache, serialised: compare shares, not absolutes)' % (
        path, len(rows), skip, tot / 1e6))
    print('%-112s %6s %12s %8s' % ('kernel', 'n', 'total_us', 'share'))
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-112s %6d %12.1f %7.2f%%' % (k, n, ns / 1e3, 100 * ns / tot))


if __name__ == '__main__':
    main()
