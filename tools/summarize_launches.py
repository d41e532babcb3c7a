"""Per-kernel summary of an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X`).
    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/..._summary.md
Per-launch times under ncu are cold-cache and serialised: compare SHARES."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    tot = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '')
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[row['Metric Unit']]
        t = tot[name]
        t[0] += 1
        t[1] += v
        t[2] = max(t[2], v)
    total = sum(v[1] for v in tot.values())
    n = sum(v[0] for v in tot.values())
    print('# %s: %d launches, %.1f ms of kernel time\n' % (path, n, total))
    print('| kernel | launches | total ms | share | longest launch ms |')
    print('|---|---|---|---|---|')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.2f | %.1f %% | %.3f |' %
              (k[:90], v[0], v[1], 100 * v[1] / total, v[2]))


if __name__ == '__main__':
    main(sys.argv[1])
