"""Key metrics per kernel from `ncu -i X.ncu-rep --page raw --csv` exports.
    python tools/ncu_summary.py gpurun_out/r2_rest_raw.csv [...] > profiles/..._summary.md"""
import csv
import sys

KEYS = [
    ('gpu__time_duration.sum', 'ms'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1/shared %'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
]


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9,
                'Tbyte': 1e12}.get(unit, 1)


def main(paths):
    print('| kernel | ms | DRAM read GB | DRAM write GB | DRAM % | L1/shared % '
          '| L2 hit % | issue % | warps active % | regs | grid |')
    print('|' + '---|' * 11)
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            u = dict(zip(hdr, units))
            out = [d['Kernel Name'].replace('void ', '')[:60]]
            for k, _ in KEYS:
                if k not in d:
                    out.append('-')
                elif 'bytes' in k:
                    out.append('%.3f' % (to_bytes(d[k], u[k]) / 1e9))
                elif k == 'gpu__time_duration.sum':
                    v = float(d[k].replace(',', ''))
                    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3}[u[k]]
                    out.append('%.3f' % v)
                else:
                    try:
                        out.append('%.1f' % float(d[k].replace(',', '')))
                    except ValueError:
                        out.append(d[k])
            print('| ' + ' | '.join(out) + ' |')


if __name__ == '__main__':
    main(sys.argv[1:])
