"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time share of
the LAST training step (delimited by the once-per-step sgd_kernel launches)."""
import csv
import sys
from collections import defaultdict


def main(path, out=None):
    rows = []
    with open(path, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(unit, 1)
        rows.append((r['Kernel Name'], ns))
    sgd = [i for i, (k, _) in enumerate(rows) if 'sgd_kernel' in k]
    if len(sgd) >= 2:
        rows = rows[sgd[-2] + 1: sgd[-1] + 1]
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for k, ns in rows:
        name = k.split('(')[0]
        if name.startswith('void '):
            name = name[5:]
        if len(name) > 90:
            name = name[:87] + '...'
        agg[name][0] += 1
        agg[name][1] += ns
    lines = [f'launches in step: {len(rows)}   summed kernel time: {tot / 1e6:.3f} ms',
             f'{"kernel":92s} {"n":>5s} {"ms":>9s} {"share":>7s}']
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f'{name:92s} {n:5d} {ns / 1e6:9.3f} {100 * ns / tot:6.1f}%')
    text = '\n'.join(lines)
    print(text)
    if out:
        with open(out, 'w') as f:
            f.write(text + '\n')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
