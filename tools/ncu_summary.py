"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name -> markdown table."""
import csv
import re
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
        ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1)
        name = re.sub(r'\(.*', '', r['Kernel Name'])
        rows.append((name, ns, r.get('Grid Size', ''), r.get('Block Size', '')))
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns, *_ in rows:
        agg[name][0] += 1
        agg[name][1] += ns
    total = sum(v[1] for v in agg.values())
    lines = [f'launches: {len(rows)}, total kernel time {total / 1e6:.3f} ms (cold-cache, serialised under ncu)', '',
             '| kernel | launches | total ms | share | avg us |', '|---|---:|---:|---:|---:|']
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f'| `{name[:110]}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.1f}% | {ns / n / 1e3:.1f} |')
    text = '\n'.join(lines)
    print(text)
    if out:
        with open(out, 'w') as f:
            f.write(text + '\n')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
