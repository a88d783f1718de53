"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total ms, share.
Usage: python scripts/launch_summary.py launches.csv [steps_in_capture]"""
import csv, sys, re, collections
path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = []
with open(path, newline='') as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1)
    name = r['Kernel Name']
    name = re.sub(r'\(.*$', '', name)
    name = name.replace('void ', '').replace('fp::', '')
    rows.append((name[:70], ns))
agg = collections.OrderedDict()
for n, ns in rows:
    c, t = agg.get(n, (0, 0.0))
    agg[n] = (c + 1, t + ns)
tot = sum(t for _, t in agg.values())
print(f'{len(rows) / steps:.0f} launches/step, {tot / 1e6 / steps:.3f} ms/step (cold-cache, serialised)')
print('| kernel | launches/step | ms/step | share |\n|---|---|---|---|')
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{n}` | {c / steps:.0f} | {t / 1e6 / steps:.3f} | {100 * t / tot:.1f}% |')
