"""ncu launch list (csv from `--metrics gpu__time_duration.sum`) -> per-kernel summary (markdown table).
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys


def main(path):
  with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
  agg = collections.defaultdict(lambda: [0, 0.0])
  tot = 0.
  for row in csv.DictReader(lines):
    try:
      v = float(row['Metric Value'].replace(',', ''))
    except (KeyError, ValueError):
      continue
    unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    name = row['Kernel Name']
    m = re.search(r'(\w+_kernel)(<[^>]*>)?', name)
    key = (m.group(1) + (m.group(2) or '')) if m else name[:60]
    key = key.replace('<unnamed>::', '')
    agg[key][0] += 1
    agg[key][1] += v
    tot += v
  print(f'total GPU time of the profiled region: {tot / 1e3:.2f} ms (ncu serialised, cold cache: compare SHARES)\n')
  print('| kernel | launches | total us | share | avg us |')
  print('|---|---:|---:|---:|---:|')
  for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{k}` | {n} | {t:.0f} | {100 * t / tot:.1f}% | {t / n:.1f} |')


if __name__ == '__main__':
  main(sys.argv[1])
