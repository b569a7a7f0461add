"""ncu launch list (csv from `--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]`)
-> per-kernel summary (markdown table).
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys

SCALE = {'ns': 1e-3, 'us': 1., 'ms': 1e3, 's': 1e6, 'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main(path):
  with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
  agg = collections.defaultdict(lambda: {'ids': set(), 't': 0., 'rd': 0., 'wr': 0.})
  for row in csv.DictReader(lines):
    try:
      v = float(row['Metric Value'].replace(',', '')) * SCALE.get(row['Metric Unit'], 1.)
    except (KeyError, ValueError):
      continue
    name = row['Kernel Name']
    m = re.search(r'(\w+_kernel)(<[^>]*>)?', name)
    key = (m.group(1) + (m.group(2) or '')) if m else name[:60]
    key = key.replace('<unnamed>::', '')
    a = agg[key]
    a['ids'].add(row['ID'])
    metric = row['Metric Name']
    if metric.startswith('gpu__time_duration'):
      a['t'] += v
    elif metric.startswith('dram__bytes_read'):
      a['rd'] += v
    elif metric.startswith('dram__bytes_write'):
      a['wr'] += v
  tot = sum(a['t'] for a in agg.values())
  dram = sum(a['rd'] + a['wr'] for a in agg.values())
  print(f'total GPU time of the profiled region: {tot / 1e3:.2f} ms (ncu serialised, cold cache: compare SHARES)')
  if dram:
    print(f'total DRAM traffic: {dram / 1e9:.2f} GB\n')
    print('| kernel | launches | total us | share | avg us | DRAM read MB | DRAM write MB | DRAM GB/s |')
    print('|---|---:|---:|---:|---:|---:|---:|---:|')
  else:
    print('\n| kernel | launches | total us | share | avg us |')
    print('|---|---:|---:|---:|---:|')
  for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['t']):
    n, t = len(a['ids']), a['t']
    line = f'| `{k}` | {n} | {t:.0f} | {100 * t / tot:.1f}% | {t / n:.1f} |'
    if dram:
      line += f" {a['rd'] / 1e6:.0f} | {a['wr'] / 1e6:.0f} | {(a['rd'] + a['wr']) / (t * 1e-6) / 1e9:.0f} |"
    print(line)


if __name__ == '__main__':
  main(sys.argv[1])
