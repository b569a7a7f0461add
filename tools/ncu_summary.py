"""Key metrics of every launch in an `ncu --set full` report, as a markdown table (what profiles/*_ncu_*.md quote).

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/x.csv      # needs only the ncu CLI, no GPU
    python tools/ncu_summary.py /tmp/x.csv

or directly `python tools/ncu_summary.py gpurun_out/x.ncu-rep` (runs the ncu export itself).
"""
import csv
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM % of peak'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM % of peak'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('smsp__issue_active.avg.pct', 'issue slots busy %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU (MUFU) pipe %'),
    ('sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
    ('launch__registers_per_thread', 'registers'),
    ('launch__occupancy_limit_shared_mem', 'CTAs/SM (smem limit)'),
    ('launch__occupancy_limit_registers', 'CTAs/SM (register limit)'),
    ('launch__cluster_size', 'cluster size'),
]


def rows_of(path):
  if path.endswith('.ncu-rep'):
    text = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    lines = text.splitlines()
  else:
    lines = open(path).read().splitlines()
  lines = [ln for ln in lines if not ln.startswith('==')]
  return list(csv.reader(lines))


def main(path):
  rows = rows_of(path)
  hdr, units, data = rows[0], rows[1], rows[2:]
  idx = {h: i for i, h in enumerate(hdr)}
  stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
  for n, r in enumerate(data):
    name = r[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    print(f'### launch {n}: `{name}` grid {r[idx["Grid Size"]]} block {r[idx["Block Size"]]}\n')
    print('| metric | value |\n|---|---:|')
    for key, label in METRICS:
      if key in idx and r[idx[key]] != '':
        print(f'| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |')
    # tensor-pipe activity: the metric's name differs between ncu versions / architectures (tcgen05 on sm_100)
    for h in hdr:
      if 'pipe_tensor' in h and 'cycles_active' in h and ('pct' in h or h.endswith('.avg')) and r[idx[h]] != '':
        print(f'| tensor pipe (`{h}`) | {r[idx[h]]} {units[idx[h]]} |')
    top = sorted(((float(r[idx[h]] or 0), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')])
                  for h in stall), reverse=True)[:5]
    print('| top stall reasons (warps per issue) | ' + ', '.join(f'{k} {v:.2f}' for v, k in top) + ' |\n')


if __name__ == '__main__':
  main(sys.argv[1])
