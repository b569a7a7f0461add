"""Opcode histogram of every kernel in the built library (cuobjdump -sass), the evidence that the hot kernels are
Blackwell-native: tcgen05.mma -> UTCHMMA(.2CTA), tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG / UTMAREDG, cp.async -> LDGSTS,
clusters -> UCGABAR, mbarrier -> SYNCS.  usage: python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'soft_truncation_b200', 'libst_b200.so')
KEY = re.compile(r'^(UTC|UTMA|LDTM|STTM|LDGSTS|UCGABAR|SYNCS|HMMA|HGMMA|MUFU|UBLKCP|REDG|ATOMG|RED\b|FFMA|LDS|STS|LDG|STG|SHFL|BAR|ELECT|ACQBULK|CCTL)')


def main():
  out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
  per = collections.OrderedDict()
  name = None
  for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
      name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
      name = re.sub(r'\(anonymous namespace\)::', '', name)
      name = re.sub(r'\(.*$', '', name)
      per[name] = collections.Counter()
      continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
    if m and name:
      per[name][m.group(1)] += 1
  print(f'# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (sm_100a), selected opcode classes + instruction totals')
  print('# tcgen05.mma = UTCHMMA (.2CTA = cta_group::2), tcgen05.ld = LDTM, tcgen05.commit = UTCBAR, tcgen05.alloc = UTCATOMSWS,')
  print('# TMA load / store / reduce = UTMALDG / UTMASTG / UTMAREDG, cp.async = LDGSTS, cluster barrier = UCGABAR, mbarrier = SYNCS\n')
  for fn, c in per.items():
    total = sum(c.values())
    sel = {k: v for k, v in c.items() if KEY.match(k)}
    agg = collections.Counter()
    for k, v in sel.items():
      base = k if k.startswith(('UTC', 'UTMA', 'LDTM', 'UCGABAR', 'UBLKCP')) else k.split('.')[0]
      agg[base] += v
    print(f'{fn}  [{total} instructions]')
    print('    ' + ', '.join(f'{k} {v}' for k, v in sorted(agg.items(), key=lambda kv: (-kv[1], kv[0]))))


if __name__ == '__main__':
  main()
