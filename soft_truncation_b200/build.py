"""Builds libst_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch headers).

    python -m soft_truncation_b200.build [--force]

The library has a pure C ABI (include/st_b200.h); it links only the static CUDA runtime and
resolves the one driver entry point it needs (cuTensorMapEncodeTiled) at run time.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(PKG, 'build')
LIB = os.path.join(PKG, 'libst_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
         '-Xcompiler', '-fvisibility=hidden', '-I', os.path.join(ROOT, 'include'), '-I', CSRC,
         '--expt-relaxed-constexpr', '-Xptxas', '-v']


def sources():
  return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
  h = hashlib.sha256(' '.join(FLAGS).encode())
  for f in sorted(os.listdir(CSRC)) + ['../../include/st_b200.h']:
    with open(os.path.join(CSRC, f), 'rb') as fh:
      h.update(f.encode())
      h.update(fh.read())
  return h.hexdigest()


def _compile(src):
  obj = os.path.join(OBJ, src[:-3] + '.o')
  cmd = [NVCC, *FLAGS, '-c', os.path.join(CSRC, src), '-o', obj]
  r = subprocess.run(cmd, capture_output=True, text=True)
  log = r.stdout + r.stderr
  with open(obj + '.log', 'w') as fh:
    fh.write(' '.join(cmd) + '\n' + log)
  if r.returncode != 0:
    raise RuntimeError(f'nvcc failed on {src}:\n{log[-6000:]}')
  return obj


def build(force=False, verbose=False):
  os.makedirs(OBJ, exist_ok=True)
  stamp = os.path.join(OBJ, 'digest')
  dig = _digest()
  if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
    return LIB
  with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
    objs = list(ex.map(_compile, sources()))
  cmd = [NVCC, '-shared', '-o', LIB, *objs, '-cudart', 'static', '-Xlinker', '--exclude-libs=ALL']
  r = subprocess.run(cmd, capture_output=True, text=True)
  if r.returncode != 0:
    raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
  with open(stamp, 'w') as fh:
    fh.write(dig)
  if verbose:
    print('built', LIB)
  return LIB


if __name__ == '__main__':
  build(force='--force' in sys.argv, verbose=True)
