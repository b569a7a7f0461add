"""Times the GEMM / implicit-GEMM shapes of the bench workload (DDPM++ CIFAR-10, batch 512) one by one with CUDA
events, for each tcgen05 kernel variant / cluster size, and prints TFLOP/s per shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soft_truncation_b200 import ops  # noqa: E402

DEV = torch.device('cuda:0')
BF = torch.bfloat16


def timeit(fn, iters=10):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(iters):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / iters * 1e-3


def conv_shapes(B):
  # (name, H, C1, C2, Cout, k)
  return [('c3 128->128 @32', 32, 128, 0, 128, 3), ('c3 256->128 @32', 32, 128, 128, 128, 3),
          ('c3 256->256 @16', 16, 256, 0, 256, 3), ('c3 512->256 @16', 16, 256, 256, 256, 3),
          ('c3 256->256 @8', 8, 256, 0, 256, 3), ('c3 512->256 @4', 4, 256, 256, 256, 3),
          ('c1 256->128 @32', 32, 128, 128, 128, 1), ('c3 64->128 @32', 32, 64, 0, 128, 3),
          ('c3 128->64 @32', 32, 128, 0, 64, 3)]


def main():
  B = int(os.environ.get('GB_BATCH', '512'))
  # single-CTA tcgen05.mma (round 1) against CTA pairs (cta_group::2), with and without programmatic dependent launch
  # single-CTA tcgen05.mma (round 1) / CTA pairs (cta_group::2) / CTA pairs + halo form of the 3x3 convolutions
  configs = [('cg1', dict(ST_TC_CG='1', ST_TC_HALO='0')), ('cg2', dict(ST_TC_CG='2', ST_TC_HALO='0')),
             ('cg2 halo', dict(ST_TC_CG='2', ST_TC_HALO='1'))]
  rows = []
  for name, H, C1, C2, Co, k in conv_shapes(B):
    Ci = C1 + C2
    x1 = torch.randn(B, H, H, C1, device=DEV).to(BF)
    x2 = torch.randn(B, H, H, C2, device=DEV).to(BF) if C2 else None
    w = (torch.randn(Co, k * k * Ci, device=DEV) * 0.02).to(BF)
    bias = torch.randn(Co, device=DEV)
    rb = torch.randn(B, Co, device=DEV)
    res = torch.randn(B, H, H, Co, device=DEV).to(BF)
    dy = torch.randn(B, H, H, Co, device=DEV).to(BF)
    dw = torch.zeros(Co, k * k * Ci, device=DEV)
    wt = (torch.randn(Ci, k * k * Co, device=DEV) * 0.02).to(BF)      # transposed weight copy: the bf16 path's data gradient
    out = torch.empty(B, H, H, Co, device=DEV, dtype=BF)
    dx = torch.empty(B, H, H, Ci, device=DEV, dtype=BF)
    flops = 2.0 * B * H * H * Co * k * k * Ci
    cases = {
        'fwd+rowbias': lambda: ops.conv_fwd(x1, w, Co, k, k, x2=x2, bias=bias, rowbias=rb, rowbias_ld=Co, out=out),
        'fwd+residual': lambda: ops.conv_fwd(x1, w, Co, k, k, x2=x2, bias=bias, residual=res, alpha=0.7, out=out),
        'dgrad': lambda: ops.conv_dgrad(dy, w, Ci, k, k, out=dx),
        'dgrad (as conv)': lambda: ops.conv_fwd(dy, wt, Ci, k, k, out=dx),
        'wgrad': lambda: ops.conv_wgrad(dy, x1, dw, k, k, x2=x2),
    }
    for cname, fn in cases.items():
      line = [f'{name:18s} {cname:15s}']
      for vname, env in configs:
        os.environ.update(env)
        t = timeit(fn)
        line.append(f'{vname} {flops / t / 1e12:7.1f}')
      rows.append('  '.join(line))
      print(rows[-1], flush=True)
  # attention-shaped GEMMs (C = 256, L = 256)
  C, L = 256, 256
  npix = B * L
  h = torch.randn(npix, C, device=DEV).to(BF)
  wqkv = (torch.randn(3 * C, C, device=DEV) * 0.05).to(BF)
  qkv = torch.randn(npix, 3 * C, device=DEV).to(BF)
  p = torch.softmax(torch.randn(B, L, L, device=DEV), -1).to(BF)
  cases = {
      'qkv proj': (2.0 * npix * 3 * C * C, lambda: ops.gemm_nt(h, wqkv)),
      'logits': (2.0 * B * L * L * C, lambda: ops.gemm_nt(qkv, qkv[:, C:], out_dtype=torch.float32, M=L, N=L, K=C, lda=3 * C,
                                                          ldb=3 * C, batch=B, sAb=L * 3 * C, sBb=L * 3 * C, sCb=L * L)),
      'p @ v': (2.0 * B * L * L * C, lambda: ops.gemm_nn(p, qkv[:, 2 * C:], C, M=L, K=L, lda=L, ldb=3 * C, batch=B, sAb=L * L,
                                                         sBb=L * 3 * C, sCb=L * C)),
  }
  for cname, (flops, fn) in cases.items():
    line = [f'attention          {cname:13s}']
    for vname, env in configs:
      os.environ.update(env)
      t = timeit(fn)
      line.append(f'{vname} {flops / t / 1e12:7.1f}')
    print('  '.join(line), flush=True)


if __name__ == '__main__':
  main()
