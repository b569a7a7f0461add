"""Is an MN-major B operand slower than a K-major one on the same GEMM?  (decides whether conv_dgrad should read a
transposed copy of the weights instead of walking [Co][tap][Ci] MN-major)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soft_truncation_b200 import ops  # noqa: E402

DEV, BF = torch.device('cuda:0'), torch.bfloat16


def timeit(fn, iters=20):
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


for (M, N, K) in [(131072, 256, 2304), (524288, 128, 1152), (32768, 256, 2304)]:
  a = torch.randn(M, K, device=DEV).to(BF)
  bk = (torch.randn(N, K, device=DEV) * 0.02).to(BF)       # [N][K]  K-major
  bn = bk.t().contiguous()                                 # [K][N]  MN-major
  out = torch.empty(M, N, device=DEV, dtype=BF)
  fl = 2.0 * M * N * K
  t_nt = timeit(lambda: ops.gemm_nt(a, bk, out=out))
  t_nn = timeit(lambda: ops.gemm_nn(a, bn, N, out=out))
  print(f'M={M} N={N} K={K}: B K-major {fl / t_nt / 1e12:7.1f} TFLOP/s   B MN-major {fl / t_nn / 1e12:7.1f} TFLOP/s', flush=True)
B = 512
for (H, C, Co) in [(16, 256, 256), (32, 128, 128), (8, 256, 256)]:
  x = torch.randn(B, H, H, C, device=DEV).to(BF)
  w = (torch.randn(Co, 9 * C, device=DEV) * 0.02).to(BF)
  dy = torch.randn(B, H, H, Co, device=DEV).to(BF)
  o1 = torch.empty(B, H, H, Co, device=DEV, dtype=BF)
  o2 = torch.empty(B, H, H, C, device=DEV, dtype=BF)
  fl = 2.0 * B * H * H * Co * 9 * C
  t_f = timeit(lambda: ops.conv_fwd(x, w, Co, out=o1))
  t_d = timeit(lambda: ops.conv_dgrad(dy, w, C, out=o2))
  # dgrad expressed as a forward conv over dy with [Ci][reversed tap][Co] weights (values irrelevant for timing)
  wt = (torch.randn(C, 9 * Co, device=DEV) * 0.02).to(BF)
  t_t = timeit(lambda: ops.conv_fwd(dy, wt, C, out=o2))
  print(f'conv {C}->{Co} @{H}: fwd {fl / t_f / 1e12:7.1f}  dgrad {fl / t_d / 1e12:7.1f}  dgrad-as-fwd(transposed weights) {fl / t_t / 1e12:7.1f}', flush=True)
