"""Runs one GEMM case of the bench workload a few times (for `ncu --set full --import-source on`).
GP_CASE in {fwd128, fwd256, dgrad128, wgrad128}; kernel variant / cluster via ST_TC_VARIANT / ST_TC_CLUSTER."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soft_truncation_b200 import ops  # noqa: E402

DEV, BF = torch.device('cuda:0'), torch.bfloat16
case = os.environ.get('GP_CASE', 'fwd128')
B = int(os.environ.get('GB_BATCH', '512'))
H, C, Co = (32, 128, 128) if case.endswith('128') else (16, 256, 256)
x = torch.randn(B, H, H, C, device=DEV).to(BF)
w = (torch.randn(Co, 9 * C, device=DEV) * 0.02).to(BF)
bias, rb = torch.randn(Co, device=DEV), torch.randn(B, Co, device=DEV)
dy = torch.randn(B, H, H, Co, device=DEV).to(BF)
dw = torch.zeros(Co, 9 * C, device=DEV)
out = torch.empty(B, H, H, Co, device=DEV, dtype=BF)
for _ in range(5):
  if case.startswith('fwd'):
    ops.conv_fwd(x, w, Co, bias=bias, rowbias=rb, rowbias_ld=Co, out=out)
  elif case.startswith('dgrad'):
    ops.conv_dgrad(dy, w, C, out=out)
  else:
    ops.conv_wgrad(dy, x, dw)
torch.cuda.synchronize()
