"""Runs one GEMM case of the bench workload a few times (for `ncu --set full --import-source on`).
GP_CASE in {fwd128, fwd256, dgrad128, wgrad128, attn_qk, attn_pv, nin768, nin256}; kernel variant / pair form via
ST_TC_VARIANT / ST_TC_CG."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from soft_truncation_b200 import ops  # noqa: E402

DEV, BF = torch.device('cuda:0'), torch.bfloat16
case = os.environ.get('GP_CASE', 'fwd128')
B = int(os.environ.get('GB_BATCH', '512'))
if case in ('attn_qk', 'attn_pv', 'nin768', 'nin256'):
  T = 256                                    # tokens per image at 16x16, C = 256
  qkv = torch.randn(B, T, 768, device=DEV).to(BF)
  h = torch.randn(B * T, 256, device=DEV).to(BF)
  w = (torch.randn(768, 256, device=DEV) * 0.05).to(BF)
  bias = torch.randn(768, device=DEV)
  s_ = torch.empty(B, T, T, device=DEV, dtype=BF)
  s32 = torch.empty(B, T, T, device=DEV, dtype=torch.float32)     # the model keeps the logits in fp32
  o_ = torch.empty(B, T, 256, device=DEV, dtype=BF)
  for _ in range(5):
    if case == 'attn_qk':                    # S = Q K^T: q, k are column slices of the packed qkv rows
      ops.gemm_nt(qkv, qkv[:, :, 256:], out=s32, M=T, N=T, K=256, batch=B, lda=768, ldb=768, sAb=T * 768, sBb=T * 768,
                  sCb=T * T, alpha=256 ** -0.5)
    elif case == 'attn_pv':                  # O = P V
      ops.gemm_nn(s_, qkv[:, :, 512:], 256, out=o_, M=T, K=T, batch=B, lda=T, ldb=768, sAb=T * T, sBb=T * 768, sCb=T * 256)
    elif case == 'nin768':
      ops.gemm_nt(h, w, bias=bias)
    else:
      ops.gemm_nt(h, w[:256], bias=bias[:256], residual=h)
  torch.cuda.synchronize()
  sys.exit(0)
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
