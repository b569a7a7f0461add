"""HBM throughput of st_upfirdn2d (the replacement of op/upfirdn2d_kernel.cu) on the three hot call shapes of the FIR
configs: upsample_2d (up 2, pad (2,1), 4x4 taps), downsample_2d (down 2, pad (1,1)) and the
conv_downsample_2d pre-filter (1:1, pad (2,2)), on C3 (CelebA-64, B=128) and C5 (CelebA-HQ-256, B=16) tensors, bf16 and
fp32.  GB/s = (input + output bytes) / CUDA-event time, inputs larger than L2 or L2 flushed between launches; the gate of
SURVEY 7.1 step 3 is >= 80 % of the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soft_truncation_b200 import ops  # noqa: E402

DEV = torch.device('cuda:0')


def fir(gain):
  k = np.outer([1., 3., 3., 1.], [1., 3., 3., 1.]).astype(np.float32)
  return torch.tensor(k / k.sum() * gain, device=DEV)


def main():
  peak = 6550.4
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    peak = json.load(open(p)).get('hbm_gbs', peak)
  flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
  cases = [('C3 res-block up   32->64', 128, 32, 128, dict(up=2, pad=(2, 1)), 4.),
           ('C3 res-block down 64->32', 128, 64, 128, dict(down=2, pad=(1, 1)), 1.),
           ('C3 pre-filter     64->64', 128, 64, 128, dict(pad=(2, 2)), 1.),
           ('C5 res-block up  128->256', 16, 128, 128, dict(up=2, pad=(2, 1)), 4.),
           ('C5 res-block down 256->128', 16, 256, 128, dict(down=2, pad=(1, 1)), 1.),
           ('C5 pyramid image down 256->128 (64-ch padded)', 16, 256, 64, dict(down=2, pad=(1, 1)), 1.)]
  print(f'st_upfirdn2d, NHWC, L2 flushed between launches; copy peak {peak:.0f} GB/s (MEASURED_PEAKS.json)')
  for name, B, H, C, kw, gain in cases:
    k = fir(gain)
    for dtype in (torch.bfloat16, torch.float32):
      x = torch.randn(B, H, H, C, device=DEV).to(dtype)
      y = ops.upfirdn2d_nhwc(x, k, **kw)
      ts = []
      for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.upfirdn2d_nhwc(x, k, **kw)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
      t = float(np.median(ts)) * 1e-3
      nbytes = x.numel() * x.element_size() + y.numel() * y.element_size()
      print(f'{name:48s} {str(dtype)[6:]:9s} {t * 1e6:8.1f} us  {nbytes / 1e6:8.1f} MB  {nbytes / t / 1e9:7.0f} GB/s  {nbytes / t / 1e9 / peak:5.2f} of copy peak')


if __name__ == '__main__':
  main()
