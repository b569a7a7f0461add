"""Per-shape GEMM time of ONE training step of the bench workload (CUDA events around every st_gemm launch):
which shapes the 38 ms of tensor-core time go to, and the TFLOP/s each one reaches inside the real step."""
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('ST_STEP_GRAPH', '0')      # the instrumented step must run eagerly (same kernels as the captured graphs)


def main():
  from soft_truncation_b200 import configs, losses, ops, sde_lib
  from soft_truncation_b200.models import utils as mutils
  from soft_truncation_b200.models.ema import ExponentialMovingAverage
  B = int(os.environ.get('GB_BATCH', '512'))
  cfg = configs.cifar10_ddpmpp_nll_st()
  cfg.device = torch.device('cuda:0')
  cfg.model.compute_dtype = 'bf16'
  torch.manual_seed(42)
  np.random.seed(42)
  sde = sde_lib.get_sde(cfg)
  model = mutils.create_model(cfg, sde)
  state = dict(model=model, optimizer=losses.get_optimizer(cfg, model.parameters()),
               ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)
  step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
  batch = torch.rand(B, 3, 32, 32, device=cfg.device) * 2 - 1
  for _ in range(3):
    step_fn(state, batch)
  recs, orig = [], ops._gemm

  def spy(**kw):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    orig(**kw)
    b.record()
    kind = {(1, 0): 'conv', (1, 2): 'dgrad', (0, 1): 'wgrad', (0, 0): 'gemm'}[(kw.get('a_mode', 0), kw.get('b_mode', 0))]
    taps = kw.get('kh', 1) * kw.get('kw', 1) if kind != 'gemm' else 1
    key = (kind, taps, kw.get('H', 0), kw['M'], kw['N'], kw['K'], kw.get('batch', 1), kw.get('split_k', 1),
           int(kw.get('residual') is not None), int(kw.get('C2', 0) > 0))
    recs.append((key, a, b, 2.0 * kw['M'] * kw['N'] * kw['K'] * kw.get('batch', 1)))
  ops._gemm = spy
  # head start for the host (45 ms of device spin): the eager launch stream stays ahead of the GPU, so every event pair
  # brackets a kernel that starts when its predecessor ends, as in the captured graph
  torch.cuda._sleep(int(0.045 * 1.7e9))
  step_fn(state, batch)
  torch.cuda.synchronize()
  ops._gemm = orig
  agg = collections.OrderedDict()
  for key, a, b, f in recs:
    e = agg.setdefault(key, [0, 0., 0.])
    e[0] += 1
    e[1] += a.elapsed_time(b)
    e[2] += f
  tot = sum(e[1] for e in agg.values())
  print(f'{len(recs)} st_gemm launches, {tot:.2f} ms, {sum(e[2] for e in agg.values()) / tot / 1e9:.0f} TFLOP/s average')
  print('kind taps H M N K batch split_k res cat | launches ms share TFLOP/s')
  for key, e in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(' '.join(str(k) for k in key), '|', e[0], f'{e[1]:.3f} {100 * e[1] / tot:.1f}% {e[2] / e[1] / 1e9:.0f}')


if __name__ == '__main__':
  main()
