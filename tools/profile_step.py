"""One training step (and optionally one sampler step) of the bench workload between cudaProfilerStart/Stop,
for `ncu --profile-from-start off ...` (see /opt/skills/guides/B200_PROFILING.md)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('ST_STEP_GRAPH', '0')      # profile the eager step: the captured graphs replay the same kernels


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--batch', type=int, default=512)
  ap.add_argument('--dtype', default='bf16')
  ap.add_argument('--warm', type=int, default=2)
  ap.add_argument('--mode', default='train', choices=['train', 'forward'])
  ap.add_argument('--config', default='vp/CIFAR10/ddpmpp_nll_st', help='config path under the reference configs/ tree')
  args = ap.parse_args()
  from soft_truncation_b200 import configs, losses, sde_lib
  from soft_truncation_b200.models import utils as mutils
  from soft_truncation_b200.models.ema import ExponentialMovingAverage
  cfg = configs.get_config(args.config)
  cfg.device = torch.device('cuda:0')
  cfg.model.compute_dtype = args.dtype
  torch.manual_seed(42)
  np.random.seed(42)
  sde = sde_lib.get_sde(cfg)
  model = mutils.create_model(cfg, sde)
  state = dict(model=model, optimizer=losses.get_optimizer(cfg, model.parameters()),
               ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)
  step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
  R = cfg.data.image_size
  batch = torch.rand(args.batch, 3, R, R, device=cfg.device)
  if cfg.data.centered:
    batch = batch * 2 - 1
  t = torch.rand(args.batch, device=cfg.device) * 999

  def run():
    if args.mode == 'train':
      step_fn(state, batch)
    else:
      with torch.no_grad():
        model.eval()
        model(batch, t)

  for _ in range(args.warm):
    run()
  torch.cuda.synchronize()
  torch.cuda.profiler.start()
  run()
  torch.cuda.synchronize()
  torch.cuda.profiler.stop()


if __name__ == '__main__':
  main()
