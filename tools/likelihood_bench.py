"""Bits/dim evaluation rate of likelihood.get_likelihood_fn on the full-size CIFAR-10 DDPM++ (bf16, random weights):
network evaluations, wall time, and the cost of one ODE right-hand side (forward + input-gradient-only backward)
against the same right-hand side with the full backward pass.

    python tools/likelihood_bench.py [batch] [rtol]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
  from soft_truncation_b200 import configs, datasets, likelihood, ops, sde_lib
  from soft_truncation_b200.models import utils as mutils
  B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
  rtol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
  dev = torch.device('cuda:0')
  cfg = configs.cifar10_ddpmpp_nll_st()
  cfg.device = dev
  cfg.model.compute_dtype = 'bf16'
  torch.manual_seed(42)
  np.random.seed(42)
  sde = sde_lib.get_sde(cfg)
  model = mutils.create_model(cfg, sde)
  model.eval()
  data = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
  inv = datasets.get_data_inverse_scaler(cfg)
  out = {'batch': B, 'rtol': rtol}

  # one right-hand side: drift + Hutchinson divergence
  score_fn = mutils.get_score_fn(cfg, sde, model, train=False, continuous=True)
  rsde = sde.reverse(score_fn, probability_flow=True, lambda_=0.)
  eps = torch.randint_like(data, 0, 2).float() * 2 - 1
  t = torch.full((B,), 0.5, device=dev)

  def rhs(input_only):
    with torch.enable_grad():
      xg = data.detach().requires_grad_(True)
      drift = rsde.sde(xg, t)[0]
      if input_only:
        with ops.input_grads_only():
          g = torch.autograd.grad(torch.sum(drift * eps), xg)[0]
      else:
        g = torch.autograd.grad(torch.sum(drift * eps), xg)[0]
    return g

  for mode in (True, False):
    for _ in range(2):
      rhs(mode)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
      rhs(mode)
    b.record()
    torch.cuda.synchronize()
    out['rhs_ms_input_grads_only' if mode else 'rhs_ms_full_backward'] = a.elapsed_time(b) / 5

  for solver in ('device', 'scipy'):
    fn = likelihood.get_likelihood_fn(cfg, sde, inv, rtol=rtol, atol=rtol, solver=solver)
    torch.manual_seed(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bpd, z, nfe = fn(model, data, eps=1e-3)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert torch.isfinite(bpd).all()
    out[solver] = {'nfe': nfe, 'seconds': dt, 'images_per_s': B / dt, 'ms_per_nfe': 1e3 * dt / nfe, 'bpd_mean': float(bpd.mean())}
  print(json.dumps(out), flush=True)


if __name__ == '__main__':
  main()
