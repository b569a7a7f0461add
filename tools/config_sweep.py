"""Full-size smoke + throughput of every BASELINE.json config (and the 'deepest' DDPM++) on one GPU.

bench.py measures configs[1]; the others are parity cases at reduced width in tests/.  This tool builds each of them at
FULL size in bf16, runs a few optimizer steps through losses.get_step_fn and a few reverse steps through
sampling.get_sampling_fn, checks finiteness and prints one JSON line per config (images/s, sampler steps/s, fraction
of the tensor roofline from SURVEY 8(d)'s GF/image).  One subprocess per config, each under its own timeout.

    python tools/config_sweep.py [c3 c4 c5 deepest]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name -> (config path, per-GPU train batch, sampler batch, forward GF/image (None = not surveyed))
CASES = {
    'c3': ('ve/CELEBA/uncsnpp_st', 128, 64, 83.957),
    'c4': ('vp/IMAGENET32/ddpmpp_nll', 512, 1024, 21.693),
    'c5': ('ve/celebahq/uncsnpp_st', 16, 16, 531.775),
    'deepest': ('vp/CIFAR10/ddpmpp_fid_st_deepest', 128, 256, None),
}


def run_case(name):
  import numpy as np
  import torch
  from soft_truncation_b200 import configs, losses, sampling, sde_lib
  from soft_truncation_b200.models import utils as mutils
  from soft_truncation_b200.models.ema import ExponentialMovingAverage
  path, B, SB, gf = CASES[name]
  dev = torch.device('cuda:0')
  cfg = configs.get_config(path)
  cfg.device = dev
  cfg.model.compute_dtype = 'bf16'
  R = cfg.data.image_size
  torch.manual_seed(42)
  np.random.seed(42)
  sde = sde_lib.get_sde(cfg)
  model = mutils.create_model(cfg, sde)
  n_params = sum(p.numel() for p in model.parameters())
  state = dict(model=model, optimizer=losses.get_optimizer(cfg, model.parameters()),
               ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)
  step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
  batch = torch.rand(B, 3, R, R, device=dev)
  if cfg.data.centered:
    batch = batch * 2 - 1
  out = {'config': path, 'image': R, 'params': n_params, 'train_batch': B}

  def timed(fn, n):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
      r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, r

  for _ in range(2):
    ls = step_fn(state, batch)
  ms, ls = timed(lambda: step_fn(state, batch), 3)
  assert torch.isfinite(ls).all(), 'non-finite training loss'
  out.update(train_ms_per_step=ms, train_images_per_s=B / ms * 1e3, loss_mean=float(ls.mean()))
  if gf:
    out['train_frac_of_tensor_roofline'] = B / ms * 1e3 * 3 * gf / 1e3 / 1387.1
  print(json.dumps(dict(out, partial='train only')), flush=True)
  torch.cuda.empty_cache()
  # ---- PC sampler: the config's own predictor / corrector on a short schedule of the same step size class
  N = 6
  cfg.sampling.method = 'pc'
  kind = cfg.training.sde.lower()
  if kind == 'vpsde':
    sde_s = sde_lib.VPSDE(truncation_time=cfg.training.truncation_time, beta_min=cfg.model.beta_min, beta_max=cfg.model.beta_max, N=N)
  elif kind == 'vesde':
    sde_s = sde_lib.VESDE(sigma_min=cfg.model.sigma_min, sigma_max=cfg.model.sigma_max, N=N)
  else:
    sde_s = None               # the reference's PC sampler does not run for the reciprocal VE SDE
  if sde_s is not None:
    model.eval()
    fn = sampling.get_sampling_fn(cfg, sde_s, (SB, 3, R, R), lambda v: v, cfg.sampling.truncation_time)
    fn(model)
    net = mutils.unwrap(model)
    ms_s, (x, nfe) = timed(lambda: fn(model), 1)
    fwd = (N * (2 if cfg.sampling.corrector.lower() != 'none' else 1)) + 1      # network evaluations incl. the denoise step
    assert torch.isfinite(x).all(), 'non-finite sample'
    out.update(sampler=f'{cfg.sampling.predictor}+{cfg.sampling.corrector}', sample_batch=SB, sampler_nfe=nfe, sampler_network_evals=fwd,
               sampler_steps_per_s=(N + 1) / ms_s * 1e3, sampler_sample_steps_per_s=(N + 1) / ms_s * 1e3 * SB)
    if gf:
      out['sampler_frac_of_tensor_roofline'] = fwd / ms_s * 1e3 * SB * gf / 1e3 / 1387.1
  out['peak_mem_gb'] = torch.cuda.max_memory_allocated() / 2 ** 30
  print(json.dumps(out), flush=True)


def main():
  if len(sys.argv) > 2 and sys.argv[1] == '--case':
    run_case(sys.argv[2])
    return
  for name in (sys.argv[1:] or list(CASES)):
    t0 = time.time()
    try:
      r = subprocess.run([sys.executable, os.path.abspath(__file__), '--case', name], capture_output=True, text=True,
                         timeout=float(os.environ.get('SWEEP_TIMEOUT', '240')))
    except subprocess.TimeoutExpired as ex:
      print(name, json.dumps({'error': 'timeout', 'tail': str(ex.stderr or '')[-800:]}), flush=True)
      continue
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    if r.returncode == 0 and lines:
      print(name, lines[-1], flush=True)
    else:
      print(name, json.dumps({'error': (r.stderr or r.stdout)[-1500:], 'rc': r.returncode, 's': time.time() - t0,
                              'partial': lines[-1] if lines else None}), flush=True)


if __name__ == '__main__':
  main()
