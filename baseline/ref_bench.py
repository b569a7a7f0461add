"""Times the UNTOUCHED reference (baseline/_ref, see ref_env.py) on the hot path's workload.

  cpu_train(...)   `losses.get_step_fn` of the reference on the host cores (the `--impl reference` arm and the
                   `cpu_baseline` of bench.py; BASELINE.md section 4.3: saturated figure at batch 64)
  gpu_train(...)   the same stock call on cuda:0 - eager fp32 (cuDNN / cuBLAS sm_100 kernels picked by PyTorch, TF32 convs
                   as torch defaults) and bf16 autocast + channels_last: the "existing Blackwell kernel" bar of
                   SURVEY 2.1 / BASELINE.md section 4.4
  gpu_sampler(...) `sampling.get_sampling_fn` (PC, Euler-Maruyama, no corrector) of the reference on cuda:0

Everything runs through the reference's own public API (`sde_lib.get_sde`, `models.utils.create_model`,
`losses.get_optimizer / optimization_manager / get_step_fn`, `models.ema.ExponentialMovingAverage`,
`sampling.get_sampling_fn`) on its own config file `configs/vp/CIFAR10/ddpmpp_nll_st.py`; none of this repo's kernels,
models or engine are on that path.  The one deviation: with more than one visible GPU the DataParallel wrapper of
`create_model` (models/utils.py:94) is pinned to device 0, so that an N=1 measurement uses one GPU.
"""
import contextlib
import os
import time

import numpy as np
import torch

from . import ref_env

CONFIG = 'vp/CIFAR10/ddpmpp_nll_st'


def available():
  return ref_env.locate() is not None


def _build(R, device, batch, seed=42):
  cfg = ref_env.ref_config(CONFIG, device)
  cfg.training.batch_size = batch
  torch.manual_seed(seed)
  np.random.seed(seed)
  sde = R.sde_lib.get_sde(cfg, None)
  if torch.device(device).type == 'cuda' and torch.cuda.device_count() > 1:
    net = R.mutils.get_model(cfg.model.name)(cfg, sde).to(cfg.device)
    model = torch.nn.DataParallel(net, device_ids=[torch.device(device).index or 0])
  else:
    model = R.mutils.create_model(cfg, sde)
  opt = R.losses.get_optimizer(cfg, model.parameters())
  ema = R.ema.ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate)
  state = dict(optimizer=opt, model=model, ema=ema, step=0)
  step_fn = R.losses.get_step_fn(cfg, sde, train=True, optimize_fn=R.losses.optimization_manager(cfg))
  return cfg, sde, state, step_fn


def cpu_train(batch=64, steps=2, warmup=1, threads=None):
  """images/s of the reference training step on the host cores."""
  R = ref_env.import_reference()
  cores = threads or os.cpu_count() or 1
  torch.set_num_threads(cores)
  cfg, sde, state, step_fn = _build(R, 'cpu', batch)
  gen = torch.Generator().manual_seed(1234)
  x = torch.rand(batch, 3, 32, 32, generator=gen) * 2 - 1
  for _ in range(warmup):
    step_fn(state, x)
  t0 = time.perf_counter()
  for _ in range(steps):
    losses_ = step_fn(state, x)
  dt = time.perf_counter() - t0
  assert torch.isfinite(losses_).all()
  return {'value': batch * steps / dt, 'unit': 'images/s', 'cores': cores, 'kind': 'reference', 'batch': batch,
          'steps': steps, 'warmup': warmup, 'ms_per_step': 1e3 * dt / steps,
          'sample': f'{steps} optimizer steps (after {warmup} warm-up) at batch {batch} of the same workload through the '
                    f'untouched reference losses.get_step_fn (baseline/_ref, fp32, torch {torch.__version__} CPU, {cores} threads)'}


def gpu_train(batch=512, steps=5, warmup=3, mode='fp32', device='cuda:0'):
  """images/s of the reference training step on one GPU.  mode: 'fp32' (stock eager) or 'bf16_autocast_channels_last'."""
  R = ref_env.import_reference()
  cfg, sde, state, step_fn = _build(R, device, batch)
  ctx = contextlib.nullcontext
  if mode != 'fp32':
    state['model'].to(memory_format=torch.channels_last)
    ctx = lambda: torch.autocast('cuda', dtype=torch.bfloat16)
  gen = torch.Generator(device=device).manual_seed(1234)
  x = torch.rand(batch, 3, 32, 32, generator=gen, device=device) * 2 - 1
  if mode != 'fp32':
    x = x.contiguous(memory_format=torch.channels_last)
  for _ in range(warmup):
    with ctx():
      step_fn(state, x)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(steps):
    with ctx():
      losses_ = step_fn(state, x)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  assert torch.isfinite(losses_).all()
  out = {'mode': mode, 'value': batch * steps / (ms * 1e-3), 'unit': 'images/s', 'batch': batch, 'steps': steps,
         'warmup': warmup, 'ms_per_step': ms / steps, 'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30,
         'tf32_conv': bool(torch.backends.cudnn.allow_tf32), 'tf32_matmul': bool(torch.backends.cuda.matmul.allow_tf32)}
  del state, step_fn
  torch.cuda.empty_cache()
  torch.cuda.reset_peak_memory_stats()
  return out


def gpu_train_best_batch(mode, steps, warmup, batches=(512, 256, 128)):
  """The reference step at the workload's batch (512), falling back to smaller batches if it does not fit."""
  err = None
  for b in batches:
    try:
      return gpu_train(b, steps, warmup, mode)
    except torch.OutOfMemoryError as ex:
      err = repr(ex)[:200]
      torch.cuda.empty_cache()
  return {'mode': mode, 'error': err}


def gpu_sampler(batch=1024, n_steps=10, device='cuda:0'):
  """PC-sampler reverse steps/s of the reference (Euler-Maruyama, no corrector, + final denoise) on one GPU."""
  R = ref_env.import_reference()
  cfg = ref_env.ref_config(CONFIG, device)
  cfg.sampling.method = 'pc'
  torch.manual_seed(42)
  sde = R.sde_lib.get_sde(cfg, None)
  if torch.cuda.device_count() > 1:
    model = torch.nn.DataParallel(R.mutils.get_model(cfg.model.name)(cfg, sde).to(cfg.device), device_ids=[0])
  else:
    model = R.mutils.create_model(cfg, sde)
  sde_s = R.sde_lib.VPSDE(truncation_time=cfg.training.truncation_time, beta_min=cfg.model.beta_min,
                          beta_max=cfg.model.beta_max, N=n_steps)      # the model keeps num_scales=1000
  fn = R.sampling.get_sampling_fn(cfg, sde_s, (batch, 3, 32, 32), lambda v: v, cfg.sampling.truncation_time)
  fn(model)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  x, nfe = fn(model)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  out = {'value': (n_steps + 1) / (ms * 1e-3), 'unit': 'steps/s', 'batch': batch, 'steps_timed': n_steps + 1,
         'mode': 'fp32 eager', 'note': 'reference sampling.get_sampling_fn (pc, euler_maruyama, none) incl. final denoise'}
  del model, fn
  torch.cuda.empty_cache()
  return out


if __name__ == '__main__':
  # `python -m baseline.ref_bench cpu <batch> <steps> <warmup>`: the CPU sample in a process of its own (bench.py runs it
  # with CUDA_VISIBLE_DEVICES='' - the reference's create_model wraps the model in DataParallel over every visible GPU)
  import json
  import sys
  kind, b, k, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
  assert kind == 'cpu'
  print('RESULT ' + json.dumps(cpu_train(batch=b, steps=k, warmup=w)), flush=True)
