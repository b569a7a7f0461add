"""Benchmark of the Soft-Truncation hot path on B200 (driver contract: one JSON line on stdout).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c2|c3|c4|c5|deepest]
                  [--mode train|sampler] [--batch B] [--sample-steps N]

Default workload (BASELINE.json configs[1], `--config c2`): DDPM++ (VP) CIFAR-10 32x32, batch 512 per GPU, bf16
compute / fp32 accumulate + master weights, soft-truncated importance-sampled DSM loss, dropout 0.1, grad-clip 1.0,
Adam, EMA: one "step" = losses.get_step_fn(...)(state, batch) on synthetic images.  `value` = images/s with the batch
resident in HBM; `e2e` = the same call fed from pinned host memory (H2D of the batch + D2H of the B losses inside the
timed region).  A secondary figure, PC-sampler reverse steps/s, is reported in `sampler`.  N > 1: one process per GPU
under torchrun, pure data parallel (weak scaling), one NCCL all-reduce of the flat gradient buffer per step.

`--config` selects the other BASELINE.json configurations at FULL size (c3 = configs[2] UNCSN++ RVE CelebA-64 B=128,
c4 = configs[3] DDPM++ ImageNet32 B=512, c5 = configs[4] NCSN++ VE CelebA-HQ-256 B=16, deepest = the README's FID
model); `--mode sampler` makes the PC sampler over the config's full schedule (`--sample-steps`, default the config's
N = 1000 / 2000) the headline value of the line.

Reference arms (the UNTOUCHED reference staged under baseline/_ref, see baseline/ref_env.py):
  * `--impl reference`: its own `losses.get_step_fn` on the host cores at the saturated batch 64;
    when baseline/_ref is absent the oracle port (oracle/ref_train.py) is timed instead and says so (`kind: "port"`).
  * `gpu_reference` (inside the default run, N = 1): the same stock call on cuda:0 at batch 512 - eager fp32 (cuDNN /
    cuBLAS sm_100 kernels) and bf16 autocast + channels_last - i.e. the "existing Blackwell kernel" bar.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

# DRAM bytes per GEMM launch (average over the GEMM launches of one B=512 bf16 c2 step), from the ncu launch list
# committed under profiles/ (dram__bytes_read.sum + dram__bytes_write.sum); see profiles/README.md for the file
GEMM_DRAM_BYTES_PER_LAUNCH = 52.16e9 / 420
TRAFFIC_SOURCE = 'profiles/r02_launches_train_step.md'

# name -> config path under the reference's configs/, per-GPU train batch, sampler batch, forward GF / image
# (2 x MACs of conv / linear / NIN / attention products; None = not surveyed)
WORKLOADS = {
    'c2': dict(path='vp/CIFAR10/ddpmpp_nll_st', batch=512, sample_batch=1024, fwd_gf=21.693, baseline='configs[1]',
               label='DDPM++ (VP) CIFAR-10 32x32', metric='DDPM++ CIFAR-10 train images/sec'),
    'c3': dict(path='ve/CELEBA/uncsnpp_st', batch=128, sample_batch=None, fwd_gf=83.957, baseline='configs[2]',
               label='UNCSN++ (RVE) CelebA 64x64', metric='UNCSN++ (RVE) CelebA-64 train images/sec'),
    'c4': dict(path='vp/IMAGENET32/ddpmpp_nll', batch=512, sample_batch=1024, fwd_gf=21.693, baseline='configs[3]',
               label='DDPM++ (VP) ImageNet32 32x32, NLL weighting', metric='DDPM++ ImageNet32 train images/sec'),
    'c5': dict(path='ve/celebahq/uncsnpp_st', batch=16, sample_batch=16, fwd_gf=531.775, baseline='configs[4]',
               label='NCSN++ (VE) CelebA-HQ 256x256', metric='NCSN++ CelebA-HQ-256 train images/sec'),
    'deepest': dict(path='vp/CIFAR10/ddpmpp_fid_st_deepest', batch=128, sample_batch=256, fwd_gf=None, baseline='README FID model',
                    label="DDPM++ 'deepest' (nf=512) CIFAR-10 32x32, step_fn_mixed", metric="DDPM++ 'deepest' CIFAR-10 train images/sec"),
}
METRIC = WORKLOADS['c2']['metric']


def peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    d = json.load(open(p))
    return d, 'measured'
  return {'hbm_gbs': 6650., 'bf16_tflops': 1590., 'bf16_tflops_sustained': 1400.}, 'fallback'


class ClockSampler:
  """Samples SM clocks / throttle reasons while the timed region runs: in-process NVML (pynvml, initialised BEFORE the
  timed region, one query every 100 ms from a thread - no process start-up or NVML initialisation inside the window),
  falling back to one looping `nvidia-smi -lms 200` process when pynvml is not importable."""
  Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
       'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.rows, self.proc, self.index = [], None, index
    self.nvml, self.handle, self.stop = None, None, threading.Event()
    try:
      import pynvml
      pynvml.nvmlInit()
      visible = os.environ.get('CUDA_VISIBLE_DEVICES')
      phys = index
      if visible:                                    # NVML enumerates physical devices
        ids = [v.strip() for v in visible.split(',') if v.strip()]
        if index < len(ids) and ids[index].isdigit():
          phys = int(ids[index])
      self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
      self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
      self.nvml = pynvml
    except Exception:
      self.nvml = None

  def __enter__(self):
    if self.nvml is not None:
      self.thread = threading.Thread(target=self._poll, daemon=True)
      self.thread.start()
      return self
    try:
      self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200',
                                    '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None
    return self

  def _poll(self):
    n = self.nvml
    flags = (('hw_slowdown', n.nvmlClocksThrottleReasonHwSlowdown), ('hw_thermal_slowdown', n.nvmlClocksThrottleReasonHwThermalSlowdown),
             ('sw_thermal_slowdown', n.nvmlClocksThrottleReasonSwThermalSlowdown), ('sw_power_cap', n.nvmlClocksThrottleReasonSwPowerCap))
    while True:
      try:
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        try:
          power = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.
        except Exception:
          power = 0.
        self.rows.append([str(sm), str(self.max_sm), str(power)] + ['Active' if mask & bit else 'Not Active' for _, bit in flags])
      except Exception:
        pass
      if self.stop.wait(0.1):
        return

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(',')])

  def __exit__(self, *a):
    self.stop.set()
    if self.nvml is not None:
      self.thread.join(timeout=2)
    if self.proc is not None:
      self.proc.terminate()
      try:
        self.proc.wait(timeout=2)
      except Exception:
        self.proc.kill()

  def summary(self):
    sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
    mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
    reasons = []
    for name, col in (('hw_slowdown', 3), ('hw_thermal_slowdown', 4), ('sw_thermal_slowdown', 5), ('sw_power_cap', 6)):
      if any(len(r) >= 7 and r[col].lower().startswith('active') for r in self.rows):
        reasons.append(name)
    return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
            'samples': len(sm), 'source': 'nvml' if self.nvml is not None else 'nvidia-smi'}


# ============================================================================================ reference arm
def cpu_reference_sample(batch, steps, warmup):
  """The reference training step on the host cores: the untouched reference when it is staged (baseline/_ref), else
  the oracle port."""
  from baseline import ref_bench
  if ref_bench.available():
    # a process of its own with the GPUs hidden: the reference's create_model wraps the model in DataParallel over
    # every visible device (models/utils.py:94), which would move a CPU run onto cuda:0
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    r = subprocess.run([sys.executable, '-m', 'baseline.ref_bench', 'cpu', str(batch), str(steps), str(warmup)], cwd=ROOT,
                       env=env, capture_output=True, text=True, timeout=1500)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')]
    if r.returncode != 0 or not lines:
      raise RuntimeError('reference CPU run failed: ' + (r.stderr or r.stdout)[-1500:])
    return json.loads(lines[-1][7:])
  from oracle import ref_model, ref_train
  from soft_truncation_b200 import configs
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  cfg = configs.cifar10_ddpmpp_nll_st()
  sde = ref_train.make_sde(cfg)
  state = ref_train.TrainState(ref_model.make_state_dict(cfg, seed=0))
  gen = torch.Generator().manual_seed(1234)
  x = torch.rand(batch, 3, 32, 32, generator=gen) * 2 - 1

  def one():
    u, z = torch.rand(batch, generator=gen), torch.randn(batch, 3, 32, 32, generator=gen)
    ref_train.train_step(state, cfg, sde, x, u, z, float(torch.rand(1, generator=gen)))

  for _ in range(warmup):
    one()
  t0 = time.perf_counter()
  for _ in range(steps):
    one()
  dt = time.perf_counter() - t0
  return {'value': batch * steps / dt, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'batch': batch, 'steps': steps,
          'warmup': warmup, 'ms_per_step': 1e3 * dt / steps,
          'sample': f'{steps} optimizer steps at batch {batch} (oracle/ref_train.train_step, fp32, dropout off: baseline/_ref not staged)'}


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  # bounded sample: the saturated CPU batch and at most 4 timed steps (~10 s each on 16 cores)
  B = int(os.environ.get('ST_BENCH_REF_BATCH', '64'))      # (the CPU test suite shrinks it)
  steps, warm = max(1, min(args.steps, 4)), max(1, min(args.warmup, 1))
  r = cpu_reference_sample(B, steps, warm)
  v = r['value']
  line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': steps,
          'warmup': warm, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
          'config': {'workload': f'DDPM++ (VP) CIFAR-10 32x32 soft-truncated DSM train step (BASELINE configs[1]), CPU, bounded sample batch {B}'},
          'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample']},
          'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
  print(json.dumps(line), flush=True)


# ============================================================================================ our arm
def _sampler_sde(sde_lib, cfg, N):
  kind = cfg.training.sde.lower()
  if kind == 'vpsde':
    return sde_lib.VPSDE(truncation_time=cfg.training.truncation_time, beta_min=cfg.model.beta_min,
                         beta_max=cfg.model.beta_max, N=N)
  if kind == 'vesde':
    return sde_lib.VESDE(sigma_min=cfg.model.sigma_min, sigma_max=cfg.model.sigma_max, N=N)
  return None      # the reference's PC sampler does not run for the reciprocal VE SDE


def run_b200(args):
  import torch.distributed as dist
  from soft_truncation_b200 import _lib, configs, losses, ops, sampling, sde_lib
  from soft_truncation_b200.models import utils as mutils
  from soft_truncation_b200.models.ema import ExponentialMovingAverage

  wl = WORKLOADS[args.config]
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    import datetime
    dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
  cfg = configs.get_config(wl['path'])
  cfg.device = dev
  cfg.model.compute_dtype = args.dtype
  B = args.batch or wl['batch']
  R = cfg.data.image_size
  if args.micro:
    cfg.optim.l2_blocks = args.micro
  cfg.training.batch_size = B * world
  # every rank builds the same initial weights (and losses.sync_replicas broadcasts rank 0's anyway), then ranks seed
  # torch apart so that their time / noise / dropout draws differ; NumPy (t_min) stays identical on all ranks
  torch.manual_seed(42)
  np.random.seed(42)
  sde = sde_lib.get_sde(cfg)
  model = mutils.create_model(cfg, sde)
  torch.manual_seed(42 + rank)
  net = mutils.unwrap(model)
  state = dict(model=model, optimizer=losses.get_optimizer(cfg, model.parameters()),
               ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)
  step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
  gen = torch.Generator(device=dev).manual_seed(1234 + rank)
  batch_dev = torch.rand(B, 3, R, R, generator=gen, device=dev)
  if cfg.data.centered:
    batch_dev = batch_dev * 2 - 1
  batch_host = batch_dev.cpu().pin_memory()
  pk, pk_kind = peaks()
  peak_tf = pk.get('bf16_tflops_sustained', pk.get('bf16_tflops'))

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(fn, steps):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      fn()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()

  train, roof, e2e_u8 = None, None, None
  if args.mode == 'train':
    _lib.lib.st_gemm_simt_fallbacks(1)
    for _ in range(max(args.warmup, 3)):
      step_fn(state, batch_dev)
    l0 = _lib.launches
    with ClockSampler(local) as clk:
      ms = timed(lambda: step_fn(state, batch_dev), args.steps)
    launches_host = _lib.launches - l0
    value = B * world * args.steps / (ms * 1e-3)

    def e2e_step():
      step_fn(state, batch_host.to(dev, non_blocking=True))
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e = B * world * args.steps / (ms_e2e * 1e-3)
    train = dict(ms=ms, value=value, ms_e2e=ms_e2e, e2e=e2e, launches_host=launches_host)

    # the same with the batch travelling as uint8 and prepared on the GPU (datasets.prepare_batch: /255, random flip,
    # uniform dequantisation, scaler in one kernel) - SURVEY 8(f)4; informative, the headline e2e stays the fp32 form
    try:
      if world > 1:
        raise RuntimeError('informative leg, measured at N=1 only')
      from soft_truncation_b200 import datasets
      u8_host = torch.randint(0, 256, (B, R, R, 3), dtype=torch.uint8).pin_memory()
      dq = cfg.data.dequantization
      cfg.data.dequantization = 'uniform'

      def e2e_u8_step():
        step_fn(state, datasets.prepare_batch(cfg, u8_host, train=True))
      try:
        e2e_u8_step()
        ms_u8 = timed(e2e_u8_step, args.steps)
      finally:
        cfg.data.dequantization = dq
      e2e_u8 = {'value': B * world * args.steps / (ms_u8 * 1e-3), 'unit': 'images/s',
                'h2d_bytes_per_step': u8_host.numel() * world, 'ms_per_step': ms_u8 / args.steps,
                'note': 'uint8 pinned host batch -> datasets.prepare_batch (flip, uniform dequantisation, scaler on the GPU) -> step_fn'}
    except Exception as ex:       # informative leg: the headline line must still print
      e2e_u8 = {'skipped': repr(ex)[:300]}

    # ---- dominant kernel: every st_gemm launch of one step bracketed by CUDA events (outside the timed region, on an
    # EAGER step: the captured-graph step the timed region replays launches exactly the same kernels)
    recs = []
    counts = {'calls': 0}
    orig, orig_check = ops._gemm, _lib.check

    def spy(**kw):
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      orig(**kw)
      b.record()
      recs.append((a, b, 2.0 * kw['M'] * kw['N'] * kw['K'] * kw.get('batch', 1)))
    ops._gemm = spy
    graph_was = losses._STEP_GRAPH
    losses._STEP_GRAPH = False
    l1 = _lib.launches
    try:
      # head start for the host: ~45 ms of device spin first, so that the (slower, eager) launch stream stays ahead of the
      # GPU and every event pair brackets a kernel that starts as soon as its predecessor ends - as in the captured graph
      torch.cuda._sleep(int(0.045 * 1.7e9))
      step_fn(state, batch_dev)
      torch.cuda.synchronize()
    finally:
      ops._gemm = orig
      losses._STEP_GRAPH = graph_was
    kernel_calls = _lib.launches - l1          # C-ABI calls of one step (each enqueues >= 1 of our kernels)
    fallbacks = int(_lib.lib.st_gemm_simt_fallbacks(0))
    if args.dtype == 'bf16' and fallbacks:
      raise RuntimeError(f'{fallbacks} bf16 GEMMs fell back to the SIMT kernel: {_lib.lib.st_gemm_simt_fallback_reason().decode()}')
    if rank == 0:
      t_ms = sum(a.elapsed_time(b) for a, b, _ in recs)
      flops = sum(f for _, _, f in recs)
      achieved = flops / (t_ms * 1e-3) / 1e12
      tc = ops.tc_available() and args.dtype == 'bf16'
      roof = {'bound': 'tensor', 'kernel': 'gemm_tc2_kernel (persistent tcgen05 implicit-GEMM conv / GEMM, all st_gemm launches of one step)' if tc else 'gemm_simt_kernel',
              'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf,
              'traffic': GEMM_DRAM_BYTES_PER_LAUNCH if (args.config == 'c2' and B == 512 and args.dtype == 'bf16') else None,
              'traffic_source': 'ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the GEMM launches of one B=512 step '
                                f'/ their count ({TRAFFIC_SOURCE})',
              'peak_source': pk_kind + ' (sustained cuBLAS bf16)', 'launches': len(recs), 'gemm_ms_per_step': t_ms,
              'gemm_share_of_step': t_ms / (ms / args.steps), 'simt_fallbacks': fallbacks,
              'whole_step_frac': (value / world * 3 * wl['fwd_gf'] / 1e3 / peak_tf) if wl['fwd_gf'] else None}
    if world > 1:
      dist.barrier()
    train['kernel_calls'] = kernel_calls

  # ---- PC sampler reverse steps/s: the config's own predictor / corrector, CUDA-graph replay
  samp = None
  clk_s = None
  try:
    SB = args.sample_batch or wl['sample_batch']
    if SB is None:
      raise RuntimeError('the reference has no working PC sampler for this SDE (SURVEY F7)')
    full = cfg.model.num_scales
    N = args.sample_steps or (full if args.mode == 'sampler' else 50)
    cfg.sampling.method = 'pc'
    sde_s = _sampler_sde(sde_lib, cfg, N)
    model.eval()
    fn = sampling.get_sampling_fn(cfg, sde_s, (SB, 3, R, R), lambda v: v, cfg.sampling.truncation_time)
    evals_per_step = 2 if cfg.sampling.corrector.lower() != 'none' else 1
    if args.mode == 'sampler':
      sde_w = _sampler_sde(sde_lib, cfg, 6)          # short warm-up call (captures the graph, fills the allocator)
      sampling.get_sampling_fn(cfg, sde_w, (SB, 3, R, R), lambda v: v, cfg.sampling.truncation_time)(model)
      fn(model)                                     # first full call captures this schedule's graph: untimed
      with ClockSampler(local) as clk_s:
        # ONE timed call: CUDA events around sampling_fn(model) for `value`; the host clock around the same call plus the
        # device->host copy of the samples for e2e (the prior draw on the CPU and its H2D copy are inside the call)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        lc0 = _lib.launches
        xs, _ = fn(model)
        e1.record()
        g = (getattr(fn, 'graph_cache', None) or {}).get('g') or {}
        # our kernels launched in the timed call: the replays of the captured reverse step + what ran eagerly (denoise step)
        samp_launches = (N * g['calls'] if g.get('graph') is not None else 0) + (_lib.launches - lc0)
        xs = xs.cpu()
        wall = time.perf_counter() - t0
        barrier()
        tms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev)
        if world > 1:
          dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_s, ms_s_e2e = tms[0].item(), tms[1].item()
        assert torch.isfinite(xs).all(), 'non-finite samples'
    else:
      fn(model)
      ms_s = timed(lambda: fn(model), 1)
      ms_s_e2e = None
    sps = (N + 1) / (ms_s * 1e-3)
    evals = N * evals_per_step + 1
    samp = {'metric': 'PC-sampler reverse steps/sec', 'value': sps, 'unit': 'steps/s', 'batch_per_gpu': SB,
            'schedule_N': N, 'steps_timed': N + 1, 'network_evals': evals, 'ms_per_step': ms_s / (N + 1),
            'sample_steps_per_sec': sps * SB * world,
            'sampler': f'{cfg.sampling.predictor}+{cfg.sampling.corrector}',
            'frac_of_tensor_roofline': (evals / (ms_s * 1e-3) * SB * wl['fwd_gf'] / 1e3 / peak_tf) if wl['fwd_gf'] else None,
            'note': f'{N} reverse steps of an N={N} schedule + final denoise through sampling.get_sampling_fn; one reverse '
                    'step is captured in a CUDA graph once (untimed first call) and replayed'}
    if ms_s_e2e is not None:
      samp['e2e_steps_per_sec'] = (N + 1) / (ms_s_e2e * 1e-3)
  except Exception as ex:   # the headline metric must still print
    samp = {'error': repr(ex)[:300]}
    if args.mode == 'sampler':
      raise

  # ---- free our state, then the reference legs (N = 1 only)
  cpu, gpu_ref = None, None
  if rank == 0 and world == 1 and args.mode == 'train' and args.config == 'c2':
    del state, step_fn, model, net
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if not args.no_gpu_reference:
      from baseline import ref_bench
      if ref_bench.available():
        gpu_ref = {'what': 'the UNTOUCHED reference (baseline/_ref) driving its own losses.get_step_fn on cuda:0: '
                           'cuDNN / cuBLAS / ATen sm_100 kernels under eager PyTorch ' + torch.__version__,
                   'our_value': train['value'], 'our_e2e': train['e2e']}
        for mode in ('fp32', 'bf16_autocast_channels_last'):
          try:
            r = ref_bench.gpu_train_best_batch(mode, steps=min(args.steps, 5), warmup=3)
          except Exception as ex:
            r = {'mode': mode, 'error': repr(ex)[:300]}
          if 'value' in r:
            r['ours_over_reference'] = train['value'] / r['value']
          gpu_ref[mode] = r
        try:
          r = ref_bench.gpu_sampler(batch=1024, n_steps=10)
          if samp and 'value' in samp:
            r['ours_over_reference'] = samp['value'] / r['value']
          gpu_ref['sampler_fp32'] = r
        except Exception as ex:
          gpu_ref['sampler_fp32'] = {'error': repr(ex)[:300]}
      else:
        gpu_ref = {'unavailable': 'baseline/_ref is not staged on this box'}
    if not args.no_cpu_baseline:
      try:
        r = cpu_reference_sample(64, 2, 1)
        cpu = {'value': r['value'], 'unit': 'images/s', 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample']}
      except Exception as ex:       # the headline line must still print
        cpu = {'error': repr(ex)[:300]}

  if rank == 0:
    common = {'n_gpus': world, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype,
              'data': 'synthetic', 'tcgen05': bool(ops.tc_available())}
    if args.mode == 'train':
      ms, value = train['ms'], train['value']
      line = {'metric': wl['metric'], 'value': value, 'unit': 'images/s', 'steps': args.steps,
              'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, **common,
              'config': {'workload': f"{wl['label']}, batch {B}/GPU, {args.dtype} compute + fp32 master, "
                                     f"{'soft-truncated ' if cfg.training.st else ''}DSM loss, dropout {cfg.model.dropout}, "
                                     f"clip+Adam+EMA (BASELINE {wl['baseline']})",
                         'global_batch': B * world, 'parallelism': f'dp{world}',
                         'step_graph': bool(losses._STEP_GRAPH),
                         'l2': 'per-step working set (>20 GB of activations) is far larger than the 126 MB L2; no flush needed'},
              'e2e': {'value': train['e2e'], 'unit': 'images/s', 'h2d_bytes_per_step': batch_host.numel() * 4 * world,
                      'd2h_bytes_per_step': B * 4 * world, 'ms_per_step': train['ms_e2e'] / args.steps},
              'e2e_u8': e2e_u8,
              # our kernels launched inside the timed region: C-ABI calls of one step (each enqueues >= 1 kernel; the
              # captured graph replays the same launches) x steps
              'gpu_launches': train['kernel_calls'] * args.steps, 'host_calls_in_timed_region': train['launches_host'],
              'clocks': clk.summary(), 'roofline': roof, 'cpu_baseline': cpu, 'gpu_reference': gpu_ref, 'sampler': samp}
    else:
      N = samp['schedule_N']
      line = {'metric': f"{wl['label']} PC-sampler sample-steps/sec", 'value': samp['sample_steps_per_sec'],
              'unit': 'sample-steps/s', 'steps': samp['steps_timed'], 'warmup': 1, 'ms_per_step': samp['ms_per_step'],
              **common,
              'config': {'workload': f"{wl['label']}, {samp['sampler']} PC sampler, batch {samp['batch_per_gpu']}/GPU, "
                                     f"full N={N} schedule + denoise, {args.dtype} compute (BASELINE {wl['baseline']})",
                         'parallelism': f'{world} independent samplers', 'l2': 'activations of one network evaluation exceed L2'},
              'steps_per_sec_per_gpu': samp['value'],
              'e2e': {'value': samp.get('e2e_steps_per_sec', 0.) * samp['batch_per_gpu'] * world, 'unit': 'sample-steps/s',
                      'h2d_bytes_per_step': samp['batch_per_gpu'] * 3 * R * R * 4 * world / samp['steps_timed'],
                      'd2h_bytes_per_step': samp['batch_per_gpu'] * 3 * R * R * 4 * world / samp['steps_timed']},
              'gpu_launches': samp_launches, 'clocks': clk_s.summary() if clk_s is not None else None,
              'roofline': {'bound': 'tensor', 'achieved': (samp['frac_of_tensor_roofline'] or 0.) * peak_tf, 'peak': peak_tf,
                           'unit': 'TFLOP/s', 'frac': samp['frac_of_tensor_roofline'], 'traffic': None,
                           'kernel': 'whole reverse step (algorithmic GF of the network evaluations / time)'},
              'cpu_baseline': None, 'sampler': samp}
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200')
  ap.add_argument('--config', default='c2', choices=sorted(WORKLOADS))
  ap.add_argument('--mode', default='train', choices=['train', 'sampler'])
  ap.add_argument('--batch', type=int, default=0, help='per-GPU train batch (0 = the workload\'s)')
  ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
  ap.add_argument('--sample-batch', type=int, default=0)
  ap.add_argument('--sample-steps', type=int, default=0, help='PC schedule length (0 = 50 in train mode, the config\'s N in sampler mode)')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-gpu-reference', action='store_true')
  ap.add_argument('--micro', type=int, default=0,
                  help='L2 blocking: run the step as this many image blocks (0 = the library default for the batch)')
  args = ap.parse_args()
  # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout when
  # the first communicator is created), so file descriptor 1 points at stderr while the benchmark runs and is restored
  # for the result line only.
  sys.stdout.flush()
  real_stdout = os.dup(1)
  os.dup2(2, 1)
  out = io.StringIO()
  try:
    with contextlib.redirect_stdout(out):
      if args.impl == 'reference':
        run_reference(args)
      else:
        run_b200(args)
  finally:
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
  lines = [ln for ln in out.getvalue().splitlines() if ln.strip()]
  for ln in lines:
    if not ln.startswith('{'):
      print(ln, file=sys.stderr)
  for ln in lines:
    if ln.startswith('{'):
      print(ln, flush=True)


if __name__ == '__main__':
  main()
