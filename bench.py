"""Benchmark of the Soft-Truncation hot path on B200 (driver contract: one JSON line on stdout).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B]

Workload (BASELINE.json configs[1]): DDPM++ (VP) CIFAR-10 32x32, batch 512 per GPU, bf16 compute / fp32
accumulate+master weights, soft-truncated importance-sampled DSM loss, dropout 0.1, grad-clip 1.0, Adam, EMA:
one "step" = losses.get_step_fn(...)(state, batch) on synthetic images.  `value` = images/s with the batch
resident in HBM; `e2e` = the same call fed from pinned host memory (H2D of the batch + D2H of the B losses inside
the timed region).  A secondary figure, PC-sampler reverse steps/s (Euler-Maruyama, batch 1024), is reported in
`sampler`.  N > 1: one process per GPU under torchrun, pure data parallel (weak scaling), one NCCL all-reduce of
the flat gradient buffer per step.

`--impl reference` times the reference's CPU path: the reference is pure Python/PyTorch and cannot travel to the
GPU box, so this arm runs the oracle port (oracle/ref_train.py, checked against reference-generated fixtures) on
all host cores, on a bounded sample (batch 8) of the same workload.
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

TRAIN_GF_PER_IMG = 65.08      # 3 x 21.693 GF forward (conv/linear/NIN/attention MACs x 2)
FWD_GF_PER_IMG = 21.693
# DRAM bytes per GEMM launch (average over the 430 launches of one B=512 bf16 step): 54.15 GB / 430, from the ncu launch
# list committed as profiles/r01_launches_train_step.md (dram__bytes_read.sum + dram__bytes_write.sum)
GEMM_DRAM_BYTES_PER_LAUNCH = 54.166e9 / 430
METRIC = 'DDPM++ CIFAR-10 train images/sec'


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
def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  from oracle import ref_model, ref_train
  from soft_truncation_b200 import configs
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  cfg = configs.cifar10_ddpmpp_nll_st()
  B = 8
  sde = ref_train.make_sde(cfg)
  state = ref_train.TrainState(ref_model.make_state_dict(cfg, seed=0))
  gen = torch.Generator().manual_seed(1234)
  batch = torch.rand(B, 3, 32, 32, generator=gen) * 2 - 1
  steps, warm = min(args.steps, 4), min(args.warmup, 1)

  def one(i):
    u, z = torch.rand(B, generator=gen), torch.randn(B, 3, 32, 32, generator=gen)
    ref_train.train_step(state, cfg, sde, batch, u, z, float(torch.rand(1, generator=gen)))

  for i in range(warm):
    one(i)
  t0 = time.perf_counter()
  for i in range(steps):
    one(i)
  dt = time.perf_counter() - t0
  v = B * steps / dt
  line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': steps,
          'warmup': warm, 'ms_per_step': 1e3 * dt / steps, 'higher_is_better': True, 'scaling': 'weak',
          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
          'config': {'workload': 'DDPM++ (VP) CIFAR-10 32x32 soft-truncated DSM train step, CPU, bounded sample batch 8'},
          'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                           'sample': f'{steps} optimizer steps at batch {B} (oracle/ref_train.train_step, fp32, dropout injected off)'},
          'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
  print(json.dumps(line), flush=True)


# ============================================================================================ our arm
def run_b200(args):
  import torch.distributed as dist
  from soft_truncation_b200 import _lib, configs, losses, ops, sampling, sde_lib
  from soft_truncation_b200.models import utils as mutils
  from soft_truncation_b200.models.ema import ExponentialMovingAverage

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    import datetime
    dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
  cfg = configs.cifar10_ddpmpp_nll_st()
  cfg.device = dev
  cfg.model.compute_dtype = args.dtype
  B = args.batch
  if args.micro:
    cfg.optim.l2_blocks = args.micro
  cfg.training.batch_size = B * world
  torch.manual_seed(42)
  np.random.seed(42)
  sde = sde_lib.get_sde(cfg)
  model = mutils.create_model(cfg, sde)
  net = mutils.unwrap(model)
  state = dict(model=model, optimizer=losses.get_optimizer(cfg, model.parameters()),
               ema=ExponentialMovingAverage(model.parameters(), decay=cfg.model.ema_rate), step=0)
  step_fn = losses.get_step_fn(cfg, sde, train=True, optimize_fn=losses.optimization_manager(cfg))
  gen = torch.Generator(device=dev).manual_seed(1234 + rank)
  batch_dev = torch.rand(B, 3, 32, 32, generator=gen, device=dev) * 2 - 1
  batch_host = batch_dev.cpu().pin_memory()

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

  for _ in range(max(args.warmup, 3)):
    step_fn(state, batch_dev)
  l0 = _lib.launches
  with ClockSampler(local) as clk:
    ms = timed(lambda: step_fn(state, batch_dev), args.steps)
  launches = _lib.launches - l0
  value = B * world * args.steps / (ms * 1e-3)

  def e2e_step():
    step_fn(state, batch_host.to(dev, non_blocking=True))
  e2e_step()
  ms_e2e = timed(e2e_step, args.steps)
  e2e = B * world * args.steps / (ms_e2e * 1e-3)

  # the same with the batch travelling as uint8 and prepared on the GPU (datasets.prepare_batch: /255, random flip,
  # uniform dequantisation, scaler in one kernel) - SURVEY 8(f)4; informative, the headline e2e stays the fp32 form
  e2e_u8 = None
  try:
    if world > 1:
      raise RuntimeError('informative leg, measured at N=1 only')
    from soft_truncation_b200 import datasets
    u8_host = torch.randint(0, 256, (B, 32, 32, 3), dtype=torch.uint8).pin_memory()
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

  # ---- dominant kernel: every st_gemm launch of one step bracketed by CUDA events (outside the timed region)
  pk, pk_kind = peaks()
  roof = None
  # every rank runs the instrumented step (it contains the gradient all-reduce); rank 0 reports
  recs = []
  orig = ops._gemm

  def spy(**kw):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    orig(**kw)
    b.record()
    recs.append((a, b, 2.0 * kw['M'] * kw['N'] * kw['K'] * kw.get('batch', 1)))
  ops._gemm = spy
  try:
    step_fn(state, batch_dev)
    torch.cuda.synchronize()
  finally:
    ops._gemm = orig
  if rank == 0:
    t_ms = sum(a.elapsed_time(b) for a, b, _ in recs)
    flops = sum(f for _, _, f in recs)
    achieved = flops / (t_ms * 1e-3) / 1e12
    peak = pk.get('bf16_tflops_sustained', pk.get('bf16_tflops'))
    roof = {'bound': 'tensor', 'kernel': 'gemm_tc2_kernel (persistent tcgen05 implicit-GEMM conv / GEMM, all st_gemm launches of one step)' if ops.tc_available() and args.dtype == 'bf16' else 'gemm_simt_kernel',
            'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
            'traffic': GEMM_DRAM_BYTES_PER_LAUNCH if B == 512 and args.dtype == 'bf16' else None,
            'traffic_source': 'ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the 430 GEMM launches of one B=512 step '
                              '/ 430 (profiles/r01_launches_train_step.md); algorithmic operand+output bytes of the same launches: see DESIGN.md',
            'peak_source': pk_kind + ' (sustained cuBLAS bf16)', 'launches': len(recs), 'gemm_ms_per_step': t_ms,
            'gemm_share_of_step': t_ms / (ms / args.steps),
            'whole_step_frac': value / world * TRAIN_GF_PER_IMG / 1e3 / peak}
  if world > 1:
    dist.barrier()

  # ---- secondary: PC sampler reverse steps/s (Euler-Maruyama, no corrector), CUDA-graph replay
  samp = None
  try:
    SB, N = args.sample_batch, args.sample_steps
    cfg.sampling.method = 'pc'
    sde_s = sde_lib.VPSDE(truncation_time=cfg.training.truncation_time, beta_min=cfg.model.beta_min,
                          beta_max=cfg.model.beta_max, N=N)
    fn = sampling.get_sampling_fn(cfg, sde_s, (SB, 3, 32, 32), lambda v: v, cfg.sampling.truncation_time)
    fn(model)
    ms_s = timed(lambda: fn(model), 1)
    sps = (N + 1) / (ms_s * 1e-3)
    samp = {'metric': 'PC-sampler reverse steps/sec', 'value': sps, 'unit': 'steps/s', 'batch_per_gpu': SB,
            'steps_timed': N + 1, 'sample_steps_per_sec': sps * SB * world,
            'frac_of_tensor_roofline': sps * SB * FWD_GF_PER_IMG / 1e3 / pk.get('bf16_tflops_sustained', 1400.),
            'note': f'{N} Euler-Maruyama steps of an N={N} VP schedule + final denoise through sampling.get_sampling_fn; '
                    'one reverse step is captured in a CUDA graph once (untimed first call) and replayed'}
  except Exception as ex:   # the headline metric must still print
    samp = {'error': repr(ex)[:300]}

  # ---- CPU baseline: the oracle port on this box's host cores, bounded sample
  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    from oracle import ref_model, ref_train
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ccfg = configs.cifar10_ddpmpp_nll_st()
    cst = ref_train.TrainState(ref_model.make_state_dict(ccfg, seed=0))
    csde = ref_train.make_sde(ccfg)
    cb = 8
    g2 = torch.Generator().manual_seed(1)
    xb = torch.rand(cb, 3, 32, 32, generator=g2) * 2 - 1
    t0 = time.perf_counter()
    n = 0
    while n < 3:
      ref_train.train_step(cst, ccfg, csde, xb, torch.rand(cb, generator=g2), torch.randn(cb, 3, 32, 32, generator=g2), 0.5)
      n += 1
    dt = time.perf_counter() - t0
    cpu = {'value': cb * n / dt, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
           'sample': f'{n} optimizer steps at batch {cb} of the same workload (oracle/ref_train.train_step, fp32)'}

  if rank == 0:
    line = {'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': f'DDPM++ (VP) CIFAR-10 32x32, batch {B}/GPU, {args.dtype} compute + fp32 master, '
                                   'soft-truncated IS-DSM loss, dropout 0.1, clip+Adam+EMA (BASELINE configs[1])',
                       'global_batch': B * world, 'parallelism': f'dp{world}',
                       'l2': 'per-step working set (>20 GB of activations) is far larger than the 126 MB L2; no flush needed'},
            'e2e': {'value': e2e, 'unit': 'images/s', 'h2d_bytes_per_step': batch_host.numel() * 4 * world,
                    'd2h_bytes_per_step': B * 4 * world, 'ms_per_step': ms_e2e / args.steps},
            'e2e_u8': e2e_u8, 'gpu_launches': launches, 'clocks': clk.summary(), 'roofline': roof, 'cpu_baseline': cpu, 'sampler': samp,
            'tcgen05': bool(ops.tc_available())}
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200')
  ap.add_argument('--batch', type=int, default=512)
  ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
  ap.add_argument('--sample-batch', type=int, default=1024)
  ap.add_argument('--sample-steps', type=int, default=50)
  ap.add_argument('--no-cpu-baseline', action='store_true')
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
