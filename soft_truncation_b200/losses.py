"""Training step of the soft-truncated denoising score matching objective
(reference losses.py:29-58, 61-168, 218-325) on the B200 kernels.

Same factory surface as the reference (`get_optimizer`, `optimization_manager`, `get_sde_loss_fn`,
`get_step_fn` -> `step_fn(state, batch) -> CPU tensor of per-sample losses`).  What changes underneath:

  * perturbation and the loss head are two fused kernels (st_dsm_perturb / st_dsm_loss); the loss is an
    autograd node that hands d(out) straight to the network's explicit backward;
  * global-norm clipping + Adam + EMA are ONE pass over the flat parameter buffer (st_sumsq +
    st_adam_ema) with the clip coefficient kept on the device (no host sync);
  * under torch.distributed (one process per GPU) the flat gradient buffer is all-reduced once per step
    over NCCL and `t_min` is broadcast from rank 0 (the reference's DataParallel replicates weights
    and reduces gradients through GPU 0 every step, models/utils.py:94).
"""
import ctypes
import inspect
import math
import os
import pickle
import warnings

import numpy as np
import torch
import torch.distributed as dist
import torch.optim as optim

from . import ops
from ._lib import check, lib
from .models import params as _params
from .models import utils as mutils
from .models.ema import ExponentialMovingAverage
from .sde_lib import VESDE, VPSDE


# ------------------------------------------------------------------------------------ optimizer
class FusedAdam(optim.Optimizer):
  """Adam over the flat fp32 parameter buffer of an NCSNpp; `state_dict()` keeps torch.optim.Adam's
  format (per-parameter 'step', 'exp_avg', 'exp_avg_sq') so checkpoints interchange with the reference
  (utils.py:29-36)."""

  def __init__(self, params, model, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0., amsgrad=False):
    if amsgrad:
      raise NotImplementedError('amsgrad is not built into the fused optimizer')
    params = list(params)
    super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))
    self.model = model
    self.m = torch.zeros_like(model._flat)
    self.v = torch.zeros_like(model._flat)
    self.t = 0
    self.gnorm_sq = torch.zeros(1, dtype=torch.float32, device=model._flat.device)
    self._bind_state()

  def _bind_state(self):
    step = torch.tensor(float(self.t))
    for e in self.model.store.entries:
      p = self.model._params[e.name]
      if e.trainable:
        self.state[p] = {'step': step, 'exp_avg': _params._logical_view(self.m, e),
                         'exp_avg_sq': _params._logical_view(self.v, e)}

  def state_dict(self):
    """torch.optim.Adam's format; the per-parameter 'step' entries are refreshed here (not 564 tensor allocations
    per optimizer step)."""
    step = torch.tensor(float(self.t))
    for st in self.state.values():
      st['step'] = step
    return super().state_dict()

  def hyper(self):
    """(lr, 1-b1^t, 1-b2^t) of the step that is about to be taken (self.t already advanced)."""
    g = self.param_groups[0]
    b1, b2 = g['betas']
    return float(g['lr']), 1. - b1 ** self.t, 1. - b2 ** self.t

  def zero_grad(self, set_to_none=False):
    self.model.zero_grad()

  def grad_norm_sq(self):
    """Sum of squared gradients, left on the device."""
    self.gnorm_sq.zero_()
    g = self.model._grad
    check(lib.st_sumsq(ops.ptr(g), g.numel(), ops.ptr(self.gnorm_sq), ops.stream()))
    return self.gnorm_sq

  @torch.no_grad()
  def step(self, closure=None, clip=-1., gnorm_sq=None, ema=None, ema_decay=0., dyn=None, advance=True):
    """`dyn`: device fp32 {lr, bc1, bc2, ema_decay} read by the kernel instead of the by-value scalars (a captured
    graph replays with the current step's values); `advance=False`: the caller keeps the step counter itself."""
    m = self.model
    if not m._flat.is_cuda:
      raise RuntimeError('FusedAdam needs the parameters on a CUDA device')
    g = self.param_groups[0]
    if advance:
      self.t += 1
    b1, b2 = g['betas']
    lr, bc1, bc2 = self.hyper()
    p16 = m._comp if m._comp is not m._flat else None
    check(lib.st_adam_ema(ops.ptr(m._flat), ops.ptr(m._grad), ops.ptr(self.m), ops.ptr(self.v),
                          ops.ptr(ema.shadow_flat) if ema is not None else None,
                          ops.ptr(ema.mask) if ema is not None else None, ops.ptr(p16), m._flat.numel(),
                          ops.ptr(gnorm_sq), float(clip), lr, float(b1), float(b2), float(g['eps']),
                          float(g['weight_decay']), bc1, bc2, float(ema_decay), ops.ptr(dyn), ops.stream()))

  def load_state_dict(self, state_dict):
    super().load_state_dict(state_dict)
    steps = [int(st['step']) for st in self.state.values() if 'step' in st]
    self.t = max(steps) if steps else 0
    for e in self.model.store.entries:
      p = self.model._params[e.name]
      if e.trainable and p in self.state:
        _params._logical_view(self.m, e).copy_(self.state[p]['exp_avg'])
        _params._logical_view(self.v, e).copy_(self.state[p]['exp_avg_sq'])
    self._bind_state()


def get_optimizer(config, params):
  """Adam with the reference's hyper-parameters (losses.py:29-41); fused when `params` are the
  parameters of an NCSNpp on a CUDA device."""
  params = list(params)
  owner = _params.flat_owner(params)
  if config.optim.optimizer == 'Adam':
    if owner is not None and owner._flat.is_cuda and not config.optim.amsgrad:
      return FusedAdam(params, owner, lr=config.optim.lr, betas=(config.optim.beta1, 0.999), eps=config.optim.eps,
                       weight_decay=config.optim.weight_decay)
    return optim.Adam(params, lr=config.optim.lr, betas=(config.optim.beta1, 0.999), eps=config.optim.eps,
                      weight_decay=config.optim.weight_decay, amsgrad=config.optim.amsgrad)
  elif config.optim.optimizer == 'AdamW':
    return optim.AdamW(params, lr=config.optim.lr, betas=(config.optim.beta1, 0.99), eps=config.optim.eps,
                       weight_decay=config.optim.weight_decay)
  raise NotImplementedError(f'Optimizer {config.optim.optimizer} not supported yet!')


def optimization_manager(config):
  """optimize_fn(optimizer, params, step, ...) with warm-up and global-norm clipping
  (reference losses.py:44-58).  `ema` (optional) folds the EMA update into the same kernel."""

  def optimize_fn(optimizer, params, step, lr=config.optim.lr, warmup=config.optim.warmup,
                  grad_clip=config.optim.grad_clip, ema=None):
    if warmup > 0:
      for g in optimizer.param_groups:
        g['lr'] = lr * np.minimum(step / warmup, 1.0)
    if isinstance(optimizer, FusedAdam):
      gn = optimizer.grad_norm_sq() if grad_clip >= 0 else None
      if ema is not None and ema.owner is optimizer.model:
        optimizer.step(clip=grad_clip, gnorm_sq=gn, ema=ema, ema_decay=ema.next_decay())
      else:
        optimizer.step(clip=grad_clip, gnorm_sq=gn)
        if ema is not None:
          ema.update(params)
      return
    if grad_clip >= 0:
      torch.nn.utils.clip_grad_norm_(params, max_norm=grad_clip)
    optimizer.step()
    if ema is not None:
      ema.update(params)

  # what the captured-graph step needs to reproduce this function without calling it (losses._StepGraph)
  optimize_fn.st_recipe = dict(lr=config.optim.lr, warmup=config.optim.warmup, grad_clip=config.optim.grad_clip)
  return optimize_fn


# ------------------------------------------------------------------------------------ loss
def _dsm_loss_raw(out, z, a, b, w, reduce_mean, gvec=None):
  """st_dsm_loss: losses[n] = w[n] * red_i (a[n]*out[n,i] + b[n]*z[n,i])^2; with `gvec` (d loss / d losses[n]) also
  d(out).  Returns (losses, dout | None)."""
  B, D = out.shape[0], out[0].numel()
  losses = torch.empty(B, dtype=torch.float32, device=out.device)
  dout = torch.empty_like(out) if gvec is not None else None
  check(lib.st_dsm_loss(ops.ptr(out), ops.ptr(z), ops.ptr(a), ops.ptr(b), ops.ptr(w), ops.ptr(losses), ops.ptr(dout),
                        ops.ptr(gvec), B, D, int(reduce_mean), ops.stream()))
  return losses, dout


class _DsmLoss(torch.autograd.Function):
  """Autograd node over st_dsm_loss; backward returns d(out)."""

  @staticmethod
  def forward(ctx, out, z, a, b, w, reduce_mean):
    out, z = out.contiguous(), z.contiguous()
    losses, _ = _dsm_loss_raw(out, z, a, b, w, reduce_mean)
    ctx.save_for_backward(out, z, a, b, w)
    ctx.reduce_mean = reduce_mean
    return losses

  @staticmethod
  def backward(ctx, dl):
    out, z, a, b, w = ctx.saved_tensors
    _, dout = _dsm_loss_raw(out, z, a, b, w, ctx.reduce_mean, gvec=dl.float().contiguous())
    return dout, None, None, None, None, None


def _vec(v, B, device):
  if torch.is_tensor(v):
    return v.to(device=device, dtype=torch.float32).expand(B).contiguous()
  return torch.full((B,), float(v), dtype=torch.float32, device=device)


def _dsm_inputs(config, sde, batch, t, Z, z):
  """Everything between the random draws and the network call, and the per-sample loss coefficients
  (reference losses.py:116-132, models/utils.py:128-190): returns (x_t, labels, a, b, w) with
  losses[n] = w[n] * reduce((a[n]*out + b[n]*z)^2) for the raw network output `out`."""
  tr = config.training
  B, dev = batch.shape[0], batch.device
  unit = torch.ones((B, 1, 1, 1), device=dev)
  mean_coeff, std = sde.marginal_prob(unit, t)
  mean_coeff = mean_coeff.reshape(B).float().contiguous()
  std = std.float().contiguous()
  xt = torch.empty_like(batch)
  check(lib.st_dsm_perturb(ops.ptr(batch), ops.ptr(z), ops.ptr(mean_coeff), ops.ptr(std), ops.ptr(xt), B,
                           batch[0].numel(), ops.stream()))
  # raw network output; the score is c[n]*out with c = -1/std (VP, ddpm_score) or 1 (models/utils.py:128-190)
  if isinstance(sde, VPSDE):
    if tr.continuous:
      if tr.unbounded_parametrization:
        c0 = tr.stabilizing_constant
        lo = sde.antiderivative(1e-5, stabilizing_constant=c0)
        labels = (sde.antiderivative(t, stabilizing_constant=c0) - lo) / \
                 (sde.antiderivative(sde.T, stabilizing_constant=c0) - lo) * 999.
      else:
        labels = t * 999
    else:
      raise NotImplementedError('discrete-time VP training is not built')
    c = -1. / std if tr.ddpm_score else torch.ones_like(std)
  else:
    labels = std if tr.continuous else torch.round((sde.T - t) * (sde.N - 1))
    c = torch.ones_like(std)
  if tr.importance_sampling or not tr.likelihood_weighting:
    # (score*std + z)^2
    a, b, w = c * std, torch.ones_like(std), 0.5 * _vec(Z, B, dev)
  else:
    g2 = sde.sde(torch.zeros((B, 1, 1, 1), device=dev), t)[1] ** 2
    a, b, w = c, 1. / std, 0.5 * _vec(Z, B, dev) * g2
  return xt, labels, a.float().contiguous(), b.float().contiguous(), w.float().contiguous()


def get_sde_loss_fn(config, sde, train, variance='scoreflow'):
  """loss_fn(model, batch, importance_sampling, t_min=None) -> per-sample losses
  (reference losses.py:61-168, core :101-132)."""
  tr = config.training
  if variance not in ('ddpm', 'scoreflow'):
    raise ValueError(f'unknown decoder variance {variance!r}')

  def _std_normal_cdf(v):      # tanh approximation of the standard normal CDF (reference :80-81)
    return 0.5 * (1.0 + torch.tanh(np.sqrt(2.0 / np.pi) * (v + 0.044715 * (v ** 3))))

  def _discretized_gaussian_ll(x, means, log_scales):
    """log-likelihood of 8-bit data rescaled to [-1, 1] under a Gaussian discretised to bins of width 2/255, with the
    open-ended first / last bin (reference losses.py:83-100)."""
    centered, inv = x - means, torch.exp(-log_scales)
    cdf_plus, cdf_min = _std_normal_cdf(inv * (centered + 1. / 255.)), _std_normal_cdf(inv * (centered - 1. / 255.))
    floor = torch.tensor(1e-12, device=x.device)
    log_plus = torch.log(torch.max(cdf_plus, floor))
    log_one_minus_min = torch.log(torch.max(1. - cdf_min, floor))
    log_mid = torch.log(torch.max(cdf_plus - cdf_min, floor))
    return torch.where(x < -0.999, log_plus, torch.where(x > 0.999, log_one_minus_min, log_mid))

  def _reconstruction_term(model, batch, t_min, z2):
    """The decoder term of the NELBO at t_min added to the DSM losses when training.reconstruction_loss is set
    (reference losses.py:134-164): a second network evaluation at t = t_min on a fresh perturbation, the Gaussian
    decoder q(x | x_{t_min}) with mean (x_t + beta^2 score) / alpha and deviation beta [/ alpha for 'scoreflow'], scored
    either by the discretised Gaussian likelihood ('lossless' dequantisation) or by the cross-entropy minus the
    entropy of the perturbation kernel at t_min."""
    B, dev = batch.shape[0], batch.device
    eps_vec = torch.ones(B, device=dev) * t_min
    mean, std = sde.marginal_prob(batch, eps_vec)
    xt = mean + std[:, None, None, None] * z2
    score = mutils.get_score_fn(config, sde, model, train=train, continuous=tr.continuous)(xt, eps_vec)
    alpha, beta = sde.marginal_prob(torch.ones_like(batch), eps_vec)
    q_mean = xt / alpha + beta[:, None, None, None] ** 2 * score / alpha
    q_std = beta if variance == 'ddpm' else beta / torch.mean(alpha, dim=(1, 2, 3))
    if config.data.dequantization == 'lossless':
      rec = -_discretized_gaussian_ll(batch, q_mean, torch.log(q_std)[:, None, None, None]).sum(dim=(1, 2, 3))
    else:
      n_dim = float(np.prod(batch.shape[1:]))
      p_entropy = n_dim / 2. * (np.log(2 * np.pi) + 2 * torch.log(std) + 1.)
      q_recon = n_dim / 2. * (np.log(2 * np.pi) + 2 * torch.log(q_std)) + \
          0.5 / (q_std ** 2) * torch.square(batch - q_mean).sum(dim=(1, 2, 3))
      rec = q_recon - p_entropy
    if tr.reduce_mean:
      rec = rec / float(np.prod(batch.shape[1:]))
    return rec

  def loss_fn(model, batch, importance_sampling, t_min=None, injected=None):
    """`injected` = dict(u=..., z=...[, z2=...]) replaces the random draws (parity tests); z2 is the noise of
    the reconstruction term's second perturbation."""
    if t_min is None:
      t_min = sde.get_t_min(config)
    B, dev = batch.shape[0], batch.device
    if injected is not None and 'u' in injected:
      t, Z = sde.time_from_uniform(injected['u'].to(dev), t_min, importance_sampling)
    else:
      t, Z = sde.get_diffusion_time(config, B, dev, t_min, importance_sampling=importance_sampling)
    z = injected['z'].to(dev) if injected is not None and 'z' in injected else torch.randn_like(batch)
    batch = batch.float().contiguous()
    z = z.float().contiguous()
    xt, labels, a, b, w = _dsm_inputs(config, sde, batch, t, Z, z)
    model_fn = mutils.get_model_fn(model, train=train)
    out = model_fn(xt, labels)
    losses = _DsmLoss.apply(out, z, a, b, w, bool(tr.reduce_mean))
    if tr.reconstruction_loss:
      z2 = injected['z2'].to(dev).float() if injected is not None and 'z2' in injected else torch.randn_like(batch)
      losses = losses + _reconstruction_term(model, batch, t_min, z2)
    return losses

  return loss_fn


# ------------------------------------------------------------------------------------ data parallel
def _world():
  return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _bcast_device():
  return torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')


def sync_numpy_rng():
  """Give every rank rank 0's NumPy global RNG state (once, at start-up).  The soft-truncation time t_min is one
  `np.random.rand()` per step (reference sde_lib.py:200-207, losses.py:284) and must be the same on every rank; with
  identical generator states every rank draws it locally, so the per-step broadcast (and its host sync) of the
  first build is gone.  The draw sequence equals the single-process reference's."""
  if _world() == 1:
    return
  dev = _bcast_device()
  blob = pickle.dumps(np.random.get_state()) if dist.get_rank() == 0 else b''
  n = torch.tensor([len(blob)], dtype=torch.int64, device=dev)
  dist.broadcast(n, 0)
  buf = torch.zeros(int(n.item()), dtype=torch.uint8, device=dev)
  if dist.get_rank() == 0:
    buf.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
  dist.broadcast(buf, 0)
  np.random.set_state(pickle.loads(bytes(buf.cpu().numpy().tobytes())))


def sync_replicas(state):
  """Make every rank's parameters, optimizer moments, EMA shadow and step counter equal to rank 0's (once at start-up
  and after a checkpoint restore): data-parallel ranks only exchange gradients afterwards, so replicas that start
  apart stay apart.  Ranks may (and should) seed torch differently so that their time / noise / dropout draws differ."""
  if _world() == 1:
    return
  net = mutils.unwrap(state['model'])
  bufs = []
  if hasattr(net, '_flat'):
    bufs.append(net._flat)
    opt, ema = state.get('optimizer'), state.get('ema')
    if isinstance(opt, FusedAdam):
      bufs += [opt.m, opt.v]
    if ema is not None and getattr(ema, 'shadow_flat', None) is not None:
      bufs.append(ema.shadow_flat)
  else:
    bufs += [p.data for p in net.parameters()]
  for b in bufs:
    dist.broadcast(b, 0)
  meta = torch.tensor([int(state.get('step', 0)), int(getattr(state.get('optimizer'), 't', 0)),
                       int(getattr(state.get('ema'), 'num_updates', 0) or 0)], dtype=torch.int64, device=_bcast_device())
  dist.broadcast(meta, 0)
  state['step'] = int(meta[0])
  if isinstance(state.get('optimizer'), FusedAdam):
    state['optimizer'].t = int(meta[1])
  if state.get('ema') is not None and getattr(state['ema'], 'num_updates', None) is not None:
    state['ema'].num_updates = int(meta[2])
  if hasattr(net, 'sync_compute_weights'):
    net.sync_compute_weights()


def shared_t_min(sde, config):
  """The soft-truncation draw of this step (NumPy RNG, reference losses.py:284).  Under torch.distributed every rank
  draws from its own copy of rank 0's generator (see sync_numpy_rng) - no collective, no host sync."""
  return sde.get_t_min(config)


_GRAD_ALLREDUCE_BF16 = os.environ.get('ST_GRAD_ALLREDUCE', 'fp32') == 'bf16'


def sync_gradients(model):
  """Data-parallel gradient reduction: ONE all-reduce (sum) of the flat fp32 gradient buffer.  The loss is
  already divided by the world size, so the sum is the global-batch mean gradient.  ST_GRAD_ALLREDUCE=bf16 halves
  the bytes on the wire (round to bf16, reduce, widen) at bf16 rounding of the summed gradient."""
  if _world() > 1:
    g = mutils.unwrap(model)._grad
    if _GRAD_ALLREDUCE_BF16 and g.is_cuda:
      h = ops.cast(g, torch.bfloat16)
      dist.all_reduce(h)
      ops.cast(h, torch.float32, out=g)
    else:
      dist.all_reduce(g)


# ------------------------------------------------------------------------------------ captured training step
_STEP_GRAPH = os.environ.get('ST_STEP_GRAPH', '1') != '0'
_HP_BYTES = 64        # u64 dropout-seed offset | fp32 Z, A(t_min), t_min, T-t_min | fp32 lr, bc1, bc2, ema decay


class _StepGraph:
  """One optimizer step as two CUDA graphs: (A) perturb -> network forward -> loss -> explicit backward,
  (B) global-norm clip + Adam + EMA; the data-parallel gradient all-reduce runs between them.

  A B=512 step is ~1250 kernel launches issued through ~750 ctypes calls; eagerly the host cannot keep the GPU fed
  through the 8x8 / 4x4 levels of the U-Net (13 us kernels against ~25 us of Python per call).  The graphs remove
  the host from the step; what changes from step to step travels through one 64-byte pinned block that graph A
  copies to the device as its first node: the soft-truncation scalars Z, A(t_min), t_min (sde_lib.time_consts), the
  warm-up learning rate, Adam bias corrections and EMA decay (st_adam_ema `dyn`), and the dropout seed offset
  (st_set_dropout_seed_offset) - so every replay draws the seeds / scalars the eager path would have used.  The
  random draws u, z stay eager torch calls outside the graph (same generator stream as the eager path).
  """

  def __init__(self, config, sde, state, optimize_fn, batch_shape):
    self.config, self.sde = config, sde
    net = mutils.unwrap(state['model'])
    dev = net._flat.device
    self.net, self.dev = net, dev
    B = batch_shape[0]
    self.key = self._key(state, batch_shape)
    self.batch = torch.empty(batch_shape, dtype=torch.float32, device=dev)
    self.u = torch.empty(B, dtype=torch.float32, device=dev)
    self.z = torch.empty(batch_shape, dtype=torch.float32, device=dev)
    self.hp_host = torch.zeros(_HP_BYTES, dtype=torch.uint8).pin_memory()
    self.hp_dev = torch.zeros(_HP_BYTES, dtype=torch.uint8, device=dev)
    self.hp_i64, self.hp_f32 = self.hp_host[:8].view(torch.int64), self.hp_host[8:].view(torch.float32)
    dev_f = self.hp_dev[8:].view(torch.float32)
    self.dyn_time = {'Z': dev_f[0], 'A': dev_f[1], 't_min': dev_f[2], 'span': dev_f[3]} if self._dynamic_t_min() else None
    self.dyn_opt = dev_f[4:8]
    self.gvec = torch.full((B,), 1. / (B * _world()), dtype=torch.float32, device=dev)
    self.recipe = getattr(optimize_fn, 'st_recipe', None)
    self.fused_opt = (self.recipe is not None and isinstance(state['optimizer'], FusedAdam)
                      and state['ema'] is not None and getattr(state['ema'], 'owner', None) is net)
    self.calls0 = None
    self.graph_a = self.graph_b = None
    self.losses = None
    self.pinned = []
    # the per-sample losses leave the device on a side stream as soon as graph A has produced them: the host gets them
    # (and starts enqueueing the next step) while the gradient all-reduce and graph B are still running
    self.side = torch.cuda.Stream(device=dev)
    self.losses_host = torch.empty(B, dtype=torch.float32).pin_memory()
    self.ev_a, self.ev_copy = torch.cuda.Event(), torch.cuda.Event()

  @staticmethod
  def _key(state, batch_shape):
    net = mutils.unwrap(state['model'])
    opt, ema = state['optimizer'], state['ema']
    ptrs = [t.data_ptr() for t in (net._flat, net._grad, net._comp)]
    if isinstance(opt, FusedAdam):
      ptrs += [opt.m.data_ptr(), opt.v.data_ptr()]
    if ema is not None and getattr(ema, 'shadow_flat', None) is not None:
      ptrs.append(ema.shadow_flat.data_ptr())
    return (id(net), id(opt), id(ema), tuple(batch_shape), tuple(ptrs), _world())

  def _dynamic_t_min(self):
    # every VP step reads its t_min-dependent scalars from device memory, soft-truncated or not: the eager expressions
    # mix 0-dim HOST tensors (Z, A) into device arithmetic, which stream capture refuses (ImageNet32 / C4 ran eagerly)
    return isinstance(self.sde, VPSDE)

  # ---- the step body (runs once, under capture)
  def _body_a(self):
    cfg, sde, net, tr = self.config, self.sde, self.net, self.config.training
    self.hp_dev.copy_(self.hp_host, non_blocking=True)
    net._grad.zero_()
    t_min = sde.eps if self.dyn_time is not None else sde.get_t_min(cfg)     # constant when not soft-truncated
    if self.dyn_time is not None:
      t, Z = sde.time_from_uniform(self.u, t_min, tr.importance_sampling, consts=self.dyn_time)
    else:
      t, Z = sde.time_from_uniform(self.u, t_min, tr.importance_sampling)
    xt, labels, a, b, w = _dsm_inputs(cfg, sde, self.batch, t, Z, self.z)
    out, ctx = net._execute(xt.float().contiguous(), labels.float().contiguous(), record=True)
    losses, dout = _dsm_loss_raw(out, self.z, a, b, w, bool(tr.reduce_mean), gvec=self.gvec)
    net._backward(ctx, dout, need_dx=False)
    return losses

  def _body_b(self, state):
    opt, ema = state['optimizer'], state['ema']
    clip = self.recipe['grad_clip']
    gn = opt.grad_norm_sq() if clip >= 0 else None
    opt.step(clip=clip, gnorm_sq=gn, ema=ema, ema_decay=0., dyn=self.dyn_opt, advance=False)

  def capture(self, state):
    net = self.net
    ops.ColsumQueue.reserve_pinned(8)
    state['model'].train()
    self.calls0 = net._calls + 1                       # _execute advances the counter before it draws seeds
    check(lib.st_set_dropout_seed_offset(ctypes.c_void_p(self.hp_dev.data_ptr())))
    try:
      torch.cuda.synchronize()
      self.graph_a = torch.cuda.CUDAGraph()
      with torch.cuda.graph(self.graph_a):
        self.losses = self._body_a()
      self.pinned = ops.ColsumQueue.take_pinned()
      if self.fused_opt:
        self.graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_b):
          self._body_b(state)
    finally:
      check(lib.st_set_dropout_seed_offset(None))
      net._calls = self.calls0 - 1                     # the capture pass itself executed nothing

  # ---- one replayed step
  def run(self, state, batch, optimize_fn, injected, finish):
    cfg, sde, net = self.config, self.sde, self.net
    t_min = injected['t_min'] if injected is not None and 't_min' in injected else shared_t_min(sde, cfg)
    if injected is not None and 'u' in injected:
      self.u.copy_(injected['u'], non_blocking=True)
    else:
      torch.rand(self.u.shape, device=self.dev, out=self.u)
    if injected is not None and 'z' in injected:
      self.z.copy_(injected['z'], non_blocking=True)
    else:
      torch.randn(self.z.shape, device=self.dev, out=self.z)
    self.batch.copy_(batch, non_blocking=True)
    state['model'].train()                              # the flag the eager path's model_fn(train=True) leaves behind
    net._calls += 1
    self.hp_i64[0] = (net._calls - self.calls0) * net.SEED_STRIDE
    if self.dyn_time is not None:
      self.hp_f32[0:4] = torch.tensor(sde.time_consts(t_min), dtype=torch.float32)
    if self.fused_opt:
      opt, ema = state['optimizer'], state['ema']
      r = self.recipe
      if r['warmup'] > 0:
        for g in opt.param_groups:
          g['lr'] = r['lr'] * np.minimum(state['step'] / r['warmup'], 1.0)
      opt.t += 1
      lr, bc1, bc2 = opt.hyper()
      self.hp_f32[4:8] = torch.tensor([lr, bc1, bc2, ema.next_decay()], dtype=torch.float32)
    self.graph_a.replay()
    main = torch.cuda.current_stream(self.dev)
    self.ev_a.record(main)
    with torch.cuda.stream(self.side):
      self.side.wait_event(self.ev_a)
      self.losses_host.copy_(self.losses, non_blocking=True)
      self.ev_copy.record(self.side)
    sync_gradients(state['model'])
    if self.fused_opt:
      self.graph_b.replay()
      state['step'] += 1
    else:
      finish(state, state['model'], reduced=True)
    main.wait_event(self.ev_copy)                       # the next replay of graph A overwrites self.losses
    self.ev_copy.synchronize()
    return self.losses_host.clone()


def _graph_eligible(config, sde, state, batch, injected):
  if not _STEP_GRAPH or not bool(getattr(config.optim, 'cuda_graph', True)):
    return False
  net = mutils.unwrap(state['model'])
  if not (hasattr(net, '_execute') and net._flat.is_cuda and batch.is_cuda):
    return False
  if config.optim.num_micro_batch != 1 or int(getattr(config.optim, 'l2_blocks', 1) or 1) != 1 or config.training.mixed:
    return False
  if config.training.reconstruction_loss:
    return False                                      # second network evaluation at t_min: eager autograd path
  if net.drop_masks is not None or net._taps is not None:
    return False                                      # injected dropout masks / activation taps: eager parity paths
  if injected is not None and any(k not in ('t_min', 'u', 'z') for k in injected):
    return False
  return True


def get_step_fn(config, sde, train, optimize_fn=None):
  """One optimizer step (reference losses.py:218-325): returns the per-sample losses on the CPU."""
  if not config.training.continuous:
    # (the reference's discrete path cannot run either: its step_fn calls loss_fn(model, batch, importance_sampling=...,
    # t_min=...) (losses.py:287) while get_smld_loss_fn / get_ddpm_loss_fn return loss_fn(model, batch) (:181,201) -> TypeError)
    raise NotImplementedError('only continuous-time training (every BASELINE config) is built')
  loss_fn = get_sde_loss_fn(config, sde, train)
  tr = config.training
  # decide ONCE whether the caller's optimize_fn takes the EMA (ours folds it into the Adam kernel); catching a
  # TypeError around the call instead would re-run the optimizer when the error came from inside it
  try:
    sig = inspect.signature(optimize_fn).parameters
    takes_ema = 'ema' in sig or any(q.kind == inspect.Parameter.VAR_KEYWORD for q in sig.values())
  except (TypeError, ValueError):
    takes_ema = False
  ctl = {'graph': None, 'eager_calls': 0, 'synced': False, 'graph_failed': False}

  def _t_min():
    return shared_t_min(sde, config)    # once per step, shared by the micro-batches (:284)

  def _finish(state, model, reduced=False):
    if not reduced:
      sync_gradients(model)
    if takes_ema:
      optimize_fn(state['optimizer'], model.parameters(), step=state['step'], ema=state['ema'])
    else:
      optimize_fn(state['optimizer'], model.parameters(), step=state['step'])
    state['step'] += 1
    if not takes_ema and state.get('ema') is not None:
      state['ema'].update(model.parameters())

  def _zero_grad(state):
    net = mutils.unwrap(state['model'])
    if hasattr(net, '_grad'):
      # through the model: the flat gradient buffer is what the explicit backward accumulates into; torch's
      # optimizer.zero_grad(set_to_none=True) would only drop the .grad views (and hide them from AdamW / clip)
      net.zero_grad()
    else:
      state['optimizer'].zero_grad()

  def _start(state):
    if not ctl['synced']:
      ctl['synced'] = True
      if _world() > 1:
        sync_numpy_rng()
        sync_replicas(state)

  def _graph_step(state, batch, injected):
    """The captured-graph form of step_fn, or None when this call has to run eagerly."""
    if ctl['graph_failed'] or not _graph_eligible(config, sde, state, batch, injected):
      return None
    ctl['eager_calls'] += 1
    g = ctl['graph']
    if g is not None and g.key != _StepGraph._key(state, batch.shape):
      g = ctl['graph'] = None
    if g is None:
      if ctl['eager_calls'] <= 2:       # two eager steps first: lazy initialisation, allocator warm-up
        return None
      try:
        g = _StepGraph(config, sde, state, optimize_fn, tuple(batch.shape))
        g.capture(state)
        ctl['graph'] = g
      except Exception as ex:           # keep training eagerly (and say so) rather than fail the step
        ctl['graph_failed'] = True
        warnings.warn(f'soft_truncation_b200: CUDA-graph capture of the training step failed ({ex!r}); running eagerly')
        return None
    return g.run(state, batch, optimize_fn, injected, _finish)

  def step_fn(state, batch, injected=None):
    model = state['model']
    if not train:
      raise NotImplementedError('step_fn(train=False) is undefined in the reference as well (losses.py:279,293)')
    _start(state)
    if batch.is_cuda:
      losses = _graph_step(state, batch, injected)
      if losses is not None:
        return losses.cpu() if losses.is_cuda else losses
    _zero_grad(state)
    B = batch.shape[0]
    nmb = config.optim.num_micro_batch
    mb = B // nmb
    pieces = []
    t_min = injected['t_min'] if injected is not None and 't_min' in injected else _t_min()
    # L2 blocking (`optim.l2_blocks`, ours): a micro-batch runs as several image blocks, each through the whole
    # network, so that a layer's activation tensor (67 MB at 32x32x128 for 256 images) is still in the 126 MB L2 when
    # the next kernel reads it.  The block losses are scaled so that the accumulated gradient is the gradient of the
    # micro-batch mean: same step as blocks == 1 up to summation order.
    blocks = max(1, int(getattr(config.optim, 'l2_blocks', 1) or 1))
    for k in range(nmb):
      sub = -(-mb // blocks)
      for lo in range(mb * k, mb * (k + 1), sub):
        hi = min(lo + sub, mb * (k + 1))
        inj = None
        if injected is not None:
          inj = {key: v[lo:hi] for key, v in injected.items() if key in ('u', 'z')}
        losses = loss_fn(model, batch[lo:hi], importance_sampling=tr.importance_sampling, t_min=t_min, injected=inj)
        (torch.sum(losses) / (mb * _world())).backward()
        pieces.append(losses.detach())
    _finish(state, model)
    return torch.cat(pieces).cpu()

  def step_fn_mixed(state, batch, injected=None):
    """Reference losses.py:295-320.  `injected` = dict(t_min=, u=(B,), z=(B,C,H,W)) replaces the draws, rows in batch
    order (per micro-batch: the importance-sampled half, then the uniform-time half)."""
    model = state['model']
    if not train:
      raise NotImplementedError('step_fn_mixed(train=False) is undefined in the reference as well (losses.py:299,318)')
    _start(state)
    _zero_grad(state)
    B = batch.shape[0]
    nmb = config.optim.num_micro_batch
    mb = B // nmb
    half = B // (2 * nmb)
    pieces = []
    t_min = injected['t_min'] if injected is not None and 't_min' in injected else _t_min()

    def inj(lo, hi):
      if injected is None:
        return None
      return {key: v[lo:hi] for key, v in injected.items() if key in ('u', 'z')}

    for k in range(nmb):
      l_is = loss_fn(model, batch[mb * k: mb * k + half], importance_sampling=True, t_min=t_min,
                     injected=inj(mb * k, mb * k + half))
      l_dd = loss_fn(model, batch[mb * k + half: mb * (k + 1)], importance_sampling=False, t_min=t_min,
                     injected=inj(mb * k + half, mb * (k + 1)))
      wgt = tr.ddpm_weight
      if tr.balanced:
        wgt = wgt * torch.mean(l_is / l_dd).detach().item()
      losses = l_is + wgt * l_dd
      (torch.mean(losses) / _world()).backward()
      pieces.append(losses.detach())
    _finish(state, model)
    return torch.cat(pieces).cpu()

  return step_fn_mixed if tr.mixed else step_fn
