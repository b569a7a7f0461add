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
import math

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
    for e in self.model.store.entries:
      p = self.model._params[e.name]
      if e.trainable:
        self.state[p] = {'step': torch.tensor(float(self.t)), 'exp_avg': _params._logical_view(self.m, e),
                         'exp_avg_sq': _params._logical_view(self.v, e)}

  def zero_grad(self, set_to_none=False):
    self.model.zero_grad()

  def grad_norm_sq(self):
    """Sum of squared gradients, left on the device."""
    self.gnorm_sq.zero_()
    g = self.model._grad
    check(lib.st_sumsq(ops.ptr(g), g.numel(), ops.ptr(self.gnorm_sq), ops.stream()))
    return self.gnorm_sq

  @torch.no_grad()
  def step(self, closure=None, clip=-1., gnorm_sq=None, ema=None, ema_decay=0.):
    m = self.model
    if not m._flat.is_cuda:
      raise RuntimeError('FusedAdam needs the parameters on a CUDA device')
    g = self.param_groups[0]
    self.t += 1
    b1, b2 = g['betas']
    p16 = m._comp if m._comp is not m._flat else None
    check(lib.st_adam_ema(ops.ptr(m._flat), ops.ptr(m._grad), ops.ptr(self.m), ops.ptr(self.v),
                          ops.ptr(ema.shadow_flat) if ema is not None else None,
                          ops.ptr(ema.mask) if ema is not None else None, ops.ptr(p16), m._flat.numel(),
                          ops.ptr(gnorm_sq), float(clip), float(g['lr']), float(b1), float(b2), float(g['eps']),
                          float(g['weight_decay']), 1. - b1 ** self.t, 1. - b2 ** self.t, float(ema_decay),
                          ops.stream()))
    for st in self.state.values():
      st['step'] = torch.tensor(float(self.t))

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

  return optimize_fn


# ------------------------------------------------------------------------------------ loss
class _DsmLoss(torch.autograd.Function):
  """losses[n] = w[n] * red_i (a[n]*out[n,i] + b[n]*z[n,i])^2 (st_dsm_loss); backward returns d(out)."""

  @staticmethod
  def forward(ctx, out, z, a, b, w, reduce_mean):
    B = out.shape[0]
    D = out[0].numel()
    out, z = out.contiguous(), z.contiguous()
    losses = torch.empty(B, dtype=torch.float32, device=out.device)
    check(lib.st_dsm_loss(ops.ptr(out), ops.ptr(z), ops.ptr(a), ops.ptr(b), ops.ptr(w), ops.ptr(losses), None, None,
                          B, D, int(reduce_mean), ops.stream()))
    ctx.save_for_backward(out, z, a, b, w)
    ctx.reduce_mean = reduce_mean
    return losses

  @staticmethod
  def backward(ctx, dl):
    out, z, a, b, w = ctx.saved_tensors
    B, D = out.shape[0], out[0].numel()
    dout = torch.empty_like(out)
    scratch = torch.empty(B, dtype=torch.float32, device=out.device)
    check(lib.st_dsm_loss(ops.ptr(out), ops.ptr(z), ops.ptr(a), ops.ptr(b), ops.ptr(w), ops.ptr(scratch),
                          ops.ptr(dout), ops.ptr(dl.float().contiguous()), B, D, int(ctx.reduce_mean), ops.stream()))
    return dout, None, None, None, None, None


def _vec(v, B, device):
  if torch.is_tensor(v):
    return v.to(device=device, dtype=torch.float32).expand(B).contiguous()
  return torch.full((B,), float(v), dtype=torch.float32, device=device)


def get_sde_loss_fn(config, sde, train, variance='scoreflow'):
  """loss_fn(model, batch, importance_sampling, t_min=None) -> per-sample losses
  (reference losses.py:61-168, core :101-132)."""
  tr = config.training
  if tr.reconstruction_loss:
    raise NotImplementedError('training.reconstruction_loss (off in every BASELINE config) is not built')

  def loss_fn(model, batch, importance_sampling, t_min=None, injected=None):
    """`injected` = dict(u=..., z=...) replaces the two random draws (parity tests)."""
    if t_min is None:
      t_min = sde.get_t_min(config)
    B, dev = batch.shape[0], batch.device
    if injected is not None and 'u' in injected:
      t, Z = sde.time_from_uniform(injected['u'].to(dev), t_min, importance_sampling)
    else:
      t, Z = sde.get_diffusion_time(config, B, dev, t_min, importance_sampling=importance_sampling)
    z = injected['z'].to(dev) if injected is not None and 'z' in injected else torch.randn_like(batch)
    unit = torch.ones((B, 1, 1, 1), device=dev)
    mean_coeff, std = sde.marginal_prob(unit, t)
    mean_coeff = mean_coeff.reshape(B).float().contiguous()
    std = std.float().contiguous()
    batch = batch.float().contiguous()
    z = z.float().contiguous()
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
    model_fn = mutils.get_model_fn(model, train=train)
    out = model_fn(xt, labels)
    if tr.importance_sampling or not tr.likelihood_weighting:
      # (score*std + z)^2
      a, b, w = c * std, torch.ones_like(std), 0.5 * _vec(Z, B, dev)
    else:
      g2 = sde.sde(torch.zeros((B, 1, 1, 1), device=dev), t)[1] ** 2
      a, b, w = c, 1. / std, 0.5 * _vec(Z, B, dev) * g2
    return _DsmLoss.apply(out, z, a.float().contiguous(), b.float().contiguous(), w.float().contiguous(),
                          bool(tr.reduce_mean))

  return loss_fn


def _world():
  return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shared_t_min(sde, config):
  """The soft-truncation draw of this step (NumPy RNG, reference losses.py:284), identical on every rank:
  rank 0 draws, the others receive it."""
  t_min = sde.get_t_min(config)
  if _world() > 1:
    box = [t_min]
    dist.broadcast_object_list(box, src=0)
    t_min = box[0]
  return t_min


def sync_gradients(model):
  """Data-parallel gradient reduction: ONE all-reduce (sum) of the flat fp32 gradient buffer.  The loss is
  already divided by the world size, so the sum is the global-batch mean gradient."""
  if _world() > 1:
    dist.all_reduce(mutils.unwrap(model)._grad)


def get_step_fn(config, sde, train, optimize_fn=None):
  """One optimizer step (reference losses.py:218-325): returns the per-sample losses on the CPU."""
  if not config.training.continuous:
    raise NotImplementedError('only continuous-time training (every BASELINE config) is built')
  loss_fn = get_sde_loss_fn(config, sde, train)
  tr = config.training

  def _t_min():
    return shared_t_min(sde, config)    # once per step, shared by the micro-batches (:284)

  def _finish(state, model):
    sync_gradients(model)
    try:
      optimize_fn(state['optimizer'], model.parameters(), step=state['step'], ema=state['ema'])
      fused_ema = True
    except TypeError:                                  # a foreign optimize_fn without the `ema` keyword
      optimize_fn(state['optimizer'], model.parameters(), step=state['step'])
      fused_ema = False
    state['step'] += 1
    if not fused_ema:
      state['ema'].update(model.parameters())

  def step_fn(state, batch, injected=None):
    model = state['model']
    optimizer = state['optimizer']
    if not train:
      raise NotImplementedError('step_fn(train=False) is undefined in the reference as well (losses.py:279,293)')
    optimizer.zero_grad()
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
    optimizer = state['optimizer']
    if not train:
      raise NotImplementedError('step_fn_mixed(train=False) is undefined in the reference as well (losses.py:299,318)')
    optimizer.zero_grad()
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
