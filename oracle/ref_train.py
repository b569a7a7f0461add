"""ORACLE (test infrastructure, never shipped or timed as the product).

Plain fp32 PyTorch-on-CPU restatement of the training step and PC sampler around the score
network, with all randomness INJECTED:

* VP / VE / RVE schedules            - reference sde_lib.py:121-207, 248-332, 334-430
* time sampling + soft truncation    - reference sde_lib.py:180-207 (VP), 314-332 (VE), 421-430 (RVE)
* score wrapper                      - reference models/utils.py:128-190
* DSM loss (IS / plain / likelihood-weighted branches) - reference losses.py:101-132
* reconstruction (decoder) term    - reference losses.py:80-100, 134-164
* step_fn / step_fn_mixed            - reference losses.py:262-293, 295-320
* warm-up / clip / Adam              - reference losses.py:44-58, torch.optim.Adam
* EMA                                - reference models/ema.py:32-51
* EM / reverse-diffusion predictors, Langevin corrector, denoise step, PC loop
                                     - reference sampling.py:185-210, 263-292, 402-431

Parity pin: tests/golden/train_golden.npz, sampler_golden.npz, variants_golden.npz, deepest_golden.npz and
lossbranch_golden.npz, recon_golden.npz (made from the untouched reference by tests/golden/make_golden.py) are checked in
tests/test_oracle.py.
"""
import math

import numpy as np
import torch

from . import ref_model


# ----------------------------------------------------------------------------- schedules
class VP:
  kind = 'vpsde'

  def __init__(self, cfg, N=None):
    self.b0, self.b1 = cfg.model.beta_min, cfg.model.beta_max
    self.eps = cfg.training.truncation_time
    self.N = N or cfg.model.num_scales
    self.T = 1

  def beta(self, t):
    return self.b0 + t * (self.b1 - self.b0)

  def std(self, t):
    lmc = -0.25 * t ** 2 * (self.b1 - self.b0) - 0.5 * t * self.b0
    return torch.sqrt(1. - torch.exp(2. * lmc))

  def mean_coeff(self, t):
    return torch.exp(-0.25 * t ** 2 * (self.b1 - self.b0) - 0.5 * t * self.b0)

  def A(self, t):
    if isinstance(t, (float, int)):
      t = torch.tensor(t).float()
    ib = 0.5 * t ** 2 * (self.b1 - self.b0) + t * self.b0
    return torch.log(1. - torch.exp(-ib)) + ib

  def time_from_uniform(self, u, t_min, importance_sampling):
    if not importance_sampling:
      return u * (self.T - t_min) + t_min, 1
    Z = self.A(self.T) - self.A(t_min)
    db = self.b1 - self.b0
    t = (-self.b0 + torch.sqrt(self.b0 ** 2 + 2 * db * torch.log(1. + torch.exp(Z * u + self.A(t_min))))) / db
    return t, Z

  def t_min_from_uniform(self, cfg, U):
    if not cfg.training.st:
      return self.eps
    k = cfg.training.k
    if k == 1.0:
      return self.eps ** (1. - U)
    return self.eps / (1. - U * (1 - self.eps ** (k - 1))) ** (1. / (k - 1))

  def labels(self, t):
    return t * 999

  def score_from_out(self, out, t):
    return -out / self.std(t)[:, None, None, None]

  def prior_scale(self):
    return 1.


class VE:
  kind = 'vesde'

  def __init__(self, cfg, N=None):
    self.smin, self.smax = cfg.model.sigma_min, cfg.model.sigma_max
    self.eps = 1e-5        # get_sde does not forward truncation_time (sde_lib.py:249,439)
    self.N = N or cfg.model.num_scales
    self.T = 1
    self.discrete_sigmas = torch.exp(torch.linspace(np.log(self.smin), np.log(self.smax), self.N))

  def std(self, t):
    return self.smin * (self.smax / self.smin) ** t

  def mean_coeff(self, t):
    return torch.ones_like(t)

  def time_from_uniform(self, u, t_min, importance_sampling):
    if importance_sampling:
      Z = 2. * torch.log(self.std(torch.tensor(1.).float())) - 2. * torch.log(self.std(torch.tensor(t_min).float()))
      return t_min + ((Z * u) / (2. * (np.log(self.smax) - np.log(self.smin)))), Z
    return u * (self.T - t_min) + t_min, 1

  def t_min_from_uniform(self, cfg, U):
    return self.eps        # st is never forwarded for VE

  def labels(self, t):
    return self.std(t)

  def score_from_out(self, out, t):
    return out

  def prior_scale(self):
    return self.smax


class RVE(VE):
  kind = 'reciprocal_vesde'

  def __init__(self, cfg, N=None):
    super().__init__(cfg, N)
    eta = cfg.training.eta
    span = 1. / self.eps - 1.
    self.b = pow(eta / self.smax, 1. / span)
    self.c = self.smax ** 2 / self.b ** 2
    self.b2 = pow(1.01, -1. / (2. * span))
    self.c2 = -pow(1.01, (1. / self.eps) / span) * (eta ** 2 - self.smin ** 2)

  def std(self, t):
    t = t.double()
    return torch.sqrt(self.c * torch.pow(self.b, 2. / t) + self.c2 * torch.pow(self.b2, 2. / t)).float()

  def time_from_uniform(self, u, t_min, importance_sampling):
    return 1. / (u * (1. / t_min - 1. / self.T) + 1. / self.T), 1


def make_sde(cfg, N=None):
  return {'vpsde': VP, 'vesde': VE, 'reciprocal_vesde': RVE}[cfg.training.sde.lower()](cfg, N)


# ----------------------------------------------------------------------------- loss / step
def score_fn(sd, cfg, sde, x, t, train=False, drop_masks=None):
  out = ref_model.unet_forward(sd, cfg, x, sde.labels(t), train=train, drop_masks=drop_masks)
  return sde.score_from_out(out, t)


def reconstruction_term(sd, cfg, sde, batch, t_min, z2, train=True, variance='scoreflow'):
  """Decoder term of reference losses.py:134-164 for the injected second noise `z2`: second score evaluation at
  t = t_min, q(x | x_tmin) = N((x_t + beta^2 s) / alpha, q_std^2) with q_std = beta ('ddpm') or beta / alpha
  ('scoreflow'); 'lossless' data: minus the discretised Gaussian log-likelihood (:83-100), otherwise the Gaussian
  cross-entropy minus the entropy of the perturbation kernel at t_min (:139 re-binds `std`)."""
  B = batch.shape[0]
  eps_vec = torch.ones(B) * t_min
  alpha, beta = sde.mean_coeff(eps_vec), sde.std(eps_vec)
  x_t = alpha[:, None, None, None] * batch + beta[:, None, None, None] * z2
  score = score_fn(sd, cfg, sde, x_t, eps_vec, train=train)
  q_mean = x_t / alpha[:, None, None, None] + beta[:, None, None, None] ** 2 * score / alpha[:, None, None, None]
  q_std = beta if variance == 'ddpm' else beta / alpha
  if cfg.data.dequantization == 'lossless':
    cdf = lambda v: 0.5 * (1.0 + torch.tanh(np.sqrt(2.0 / np.pi) * (v + 0.044715 * (v ** 3))))
    inv = 1. / q_std[:, None, None, None]
    plus, minus = cdf(inv * (batch - q_mean + 1. / 255.)), cdf(inv * (batch - q_mean - 1. / 255.))
    lo = torch.tensor(1e-12)
    ll = torch.where(batch < -0.999, torch.log(torch.max(plus, lo)),
                     torch.where(batch > 0.999, torch.log(torch.max(1. - minus, lo)), torch.log(torch.max(plus - minus, lo))))
    rec = -ll.sum(dim=(1, 2, 3))
  else:
    n = float(np.prod(batch.shape[1:]))
    p_entropy = n / 2. * (np.log(2 * np.pi) + 2 * torch.log(beta) + 1.)
    rec = n / 2. * (np.log(2 * np.pi) + 2 * torch.log(q_std)) + 0.5 / q_std ** 2 * torch.square(batch - q_mean).sum(dim=(1, 2, 3)) \
        - p_entropy
  if cfg.training.reduce_mean:
    rec = rec / float(np.prod(batch.shape[1:]))
  return rec


def dsm_losses(sd, cfg, sde, batch, u, z, t_min, train=True, drop_masks=None, importance_sampling=None, z2=None,
               variance='scoreflow'):
  """Per-sample losses of reference losses.py:101-132 (+ the reconstruction term :134-164 when the config asks for it)
  for injected uniforms `u` and noise `z` (`z2`: the second perturbation of the reconstruction term).
  `importance_sampling` (the loss_fn ARGUMENT, :101,112) only selects how the times are drawn; the loss formula is keyed
  on the config (:122) - the two differ inside step_fn_mixed."""
  tr = cfg.training
  t, Z = sde.time_from_uniform(u, t_min, tr.importance_sampling if importance_sampling is None else importance_sampling)
  std = sde.std(t)
  x_t = sde.mean_coeff(t)[:, None, None, None] * batch + std[:, None, None, None] * z
  score = score_fn(sd, cfg, sde, x_t, t, train=train, drop_masks=drop_masks)
  reduce = (lambda v: v.mean(dim=-1)) if tr.reduce_mean else (lambda v: 0.5 * v.sum(dim=-1))
  if tr.reconstruction_loss:
    assert tr.importance_sampling or not tr.likelihood_weighting
    sq = torch.square(score * std[:, None, None, None] + z)
    return 0.5 * Z * reduce(sq.reshape(sq.shape[0], -1)) + \
        reconstruction_term(sd, cfg, sde, batch, t_min, z2, train=train, variance=variance)
  if tr.importance_sampling or not tr.likelihood_weighting:
    sq = torch.square(score * std[:, None, None, None] + z)
    return 0.5 * Z * reduce(sq.reshape(sq.shape[0], -1))
  if sde.kind == 'vpsde':
    g2 = sde.beta(t)                       # diffusion^2 of the VP SDE (sde_lib.py:145-149)
  elif sde.kind == 'vesde':
    # diffusion = sigma(t) sqrt(2 ln(sigma_max / sigma_min)) (sde_lib.py:263-268)
    g2 = sde.std(t) ** 2 * (2. * (np.log(sde.smax) - np.log(sde.smin)))
  else:
    raise NotImplementedError
  sq = torch.square(score + z / std[:, None, None, None])
  return 0.5 * Z * reduce(sq.reshape(sq.shape[0], -1)) * g2


class TrainState:
  """Parameters + Adam moments + EMA shadows for the oracle's step."""

  def __init__(self, sd):
    self.sd = {k: v.clone() for k, v in sd.items()}
    # everything but the `sigmas` buffer and the frozen GaussianFourierProjection.W
    # (`all_modules.0.W`, models/layerspp.py:50) is trained
    self.trainable = [k for k in self.sd
                      if k != 'sigmas' and not (k.endswith('.W') and k.count('.') == 2)]
    self.m = {k: torch.zeros_like(self.sd[k]) for k in self.trainable}
    self.v = {k: torch.zeros_like(self.sd[k]) for k in self.trainable}
    self.ema = {k: self.sd[k].clone() for k in self.trainable}
    self.step = 0
    self.ema_updates = 0


def train_step(state, cfg, sde, batch, u, z, U_tmin, train=True, drop_masks=None):
  """One optimizer step (reference losses.py:262-293, or :295-320 when training.mixed) with injected randomness.
  Returns (per-sample losses, {name: grad})."""
  o = cfg.optim
  for k in state.trainable:
    state.sd[k].requires_grad_(True)
    state.sd[k].grad = None
  t_min = sde.t_min_from_uniform(cfg, U_tmin)
  if cfg.training.mixed:
    # step_fn_mixed (:295-320), one micro-batch: first half of the batch with importance-sampled times, second half
    # with uniform times, combined per sample pair
    assert o.num_micro_batch == 1 and drop_masks is None
    h = batch.shape[0] // 2
    l_is = dsm_losses(state.sd, cfg, sde, batch[:h], u[:h], z[:h], t_min, train=train, importance_sampling=True)
    l_dd = dsm_losses(state.sd, cfg, sde, batch[h:], u[h:], z[h:], t_min, train=train, importance_sampling=False)
    wgt = cfg.training.ddpm_weight
    if cfg.training.balanced:
      wgt = wgt * torch.mean(l_is / l_dd).detach().item()
    losses = l_is + wgt * l_dd
  else:
    losses = dsm_losses(state.sd, cfg, sde, batch, u, z, t_min, train=train, drop_masks=drop_masks)
  torch.mean(losses).backward()
  grads = {k: state.sd[k].grad.detach().clone() for k in state.trainable}
  with torch.no_grad():
    lr = o.lr * np.minimum(state.step / o.warmup, 1.0) if o.warmup > 0 else o.lr
    if o.grad_clip >= 0:
      total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
      coef = torch.clamp(o.grad_clip / (total + 1e-6), max=1.0)
    else:
      coef = torch.tensor(1.)
    b1, b2 = o.beta1, 0.999
    n = state.step + 1
    for k in state.trainable:
      g = grads[k] * coef
      if o.weight_decay:
        g = g + o.weight_decay * state.sd[k]
      state.m[k].mul_(b1).add_(g, alpha=1 - b1)
      state.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
      denom = (state.v[k].sqrt() / math.sqrt(1 - b2 ** n)).add_(o.eps)
      state.sd[k].data.addcdiv_(state.m[k], denom, value=-float(lr) / (1 - b1 ** n))
    state.step += 1
    state.ema_updates += 1
    d = min(cfg.model.ema_rate, (1 + state.ema_updates) / (10 + state.ema_updates))
    for k in state.trainable:
      state.ema[k].sub_((1.0 - d) * (state.ema[k] - state.sd[k].data))
  for k in state.trainable:
    state.sd[k].requires_grad_(False)
  return losses.detach(), grads


# ----------------------------------------------------------------------------- PC sampler
def pc_sample(sd, cfg, sde, x_T, noises, eps, predictor='euler_maruyama', corrector='none',
              snr=0.16, n_steps=1, denoise=True, trace=None):
  """Reference sampling.py:410-431 with injected prior draw `x_T` and per-step noises.

  `noises[i]` is the predictor noise of step i; with the Langevin corrector `noises[i]` is a
  tuple (corrector noises..., predictor noise).  Returns x (network range, before the inverse
  scaler).  `trace`, if a list, receives x after every step.
  """
  with torch.no_grad():
    x = x_T.clone()
    B = x.shape[0]
    ts = torch.linspace(sde.T, eps, sde.N)
    x_mean = x
    for i in range(sde.N):
      t = torch.ones(B) * ts[i]
      step_noise = noises[i]
      if corrector == 'langevin':
        *cn, pn = step_noise
        if sde.kind == 'vpsde':
          idx = (t * (sde.N - 1) / sde.T).long()
          alpha = (1. - torch.linspace(sde.b0 / sde.N, sde.b1 / sde.N, sde.N))[idx]
        else:
          alpha = torch.ones_like(t)
        for j in range(n_steps):
          grad = score_fn(sd, cfg, sde, x, t)
          noise = cn[j]
          gnorm = torch.norm(grad.reshape(B, -1), dim=-1).mean()
          nnorm = torch.norm(noise.reshape(B, -1), dim=-1).mean()
          step = (snr * nnorm / gnorm) ** 2 * 2 * alpha
          x_mean = x + step[:, None, None, None] * grad
          x = x_mean + torch.sqrt(step * 2)[:, None, None, None] * noise
      else:
        pn = step_noise
      if predictor == 'euler_maruyama':
        dt = -1. / sde.N
        score = score_fn(sd, cfg, sde, x, t)
        if sde.kind == 'vpsde':
          beta = sde.beta(t)
          drift = -0.5 * beta[:, None, None, None] * x
          g = torch.sqrt(beta)
        else:
          drift = torch.zeros_like(x)
          g = sde.std(t) * torch.sqrt(torch.tensor(2 * (np.log(sde.smax) - np.log(sde.smin))))
        drift = drift - g[:, None, None, None] ** 2 * score * 1.0
        x_mean = x + drift * dt
        x = x_mean + g[:, None, None, None] * np.sqrt(-dt) * pn
      elif predictor == 'reverse_diffusion':
        idx = (t * (sde.N - 1) / sde.T).long()
        if sde.kind == 'vpsde':
          dbeta = torch.linspace(sde.b0 / sde.N, sde.b1 / sde.N, sde.N)[idx]
          f = torch.sqrt(1. - dbeta)[:, None, None, None] * x - x
          G = torch.sqrt(dbeta)
        else:
          sig = sde.discrete_sigmas[idx]
          prev = torch.where(idx == 0, torch.zeros_like(t), sde.discrete_sigmas[idx - 1])
          f = torch.zeros_like(x)
          G = torch.sqrt(sig ** 2 - prev ** 2)
        rev_f = f - G[:, None, None, None] ** 2 * score_fn(sd, cfg, sde, x, t) * 1.0
        x_mean = x - rev_f
        x = x_mean + G[:, None, None, None] * pn
      elif predictor != 'none':
        raise NotImplementedError(predictor)
      if trace is not None:
        trace.append(x.clone())
    # final denoise: reverse-diffusion predictor, probability flow, t = sde.eps -> 0
    xin = x_mean if denoise else x
    t = torch.ones(B) * sde.eps
    if sde.kind == 'vpsde':
      G = torch.sqrt((t - 0.) * sde.beta(t))
      f = torch.sqrt(1. - G ** 2)[:, None, None, None] * xin - xin
    else:
      G = torch.sqrt(sde.std(t) ** 2 - sde.std(torch.zeros_like(t)) ** 2)
      f = torch.zeros_like(xin)
    rev_f = f - G[:, None, None, None] ** 2 * score_fn(sd, cfg, sde, xin, t) * 0.5
    return xin - rev_f
