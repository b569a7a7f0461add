"""Forward/reverse SDE schedules of the Soft-Truncation hot path (host side).

Mirrors the public surface of the reference's `sde_lib` (sde_lib.py:8-445):
`SDE`, `VPSDE`, `subVPSDE`, `VESDE`, `reciprocal_VESDE`, `get_sde`, with the same
method names, argument meaning and quirks.  Everything
here is O(batch) scalar math: it stays in PyTorch on whatever device `t` lives on; the
O(B*C*H*W) tensor work that consumes these scalars runs in the CUDA library
(csrc/elementwise.cu: st_perturb / st_dsm_loss / st_em_step / st_rd_step / ...).

Arithmetic is written in the same operation order as the reference so that the
per-sample scalars (t, std, Z, beta, G) agree bit-for-bit in fp32.
"""
import abc
import math

import numpy as np
import torch


def _bcast(v):
  """(B,) -> (B,1,1,1)"""
  return v[:, None, None, None]


def _soft_truncation_draw(eps, k):
  """One draw of the smallest diffusion time (sde_lib.py:200-207).

  k == 1: eps**(1-U);  k != 1: eps / (1 - U (1 - eps**(k-1)))**(1/(k-1)).  U comes from
  NumPy's global RNG, exactly one draw per call, like the reference.
  """
  u = np.random.rand()
  if k == 1.0:
    return eps ** (1. - u)
  return eps / (1. - u * (1 - eps ** (k - 1))) ** (1. / (k - 1))


class ReverseSDE:
  """Reverse-time SDE/ODE built by `SDE.reverse` (sde_lib.py:75-119).

  drift_rev = f - g^2 * score * w,  w = 1/2 for the probability-flow ODE and
  (1 + lambda^2)/2 otherwise; diffusion_rev = lambda * g.
  """

  def __init__(self, forward_sde, score_fn, probability_flow, lambda_):
    assert probability_flow == (lambda_ == 0.)
    self.forward_sde = forward_sde
    self.score_fn = score_fn
    self.N = forward_sde.N
    self.probability_flow = probability_flow
    self.lambda_ = lambda_
    self.weight = 0.5 if probability_flow else 0.5 * (1. + lambda_ ** 2)

  @property
  def T(self):
    return self.forward_sde.T

  def sde(self, x, t):
    drift, diffusion = self.forward_sde.sde(x, t)
    score = self.score_fn(x, t)
    drift = drift - _bcast(diffusion) ** 2 * score * self.weight
    return drift, self.lambda_ * diffusion

  def discretize(self, x, t, next_t=None):
    f, G = self.forward_sde.discretize(x, t, next_t)
    rev_f = f - _bcast(G) ** 2 * self.score_fn(x, t) * self.weight
    return rev_f, self.lambda_ * G


class SDE(abc.ABC):
  """Abstract forward SDE (sde_lib.py:8-73)."""

  def __init__(self, N):
    super().__init__()
    self.N = N

  @property
  @abc.abstractmethod
  def T(self):
    """End time."""

  @abc.abstractmethod
  def sde(self, x, t):
    """(drift, diffusion) at (x, t)."""

  @abc.abstractmethod
  def marginal_prob(self, x, t):
    """(mean, std) of p_t(x_t | x_0 = x)."""

  @abc.abstractmethod
  def prior_sampling(self, shape):
    """One draw from p_T, on the CPU like the reference."""

  @abc.abstractmethod
  def prior_logp(self, z):
    """log p_T(z)."""

  def get_diffusion_time(self, config):
    pass

  def discretize(self, x, t, next_t=None):
    """Euler-Maruyama default: f = drift/N, G = diffusion * sqrt(1/N) (sde_lib.py:55-73)."""
    dt = 1 / self.N
    drift, diffusion = self.sde(x, t)
    return drift * dt, diffusion * float(torch.sqrt(torch.tensor(dt)))      # fp32 sqrt on the host: graph-capturable

  def reverse(self, score_fn, probability_flow=False, lambda_=1.):
    return ReverseSDE(self, score_fn, probability_flow, lambda_)

  def table(self, name, device):
    """Device-resident copy of one of the per-step tables (`discrete_sigmas`, `discrete_betas`, `alphas`, ...).  The
    reference re-uploads (or CPU-indexes) them on every call (sde_lib.py:170-171,295); a sampler step captured in a
    CUDA graph cannot copy from pageable host memory, so the copy is made once per device and kept."""
    cache = self.__dict__.setdefault('_table_cache', {})
    src = getattr(self, name)
    key = (name, str(device))
    t = cache.get(key)
    if t is None or t.shape != src.shape:
      t = cache[key] = src.to(device)
    return t


def _gaussian_prior_logp(z, sigma):
  n = np.prod(z.shape[1:])
  return -n / 2. * np.log(2 * np.pi * sigma ** 2) - torch.sum(z ** 2, dim=(1, 2, 3)) / (2 * sigma ** 2)


class _LinearBeta:
  """beta(t) = beta_0 + t (beta_1 - beta_0) helpers shared by VP / sub-VP."""

  def _beta(self, t):
    return self.beta_0 + t * (self.beta_1 - self.beta_0)

  def _log_mean_coeff(self, t):
    return -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0


class VPSDE(_LinearBeta, SDE):
  """Variance-preserving SDE with soft truncation (sde_lib.py:121-207)."""

  def __init__(self, truncation_time=1e-5, beta_min=0.1, beta_max=20, N=1000):
    super().__init__(N)
    self.beta_0, self.beta_1, self.eps = beta_min, beta_max, truncation_time
    self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
    self.alphas = 1. - self.discrete_betas
    self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
    self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
    self.sqrt_1m_alphas_cumprod = torch.sqrt(1. - self.alphas_cumprod)

  @property
  def T(self):
    return 1

  def sde(self, x, t):
    beta_t = self._beta(t)
    return -0.5 * _bcast(beta_t) * x, torch.sqrt(beta_t)

  def marginal_prob(self, x, t):
    lmc = self._log_mean_coeff(t)
    return torch.exp(_bcast(lmc)) * x, torch.sqrt(1. - torch.exp(2. * lmc))

  def prior_sampling(self, shape):
    return torch.randn(*shape)

  def prior_logp(self, z):
    return _gaussian_prior_logp(z, 1.)

  def discretize(self, x, t, next_t=None):
    """DDPM grid step, or the exact one-interval step used by the denoiser (sde_lib.py:166-178)."""
    if next_t is None:
      idx = (t * (self.N - 1) / self.T).long()
      beta = self.table('discrete_betas', x.device)[idx]
      alpha = self.table('alphas', x.device)[idx]
      return _bcast(torch.sqrt(alpha)) * x - x, torch.sqrt(beta)
    G = torch.sqrt((t - next_t) * self._beta(t))
    return _bcast(torch.sqrt(1. - G ** 2)) * x - x, G

  def integral_beta(self, t):
    return 0.5 * t ** 2 * (self.beta_1 - self.beta_0) + t * self.beta_0

  def antiderivative(self, t, stabilizing_constant=0.):
    if isinstance(t, (float, int)):
      t = torch.tensor(t).float()
    ib = self.integral_beta(t)
    return torch.log(1. - torch.exp(-ib) + stabilizing_constant) + ib

  def normalizing_constant(self, t_min):
    return self.antiderivative(self.T) - self.antiderivative(t_min)

  def importance_time_from_uniform(self, u, t_min, consts=None):
    """Inverse-CDF map u in [0,1) -> t for the importance-sampled time (sde_lib.py:191-196).  `consts` (optional):
    the t_min-dependent scalars as DEVICE 0-dim fp32 tensors (see `time_consts`) - same values, same operation
    order, but a captured CUDA graph re-reads them at every replay instead of baking one step's t_min in."""
    Z = self.normalizing_constant(t_min) if consts is None else consts['Z']
    A = self.antiderivative(t_min) if consts is None else consts['A']
    db = self.beta_1 - self.beta_0
    t = (-self.beta_0 + torch.sqrt(self.beta_0 ** 2 + 2 * db *
                                   torch.log(1. + torch.exp(Z * u + A)))) / db
    return t, Z.detach()

  def time_consts(self, t_min):
    """Host-side fp32 values of everything in `time_from_uniform` that depends on t_min: [Z, A(t_min), t_min, T - t_min]
    (Z, A evaluated exactly as the reference does, as fp32 tensor ops on the host)."""
    return [float(self.normalizing_constant(t_min)), float(self.antiderivative(t_min)), float(t_min), float(self.T - t_min)]

  def time_from_uniform(self, u, t_min, importance_sampling=True, consts=None):
    """Diffusion times for given uniforms `u` (the deterministic part of get_diffusion_time)."""
    if importance_sampling:
      return self.importance_time_from_uniform(u, t_min, consts)
    if consts is not None:
      return u * consts['span'] + consts['t_min'], 1
    return u * (self.T - t_min) + t_min, 1

  def get_diffusion_time(self, config, batch_size, batch_device, t_min, importance_sampling=True):
    return self.time_from_uniform(torch.rand(batch_size, device=batch_device), t_min, importance_sampling)

  def get_t_min(self, config):
    if config.training.st:
      return _soft_truncation_draw(self.eps, config.training.k)
    return self.eps


class subVPSDE(_LinearBeta, SDE):
  """sub-VP SDE (sde_lib.py:209-246); kept for registry completeness."""

  def __init__(self, truncation_time=1e-5, beta_min=0.1, beta_max=20, N=1000):
    super().__init__(N)
    self.beta_0, self.beta_1 = beta_min, beta_max

  @property
  def T(self):
    return 1

  def sde(self, x, t):
    beta_t = self._beta(t)
    discount = 1. - torch.exp(-2 * self.beta_0 * t - (self.beta_1 - self.beta_0) * t ** 2)
    return -0.5 * _bcast(beta_t) * x, torch.sqrt(beta_t * discount)

  def marginal_prob(self, x, t):
    lmc = self._log_mean_coeff(t)
    return _bcast(torch.exp(lmc)) * x, 1 - torch.exp(2. * lmc)

  def prior_sampling(self, shape, data_mean=None):
    return torch.randn(*shape)

  def prior_logp(self, z):
    return _gaussian_prior_logp(z, 1.)


class VESDE(SDE):
  """Variance-exploding SDE (sde_lib.py:248-332)."""

  def __init__(self, sigma_min=0.01, sigma_max=50, N=1000, truncation_time=1e-5):
    super().__init__(N)
    self.sigma_min, self.sigma_max, self.eps = sigma_min, sigma_max, truncation_time
    self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))

  @property
  def T(self):
    return 1

  def _sigma(self, t):
    return self.sigma_min * (self.sigma_max / self.sigma_min) ** t

  def sde(self, x, t):
    # the reference builds this scalar as a device tensor on every call (sde_lib.py:255); a host->device copy from
    # pageable memory cannot be captured in a CUDA graph, so the fp32 value is computed once on the host
    growth = float(torch.sqrt(torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min)))))
    return torch.zeros_like(x), self._sigma(t) * growth

  def marginal_prob(self, x, t):
    return x, self._sigma(t)

  def prior_sampling(self, shape):
    return torch.randn(*shape) * self.sigma_max

  def prior_logp(self, z):
    return _gaussian_prior_logp(z, self.sigma_max)

  def discretize(self, x, t, next_t=None):
    """SMLD grid step or the exact step to `next_t == 0` (sde_lib.py:288-304)."""
    if next_t is None:
      idx = (t * (self.N - 1) / self.T).long()
      table = self.table('discrete_sigmas', t.device)   # the reference indexes the CPU table (quirk 21)
      sigma = table[idx]
      prev_sigma = torch.where(idx == 0, torch.zeros_like(t), table[idx - 1])
    else:
      if next_t[0].item() != 0.:
        raise NotImplementedError
      sigma, prev_sigma = self._sigma(t), self._sigma(next_t)
    return torch.zeros_like(x), torch.sqrt(sigma ** 2 - prev_sigma ** 2)

  def antiderivative(self, t):
    if isinstance(t, (float, int)):
      t = torch.tensor(t).float()
    return 2. * torch.log(self._sigma(t))

  def normalizing_constant(self, t_min):
    return self.antiderivative(self.T) - self.antiderivative(t_min)

  def get_diffusion_time(self, config, batch_size, batch_device, t_min, importance_sampling=None):
    if importance_sampling is None:
      importance_sampling = config.training.importance_sampling
    return self.time_from_uniform(torch.rand(batch_size, device=batch_device), t_min, importance_sampling)

  def time_from_uniform(self, u, t_min, importance_sampling=False):
    if importance_sampling:
      Z = self.normalizing_constant(t_min)
      return t_min + ((Z * u) / (2. * (np.log(self.sigma_max) - np.log(self.sigma_min)))), Z.detach()
    return u * (self.T - t_min) + t_min, 1

  def get_t_min(self, config, st=False):
    # NB the training loop calls get_t_min(config) -> st stays False.
    return _soft_truncation_draw(self.eps, config.training.k) if st else self.eps


class reciprocal_VESDE(SDE):
  """Reciprocal VE SDE, sigma^2(t) = c b^(2/t) + c2 b2^(2/t) (sde_lib.py:334-430)."""

  def __init__(self, eta=1e-5, sigma_min=0.01, sigma_max=50, N=1000):
    super().__init__(N)
    self.sigma_min, self.sigma_max, self.eta, self.eps = sigma_min, sigma_max, eta, 1e-5
    span = 1. / self.eps - 1.
    self.base_sigma = pow(eta / sigma_max, 1. / span)
    self.const = sigma_max ** 2 / self.base_sigma ** 2
    self.base_sigma_2 = pow(1.01, -1. / (2. * span))
    self.const_2 = -pow(1.01, (1. / self.eps) / span) * (eta ** 2 - sigma_min ** 2)

    self.t_0 = torch.tensor(self.get_time())
    self.sigma_0 = torch.sqrt(self.const * torch.pow(self.base_sigma, 2. * self.t_0)
                              + self.const_2 * torch.pow(self.base_sigma_2, 2. * self.t_0))
    self.k_1 = -self.t_0 * self.sigma_0 / np.log(self.base_sigma)
    self.k_2 = -self.k_1 / self.sigma_0
    self.constant_ = 1. / torch.log(self.sigma_0 / self.sigma_max)
    self.c_1_ = (self.sigma_0 / np.log(self.base_sigma) * (np.log(self.sigma_0) - np.log(self.sigma_max))
                 / (self.t_0 - 1. / self.T))
    self.c_2_ = self.sigma_0 - (self.c_1_ / self.sigma_0)
    self.c_2__ = np.log(self.sigma_0) + self.c_1_ / self.sigma_0
    self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))

  @property
  def T(self):
    return 1

  def sde(self, x, t):
    g2 = (-(2. * self.const * np.log(self.base_sigma)) * torch.pow(self.base_sigma, 2. / t) / (t ** 2)
          + (2. * self.const_2 * np.log(self.base_sigma_2) * torch.pow(self.base_sigma_2, 2. / t) / (t ** 2)))
    return torch.zeros_like(x), torch.sqrt(g2)

  def marginal_prob(self, x, t):
    # evaluated in float64 (the b^(2/t) terms underflow fp32 near t_min); the reference does it on
    # the CPU (sde_lib.py:381-385), here it stays on t's device - same IEEE double arithmetic.
    t64 = t.double()
    std = torch.sqrt(self.const * torch.pow(self.base_sigma, 2. / t64)
                     + self.const_2 * torch.pow(self.base_sigma_2, 2. / t64))
    return x, std.float().to(x.device)

  def prior_sampling(self, shape):
    return torch.randn(*shape) * self.sigma_max

  def prior_logp(self, z):
    return _gaussian_prior_logp(z, self.sigma_max)

  def discretize(self, x, t, next_t=None):
    # The reference dereferences next_t.type with next_t=None (sde_lib.py:404): PC sampling with
    # this SDE is broken as shipped.  Same error class here.
    if next_t is None:
      raise AttributeError("'NoneType' object has no attribute 'type'")
    sigma = self.marginal_prob(x, t)[1]
    next_sigma = self.marginal_prob(x, next_t)[1]
    return torch.zeros_like(x), torch.sqrt(sigma ** 2 - next_sigma ** 2)

  def get_time(self, sigma_level=0.01):
    return (np.log((-self.sigma_min ** 2 + self.eta ** 2 + sigma_level ** 2) / self.const)
            / (2. * np.log(self.base_sigma)))

  def transform(self, sigmas):
    return ((sigmas > 0.01) * torch.log(sigmas)
            + (sigmas < 0.01) * (-self.c_1_ / (sigmas + 1e-4) + self.c_2__))

  def get_diffusion_time(self, config, batch_size, batch_device, t_min, importance_sampling=False):
    return self.time_from_uniform(torch.rand(batch_size, device=batch_device), t_min)

  def time_from_uniform(self, u, t_min, importance_sampling=False):
    inv_t = u * (1. / t_min - 1. / self.T) + 1. / self.T
    return 1. / inv_t, 1

  def get_t_min(self, config, st=False):
    if st:
      return 1. / (np.random.rand() * (1. / self.eps - 1. / self.T) + 1. / self.T)
    return self.eps


def get_sde(config, state=None):
  """Factory keyed on `config.training.sde` (sde_lib.py:433-445)."""
  kind = config.training.sde.lower()
  m = config.model
  if kind == 'vpsde':
    return VPSDE(truncation_time=config.training.truncation_time, beta_min=m.beta_min,
                 beta_max=m.beta_max, N=m.num_scales)
  if kind == 'subvpsde':
    return subVPSDE(truncation_time=config.training.truncation_time, beta_min=m.beta_min,
                    beta_max=m.beta_max, N=m.num_scales)
  if kind == 'vesde':
    return VESDE(sigma_min=m.sigma_min, sigma_max=m.sigma_max, N=m.num_scales)
  if kind == 'reciprocal_vesde':
    return reciprocal_VESDE(sigma_min=m.sigma_min, sigma_max=m.sigma_max, N=m.num_scales,
                            eta=config.training.eta)
  raise NotImplementedError(f"SDE {config.training.sde} unknown.")
