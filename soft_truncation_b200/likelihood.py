"""Likelihood (bits/dim) and NELBO estimators on the B200 path - SURVEY 8(f)2.

Same surface as the reference's `likelihood.py`:

* `get_div_fn(fn)`                          - Hutchinson-Skilling trace estimator (likelihood.py:27-37)
* `get_likelihood_fn(config, sde, inverse_scaler, ...)` -> `likelihood_fn(model, data, logdet=0., eps=1e-5,
  mode='correct')` -> `(bpd, z, nfe)`      - probability-flow ODE + divergence (likelihood.py:41-134)
* `get_elbo_fn(config, sde, inverse_scaler, ...)` -> `loss_fn(model, batch, logdet=0., eps=1e-5)` ->
  `(nelbo_bpd, residual_bpd)`                (likelihood.py:136-208)
* `get_likelihood_residual_fn(config, sde, score_fn, variance)` (likelihood.py:210-313)

What changes is where the work runs.  The score network's vector-Jacobian product eps^T d(drift)/dx goes through the
explicit backward of `models/ncsnpp.py` restricted to the input gradient (`ops.input_grads_only`: the data-gradient
GEMMs and GroupNorm backward kernels only - no weight / bias / GroupNorm-parameter gradients), and the ODE state
[x, delta_logp] stays on the device inside `ode.solve_ivp_rk45` (scipy's RK45 algorithm; the reference round-trips
the state through numpy for each of the ~10^2-10^3 function evaluations).  `solver='scipy'` keeps the reference's host
loop.  Random draws can be injected for parity tests (`injected=dict(...)`).
"""
import numpy as np
import torch
from scipy import integrate

from . import ode, ops
from .models import utils as mutils


def get_div_fn(fn):
  """div_fn(x, t, eps) ~ tr(d fn / dx) estimated as eps^T (d fn/dx) eps (reference likelihood.py:27-37)."""

  def div_fn(x, t, eps):
    with torch.enable_grad():
      xg = x.detach().requires_grad_(True)
      with ops.input_grads_only():
        proj = torch.sum(fn(xg, t) * eps)
        vjp = torch.autograd.grad(proj, xg)[0]
    return torch.sum(vjp * eps, dim=tuple(range(1, x.dim())))

  return div_fn


def _hutchinson_noise(like, kind, injected, key='epsilon'):
  if injected is not None and key in injected:
    return injected[key].to(like.device).float()
  if kind == 'Gaussian':
    return torch.randn_like(like)
  if kind == 'Rademacher':
    return torch.randint_like(like, low=0, high=2).float() * 2 - 1.
  raise NotImplementedError(f"Hutchinson type {kind} unknown.")


def _draw(like, injected, key):
  return injected[key].to(like.device).float() if injected is not None and key in injected else torch.randn_like(like)


def get_likelihood_fn(config, sde, inverse_scaler, hutchinson_type='Rademacher', rtol=1e-5, atol=1e-5, method='RK45',
                      solver='device'):
  """`solver`: 'device' (state resident on the GPU, `ode.solve_ivp_rk45`) or 'scipy' (the reference's host loop)."""
  if solver == 'device' and method != 'RK45':
    raise NotImplementedError("the device solver implements RK45; pass solver='scipy' for other methods")

  def drift_fn(model, x, t):
    score_fn = mutils.get_score_fn(config, sde, model, train=False, continuous=True)
    rsde = sde.reverse(score_fn, probability_flow=config.eval.probability_flow, lambda_=config.eval.lambda_)
    return rsde.sde(x, t)[0]

  def drift_and_div(model, x, t, noise):
    """One network forward + one input-gradient backward give both the drift and its divergence estimate."""
    with torch.enable_grad():
      xg = x.detach().requires_grad_(True)
      with ops.input_grads_only():
        drift = drift_fn(model, xg, t)
        vjp = torch.autograd.grad(torch.sum(drift * noise), xg)[0]
    return drift.detach(), torch.sum(vjp * noise, dim=(1, 2, 3))

  def likelihood_fn(model, data, logdet=0., eps=1e-5, mode='correct', injected=None):
    with torch.no_grad():
      score_fn = mutils.get_score_fn(config, sde, model, train=False, continuous=True)
      shape = data.shape
      B, dev = shape[0], data.device
      n_dim = int(np.prod(shape[1:]))
      epsilon = _hutchinson_noise(data, hutchinson_type, injected)
      if mode == 'correct':
        z = _draw(data, injected, 'z')
        mean, std = sde.marginal_prob(data, torch.ones(B, device=dev) * eps)
        start = mean + std[:, None, None, None] * z
      elif mode == 'wrong':
        start = data
      else:
        raise NotImplementedError

      if solver == 'device':
        def rhs(t, state):
          x = state[:B * n_dim].reshape(shape).float()
          drift, div = drift_and_div(model, x, torch.ones(B, device=dev) * t, epsilon)
          return torch.cat([drift.reshape(-1), div.reshape(-1)])

        init = torch.cat([start.reshape(-1).double(), torch.zeros(B, dtype=torch.float64, device=dev)])
        sol = ode.solve_ivp_rk45(rhs, (eps, sde.T), init, rtol=rtol, atol=atol)
        if not sol.success:
          raise RuntimeError('likelihood ODE: step size underflow')
        nfe, zp = sol.nfev, sol.y
        z = zp[:B * n_dim].reshape(shape).float()
        delta_logp = zp[B * n_dim:].float()
      else:
        def ode_func(t, x):
          sample = mutils.from_flattened_numpy(x[:-B], shape).to(dev).type(torch.float32)
          drift, div = drift_and_div(model, sample, torch.ones(B, device=dev) * t, epsilon)
          return np.concatenate([mutils.to_flattened_numpy(drift), mutils.to_flattened_numpy(div)], axis=0)

        init = np.concatenate([mutils.to_flattened_numpy(start), np.zeros((B,))], axis=0)
        solution = integrate.solve_ivp(ode_func, (eps, sde.T), init, rtol=rtol, atol=atol, method=method)
        nfe, zp = solution.nfev, solution.y[:, -1]
        z = mutils.from_flattened_numpy(zp[:-B], shape).to(dev).type(torch.float32)
        delta_logp = mutils.from_flattened_numpy(zp[-B:], (B,)).to(dev).type(torch.float32)

      prior_logp = sde.prior_logp(z)
      if mode == 'correct':
        residual_fn = get_likelihood_residual_fn(config, sde, score_fn, variance='scoreflow')
        delta_logp = delta_logp - residual_fn(data, eps, injected=injected)
      bpd = -(prior_logp + delta_logp + logdet) / np.log(2) / n_dim
      # log-likelihood -> bits/dim of 8-bit data (the reference's offset, likelihood.py:128-130)
      bpd = bpd + (7. - inverse_scaler(-1.))
      return bpd, z, nfe

  return likelihood_fn


def get_elbo_fn(config, sde, inverse_scaler=None, hutchinson_type='Rademacher'):
  """Single-sample NELBO estimate in bits/dim (reference likelihood.py:136-208)."""
  rve = config.training.sde.lower() == 'reciprocal_vesde'

  @torch.enable_grad()
  def loss_fn(model, batch, logdet=0., eps=1e-5, injected=None):
    score_fn = mutils.get_score_fn(config, sde, model, train=False, continuous=True)
    B, dev = batch.shape[0], batch.device
    n_dim = int(np.prod(batch.shape[1:]))
    if injected is not None and 'u' in injected:
      time, Z = sde.time_from_uniform(injected['u'].to(dev), eps, True)
    else:
      time, Z = sde.get_diffusion_time(config, B, dev, eps, importance_sampling=True)
    qt = 1. / (1. / eps - 1. / sde.T) if rve else 1. / (sde.T - eps)
    z = _draw(batch, injected, 'z')
    mean, std = sde.marginal_prob(batch, time)
    s4 = std[:, None, None, None]
    xt = (mean + s4 * z).detach().requires_grad_(True)
    epsilon = _hutchinson_noise(batch, hutchinson_type, injected)
    with ops.input_grads_only():
      score = score_fn(xt, time)
      f, g = sde.sde(xt, time)
      a = s4 * score
      mu = s4 ** 2 * score - s4 ** 2 / g[:, None, None, None] ** 2 * f
      vjp = torch.autograd.grad(mu, xt, epsilon, create_graph=False)[0]
    Mu = -(vjp * epsilon).reshape(B, -1).sum(1) * Z / qt
    Nu = -(a.detach() ** 2).reshape(B, -1).sum(1) * Z / 2 / qt
    lp_z = _draw(batch, injected, 'lp_z')
    lp_mean, lp_std = sde.marginal_prob(batch, torch.ones_like(time) * sde.T)
    lp = sde.prior_logp(lp_mean + lp_std[:, None, None, None] * lp_z)
    weight = 2. * eps * np.log(sde.sigma_max / sde.sigma_min) if rve else 1.
    elbos = lp + (Mu + Nu) * weight
    with torch.no_grad():
      residual_fn = get_likelihood_residual_fn(config, sde, score_fn, variance='scoreflow')
      residual = residual_fn(batch, eps, injected=injected)
    nelbo = -(elbos + logdet) / n_dim / np.log(2) + 7. - inverse_scaler(-1.)
    return nelbo.detach(), residual / n_dim / np.log(2)

  return loss_fn


def get_likelihood_residual_fn(config, sde, score_fn, variance='ddpm'):
  """Reconstruction term of the truncated likelihood bound at time eps: the Tweedie denoiser q(x_0 | x_eps) against
  the perturbation entropy (reference likelihood.py:210-313); discretised-Gaussian decoder for lossless data."""

  def _std_normal_cdf(x):
    return 0.5 * (1.0 + torch.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * (x ** 3))))

  def _discretized_gaussian_logp(x, means, log_scales):
    # data are integers in [0, 255] rescaled to [-1, 1]: bin half-width 1/255
    assert x.shape == means.shape
    floor = torch.tensor(1e-12, device=x.device)
    d = x - means
    inv = torch.exp(-log_scales)
    cdf_hi = _std_normal_cdf(inv * (d + 1. / 255.))
    cdf_lo = _std_normal_cdf(inv * (d - 1. / 255.))
    inner = torch.log(torch.max(cdf_hi - cdf_lo, floor))
    upper = torch.log(torch.max(1. - cdf_lo, floor))
    lower = torch.log(torch.max(cdf_hi, floor))
    return torch.where(x < -0.999, lower, torch.where(x > 0.999, upper, inner))

  def _posterior(batch, eps, injected):
    eps = sde.eps if eps is None else eps
    B = batch.shape[0]
    eps_vec = torch.ones(B, device=batch.device) * eps
    mean, std = sde.marginal_prob(batch, eps_vec)
    z = _draw(batch, injected, 'z_res')
    xe = mean + std[:, None, None, None] * z
    score = score_fn(xe, eps_vec)
    alpha, beta = sde.marginal_prob(torch.ones_like(batch), eps_vec)
    q_mean = xe / alpha + beta[:, None, None, None] ** 2 * score / alpha
    if variance == 'ddpm':
      q_std = beta
    elif variance == 'scoreflow':
      q_std = beta / torch.mean(alpha, axis=(1, 2, 3))
    else:
      raise ValueError(f'variance {variance!r} unknown')
    n_dim = float(np.prod(batch.shape[1:]))
    p_entropy = n_dim / 2. * (np.log(2 * np.pi) + 2 * torch.log(std) + 1.)
    return q_mean, q_std, p_entropy, n_dim

  def residual_lossless(batch, eps=None, injected=None):
    q_mean, q_std, p_entropy, _ = _posterior(batch, eps, injected)
    if not config.data.centered:
      batch, q_mean, q_std = 2. * batch - 1., 2. * q_mean - 1., 2. * q_std
    nll = -_discretized_gaussian_logp(batch, q_mean, torch.log(q_std)[:, None, None, None].expand_as(batch))
    return nll.sum(axis=(1, 2, 3)) - p_entropy

  def residual_gaussian(batch, eps=None, injected=None):
    q_mean, q_std, p_entropy, n_dim = _posterior(batch, eps, injected)
    q_recon = n_dim / 2. * (np.log(2 * np.pi) + 2 * torch.log(q_std)) + \
        0.5 / (q_std ** 2) * torch.square(batch - q_mean).sum(axis=(1, 2, 3))
    return q_recon - p_entropy

  return residual_lossless if config.data.dequantization == 'lossless' else residual_gaussian
