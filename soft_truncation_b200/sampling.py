"""Reverse-SDE predictor-corrector sampling (reference sampling.py:30-125, 185-433) and the
probability-flow ODE sampler (:436-504) on the B200 kernels.

Same registry surface as the reference (`register_predictor/corrector`, `get_predictor/corrector`,
`get_sampling_fn(config, sde, shape, inverse_scaler, eps)` -> `sampling_fn(model) -> (x, nfe)`).
Each built-in update is: one score-network evaluation (hand-written CUDA), a handful of B-length
schedule scalars from sde_lib, and ONE fused elementwise kernel (st_pc_update) producing both x_mean
and the noised x.  The latent never leaves HBM and there is no host synchronisation inside the loop;
with `sampling.cuda_graph` (default on) one reverse step is captured once and replayed N times.
"""
import abc
import functools

import numpy as np
import torch
from scipy import integrate

from . import ode, ops, sde_lib
from ._lib import check, lib
from .models import utils as mutils
from .models.utils import from_flattened_numpy, get_score_fn, to_flattened_numpy

_CORRECTORS = {}
_PREDICTORS = {}


def register_predictor(cls=None, *, name=None):
  """A decorator for registering predictor classes."""

  def _register(cls):
    local_name = cls.__name__ if name is None else name
    if local_name in _PREDICTORS:
      raise ValueError(f'Already registered model with name: {local_name}')
    _PREDICTORS[local_name] = cls
    return cls

  return _register if cls is None else _register(cls)


def register_corrector(cls=None, *, name=None):
  """A decorator for registering corrector classes."""

  def _register(cls):
    local_name = cls.__name__ if name is None else name
    if local_name in _CORRECTORS:
      raise ValueError(f'Already registered model with name: {local_name}')
    _CORRECTORS[local_name] = cls
    return cls

  return _register if cls is None else _register(cls)


def get_predictor(name):
  return _PREDICTORS[name]


def get_corrector(name):
  return _CORRECTORS[name]


def get_sampling_fn(config, sde, shape, inverse_scaler, eps):
  """Create a sampling function (reference sampling.py:80-125)."""
  sampler_name = config.sampling.method
  if sampler_name.lower() == 'ode':
    return get_ode_sampler(config=config, sde=sde, shape=shape, inverse_scaler=inverse_scaler,
                           denoise=config.sampling.noise_removal, eps=eps, device=config.device)
  elif sampler_name.lower() == 'pc':
    predictor = get_predictor(config.sampling.predictor.lower())
    corrector = get_corrector(config.sampling.corrector.lower())
    return get_pc_sampler(config=config, sde=sde, shape=shape, predictor=predictor, corrector=corrector,
                          inverse_scaler=inverse_scaler, snr=config.sampling.snr,
                          n_steps=config.sampling.n_steps_each, probability_flow=config.sampling.probability_flow,
                          continuous=config.training.continuous, denoise=config.sampling.noise_removal, eps=eps,
                          device=config.device)
  raise ValueError(f"Sampler name {sampler_name} unknown.")


# ------------------------------------------------------------------------------------ fused update
def _f32(v, like):
  return v.reshape(-1).to(device=like.device, dtype=torch.float32).contiguous()


def fused_update(x, s, noise, ca, cb, cc):
  """x_mean = ca[n]*x + cb[n]*s;  x_new = x_mean + cc[n]*noise  -> (x_new, x_mean), one kernel."""
  x, s = x.contiguous(), s.contiguous()
  B, D = x.shape[0], x[0].numel()
  x_mean, x_new = torch.empty_like(x), torch.empty_like(x)
  check(lib.st_pc_update(ops.ptr(x), ops.ptr(s), ops.ptr(noise.contiguous()) if noise is not None else None,
                         ops.ptr(_f32(ca, x)), ops.ptr(_f32(cb, x)), ops.ptr(_f32(cc, x)), ops.ptr(x_mean),
                         ops.ptr(x_new), B, D, ops.stream()))
  return x_new, x_mean


def _draw(x, injected):
  """Per-step Gaussian noise: torch.randn_like as in the reference, or the next injected tensor."""
  if injected is not None:
    return injected.pop(0).to(x.device)
  return torch.randn_like(x)


class Predictor(abc.ABC):
  """The abstract class for a predictor algorithm (reference sampling.py:127-157)."""

  def __init__(self, sde, score_fn, probability_flow=False, logsnr_model=None):
    super().__init__()
    self.sde = sde
    lambda_ = 0. if probability_flow else 1.
    self.rsde = sde.reverse(score_fn, probability_flow, lambda_=lambda_)
    self.score_fn = score_fn
    self.noise_source = None

  @abc.abstractmethod
  def update_fn(self, x, t, next_t=None):
    pass


class Corrector(abc.ABC):
  """The abstract class for a corrector algorithm (reference sampling.py:160-182)."""

  def __init__(self, sde, score_fn, snr, n_steps):
    super().__init__()
    self.sde = sde
    self.score_fn = score_fn
    self.snr = snr
    self.n_steps = n_steps
    self.noise_source = None

  @abc.abstractmethod
  def update_fn(self, x, t):
    pass


def _unit(x):
  return torch.ones((x.shape[0], 1, 1, 1), device=x.device)


@register_predictor(name='euler_maruyama')
class EulerMaruyamaPredictor(Predictor):
  """x_mean = x + [f(x,t) - g^2 w score] dt, x = x_mean + lambda g sqrt(-dt) z, dt = -1/N
  (reference sampling.py:185-196).  Every SDE here has a drift linear in x, f = fa(t) x, so
  x_mean = (1 + fa dt) x + (-g^2 w dt) score."""

  def __init__(self, config, sde, score_fn, probability_flow=False):
    super().__init__(sde, score_fn, probability_flow)

  def update_fn(self, x, t):
    dt = -1. / self.rsde.N
    z = _draw(x, self.noise_source)
    fa, g = self.sde.sde(_unit(x), t)
    score = self.score_fn(x, t)
    ca = 1. + fa.reshape(-1) * dt
    cb = -(g ** 2) * self.rsde.weight * dt
    cc = self.rsde.lambda_ * g * np.sqrt(-dt)
    return fused_update(x, score, z, ca, cb, cc)


@register_predictor(name='reverse_diffusion')
class ReverseDiffusionPredictor(Predictor):
  """x_mean = x - [f - G^2 w score], x = x_mean + lambda G z with (f, G) = sde.discretize
  (reference sampling.py:199-210); f = fd(t) x for every SDE here."""

  def __init__(self, config, sde, score_fn, probability_flow=False, logsnr_model=None):
    super().__init__(sde, score_fn, probability_flow, logsnr_model)
    self.config = config

  def update_fn(self, x, t, next_t=None):
    fd, G = self.sde.discretize(_unit(x), t, next_t)
    z = _draw(x, self.noise_source)
    score = self.score_fn(x, t)
    ca = 1. - fd.reshape(-1)
    cb = (G ** 2) * self.rsde.weight
    cc = self.rsde.lambda_ * G
    return fused_update(x, score, z, ca, cb, cc)


@register_predictor(name='ancestral_sampling')
class AncestralSamplingPredictor(Predictor):
  """The ancestral sampling predictor for VE/VP SDEs (reference sampling.py:213-250)."""

  def __init__(self, config, sde, score_fn, probability_flow=False):
    super().__init__(sde, score_fn, probability_flow)
    if not isinstance(sde, sde_lib.VPSDE) and not isinstance(sde, sde_lib.VESDE):
      raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")
    assert not probability_flow, "Probability flow not supported by ancestral sampling"

  def update_fn(self, x, t):
    sde = self.sde
    timestep = (t * (sde.N - 1) / sde.T).long()
    noise = _draw(x, self.noise_source)
    score = self.score_fn(x, t)
    if isinstance(sde, sde_lib.VESDE):
      sig = sde.table('discrete_sigmas', t.device)
      sigma = sig[timestep]
      adjacent = torch.where(timestep == 0, torch.zeros_like(t), sig[timestep - 1])
      ca, cb = torch.ones_like(t), sigma ** 2 - adjacent ** 2
      cc = torch.sqrt((adjacent ** 2 * (sigma ** 2 - adjacent ** 2)) / (sigma ** 2))
    else:
      beta = sde.table('discrete_betas', t.device)[timestep]
      ca = 1. / torch.sqrt(1. - beta)
      cb = beta * ca
      cc = torch.sqrt(beta)
    return fused_update(x, score, noise, ca, cb, cc)


@register_predictor(name='none')
class NonePredictor(Predictor):
  """An empty predictor that does nothing."""

  def __init__(self, sde, score_fn, probability_flow=False):
    pass

  def update_fn(self, x, t):
    return x, x


def _langevin_alpha(sde, t):
  if isinstance(sde, (sde_lib.VPSDE, sde_lib.subVPSDE)):
    timestep = (t * (sde.N - 1) / sde.T).long()
    return sde.table('alphas', t.device)[timestep].float()
  return torch.ones_like(t)


@register_corrector(name='langevin')
class LangevinCorrector(Corrector):
  """Langevin MCMC corrector (reference sampling.py:263-292).  The batch means of the per-sample
  score / noise norms and the step size stay on the device (st_batch_norms, st_langevin_coeffs)."""

  def __init__(self, sde, score_fn, snr, n_steps):
    super().__init__(sde, score_fn, snr, n_steps)
    if not isinstance(sde, (sde_lib.VPSDE, sde_lib.VESDE, sde_lib.subVPSDE)):
      raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")

  def update_fn(self, x, t):
    alpha = _langevin_alpha(self.sde, t).contiguous()
    B, D = x.shape[0], x[0].numel()
    x_mean = x
    for _ in range(self.n_steps):
      grad = self.score_fn(x, t).contiguous()
      noise = _draw(x, self.noise_source).contiguous()
      norms = torch.empty(2, dtype=torch.float32, device=x.device)
      coef = torch.empty((3, B), dtype=torch.float32, device=x.device)
      check(lib.st_batch_norms(ops.ptr(grad), ops.ptr(noise), ops.ptr(norms), B, D, ops.stream()))
      check(lib.st_langevin_coeffs(ops.ptr(norms), ops.ptr(alpha), float(self.snr), ops.ptr(coef[0]), ops.ptr(coef[1]),
                                   ops.ptr(coef[2]), B, ops.stream()))
      x, x_mean = fused_update(x, grad, noise, coef[0], coef[1], coef[2])
    return x, x_mean


@register_corrector(name='ald')
class AnnealedLangevinDynamics(Corrector):
  """Annealed Langevin dynamics of NCSN/NCSNv2 (reference sampling.py:295-329)."""

  def __init__(self, sde, score_fn, snr, n_steps):
    super().__init__(sde, score_fn, snr, n_steps)
    if not isinstance(sde, (sde_lib.VPSDE, sde_lib.VESDE, sde_lib.subVPSDE)):
      raise NotImplementedError(f"SDE class {sde.__class__.__name__} not yet supported.")

  def update_fn(self, x, t):
    alpha = _langevin_alpha(self.sde, t)
    std = self.sde.marginal_prob(_unit(x), t)[1]
    x_mean = x
    for _ in range(self.n_steps):
      grad = self.score_fn(x, t)
      noise = _draw(x, self.noise_source)
      step_size = (self.snr * std) ** 2 * 2 * alpha
      x, x_mean = fused_update(x, grad, noise, torch.ones_like(step_size), step_size, torch.sqrt(step_size * 2))
    return x, x_mean


@register_corrector(name='none')
class NoneCorrector(Corrector):
  """An empty corrector that does nothing."""

  def __init__(self, sde, score_fn, snr, n_steps):
    pass

  def update_fn(self, x, t):
    return x, x


def shared_predictor_update_fn(x, t, sde, model, predictor, probability_flow, continuous, config, noise_source=None):
  """Configures a predictor and runs one update (reference sampling.py:343-351)."""
  score_fn = mutils.get_score_fn(config, sde, model, train=False, continuous=continuous)
  if predictor is None:
    predictor_obj = NonePredictor(sde, score_fn, probability_flow)
  else:
    predictor_obj = predictor(config, sde, score_fn, probability_flow)
  predictor_obj.noise_source = noise_source
  return predictor_obj.update_fn(x, t)


def shared_corrector_update_fn(x, t, sde, model, corrector, continuous, snr, n_steps, config, noise_source=None):
  """Configures a corrector and runs one update (reference sampling.py:354-362)."""
  score_fn = mutils.get_score_fn(config, sde, model, train=False, continuous=continuous)
  if corrector is None:
    corrector_obj = NoneCorrector(sde, score_fn, snr, n_steps)
  else:
    corrector_obj = corrector(sde, score_fn, snr, n_steps)
  corrector_obj.noise_source = noise_source
  return corrector_obj.update_fn(x, t)


def get_pc_sampler(config, sde, shape, predictor, corrector, inverse_scaler, snr, n_steps=1, probability_flow=False,
                   continuous=False, denoise=True, eps=1e-3, device='cuda'):
  """Create a Predictor-Corrector sampler (reference sampling.py:365-433).

  pc_sampler(model, x_init=None, noises=None, trace=None):
    `x_init` / `noises` (a list of per-update noise tensors, consumed in order) replace the random draws
    for parity tests; `trace`, if a list, receives x after every reverse step.
  """
  predictor_update_fn = functools.partial(shared_predictor_update_fn, sde=sde, predictor=predictor,
                                          probability_flow=probability_flow, continuous=continuous, config=config)
  corrector_update_fn = functools.partial(shared_corrector_update_fn, sde=sde, corrector=corrector,
                                          continuous=continuous, snr=snr, n_steps=n_steps, config=config)

  def denoise_update_fn(model, x):
    score_fn = get_score_fn(config, sde, model, train=False, continuous=True)
    predictor_obj = ReverseDiffusionPredictor(config, sde, score_fn, probability_flow=True)
    vec_eps = torch.ones(x.shape[0], device=x.device) * sde.eps
    _, x = predictor_obj.update_fn(x, vec_eps, torch.zeros_like(vec_eps))
    return x

  use_graph = bool(config.sampling.get('cuda_graph', True)) if hasattr(config.sampling, 'get') else True
  graph_cache = {}

  def pc_sampler(model, x_init=None, noises=None, trace=None):
    with torch.no_grad():
      x = (sde.prior_sampling(shape) if x_init is None else x_init).to(device)
      timesteps = torch.linspace(sde.T, eps, sde.N, device=device)
      noise_source = list(noises) if noises is not None else None
      B = shape[0]
      graphable = use_graph and noise_source is None and trace is None and x.is_cuda and sde.N > 4

      def one_step(x, vec_t):
        x, x_mean = corrector_update_fn(x, vec_t, model=model, noise_source=noise_source)
        x, x_mean = predictor_update_fn(x, vec_t, model=model, noise_source=noise_source)
        return x, x_mean

      if not graphable:
        x_mean = x
        for i in range(sde.N):
          vec_t = torch.ones(B, device=device) * timesteps[i]
          x, x_mean = one_step(x, vec_t)
          if trace is not None:
            trace.append(x.clone())
      else:
        # One reverse step captured into a CUDA graph and replayed: the step index lives on the device,
        # so all N steps are enqueued back to back with no host round trip.  The graph is kept for later calls
        # with the same network buffers (it holds raw pointers into them).
        net = mutils.unwrap(model)
        key = (id(net),) + tuple(int(t.data_ptr()) for t in net.buffers_for_graph_key())
        g = graph_cache.get('g')
        if g is None or g['key'] != key:
          step = torch.zeros(1, dtype=torch.long, device=device)
          sx, sx_mean = x.clone(), x.clone()
          sts = timesteps.clone()
          side = torch.cuda.Stream()
          side.wait_stream(torch.cuda.current_stream())
          rng = torch.cuda.get_rng_state(device)     # the warm-up draws must not shift the caller's noise stream
          with torch.cuda.stream(side):
            for _ in range(2):                       # warm-up outside capture (allocator, lazy inits)
              one_step(sx, torch.ones(B, device=device) * sts[0])
          torch.cuda.current_stream().wait_stream(side)
          torch.cuda.set_rng_state(rng, device)
          try:
            from . import _lib
            graph = torch.cuda.CUDAGraph()
            calls0 = _lib.launches
            with torch.cuda.graph(graph):
              vec_t = torch.ones(B, device=device) * sts.index_select(0, step)
              nx, nx_mean = one_step(sx, vec_t)
              sx.copy_(nx)
              sx_mean.copy_(nx_mean)
              step.add_(1)
            # `calls`: C-ABI calls one replay stands for (each enqueued >= 1 of the library's kernels at capture)
            g = dict(key=key, graph=graph, step=step, sx=sx, sx_mean=sx_mean, sts=sts, calls=_lib.launches - calls0)
          except Exception as ex:                    # a predictor / corrector that cannot be captured: eager loop
            import warnings
            warnings.warn(f'soft_truncation_b200: CUDA-graph capture of the reverse step failed ({ex!r}); sampling eagerly')
            torch.cuda.synchronize()
            g = dict(key=key, graph=None)
          graph_cache['g'] = g
        if g['graph'] is None:
          x_mean = x
          for i in range(sde.N):
            x, x_mean = one_step(x, torch.ones(B, device=device) * timesteps[i])
          x_mean = x = denoise_update_fn(model, x_mean if denoise else x)
          return inverse_scaler(x_mean if denoise else x), sde.N * (n_steps + 1)
        g['sx'].copy_(x)
        g['sts'].copy_(timesteps)
        g['step'].zero_()
        for _ in range(sde.N):
          g['graph'].replay()
        x, x_mean = g['sx'].clone(), g['sx_mean'].clone()

      x_mean = x = denoise_update_fn(model, x_mean if denoise else x)
      return inverse_scaler(x_mean if denoise else x), sde.N * (n_steps + 1)

  pc_sampler.graph_cache = graph_cache      # (bench.py reads the per-replay launch count from it)
  return pc_sampler


def get_ode_sampler(config, sde, shape, inverse_scaler, denoise=False, rtol=1e-5, atol=1e-5, method='RK45', eps=1e-3,
                    device='cuda', solver='device'):
  """Probability-flow ODE sampler (reference sampling.py:436-504).  `solver='device'` integrates with
  `ode.solve_ivp_rk45` - scipy's RK45 algorithm with the state resident on the GPU - instead of moving the state
  through numpy for every function evaluation; `solver='scipy'` is the reference's black-box host loop."""
  if solver == 'device' and method != 'RK45':
    solver = 'scipy'

  def denoise_update_fn(model, x):
    score_fn = get_score_fn(config, sde, model, train=False, continuous=True)
    predictor_obj = ReverseDiffusionPredictor(config, sde, score_fn, probability_flow=False)
    vec_eps = torch.ones(x.shape[0], device=x.device) * sde.eps
    _, x = predictor_obj.update_fn(x, vec_eps, torch.zeros_like(vec_eps))
    return x

  def drift_fn(model, x, t):
    score_fn = get_score_fn(config, sde, model, train=False, continuous=True)
    rsde = sde.reverse(score_fn, probability_flow=True, lambda_=0.)
    return rsde.sde(x, t)[0]

  def ode_sampler(model, x_init=None):
    with torch.no_grad():
      x = (sde.prior_sampling(shape) if x_init is None else x_init).to(device)

      if solver == 'device':
        def rhs(t, state):
          xs = state.reshape(shape).float()
          return drift_fn(model, xs, torch.ones(shape[0], device=xs.device) * t).reshape(-1)

        sol = ode.solve_ivp_rk45(rhs, (sde.T, eps), x.reshape(-1).double(), rtol=rtol, atol=atol)
        if not sol.success:
          raise RuntimeError('ODE sampler: step size underflow')
        nfe = sol.nfev
        x = sol.y.reshape(shape).float()
      else:
        def ode_func(t, x):
          x = from_flattened_numpy(x, shape).to(device).type(torch.float32)
          vec_t = torch.ones(shape[0], device=x.device) * t
          return to_flattened_numpy(drift_fn(model, x, vec_t))

        solution = integrate.solve_ivp(ode_func, (sde.T, eps), to_flattened_numpy(x), rtol=rtol, atol=atol, method=method)
        nfe = solution.nfev
        x = torch.tensor(solution.y[:, -1]).reshape(shape).to(device).type(torch.float32)
      if denoise:
        x = denoise_update_fn(model, x)
      return inverse_scaler(x), nfe

  return ode_sampler
