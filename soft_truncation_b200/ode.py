"""Adaptive Dormand-Prince RK45 integrator whose state stays on the device.

The reference integrates the probability-flow ODE with `scipy.integrate.solve_ivp(method='RK45')`
(sampling.py:492, likelihood.py:111): every function evaluation moves the whole state device -> numpy -> device
(models/utils.py:193-200).  This solver follows the same published algorithm step for step - Dormand-Prince 5(4)
tableau, RMS error norm against `atol + rtol*max(|y|,|y_new|)`, safety 0.9, step factors clamped to [0.2, 10],
Hairer's initial step - with the state as ONE fp64 torch tensor on the device; only the scalar error norm of each
step is read back.  `tests/test_host.py` checks it against scipy on the same right-hand side (same accepted steps,
same nfev).
"""
import math
from types import SimpleNamespace

import torch

_C = (0., 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.)
_A = ((),
      (1 / 5,),
      (3 / 40, 9 / 40),
      (44 / 45, -56 / 15, 32 / 9),
      (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
      (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656))
_B = (35 / 384, 0., 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84)
_E = (-71 / 57600, 0., 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40)
_SAFETY, _MIN_FACTOR, _MAX_FACTOR = 0.9, 0.2, 10.
_ERR_EXP = -1. / 5.          # -1 / (error estimator order + 1)


def _rms(v):
  return (torch.linalg.vector_norm(v) / math.sqrt(v.numel())).item()


def _initial_step(fun, t0, y0, t1, f0, direction, rtol, atol):
  span = abs(t1 - t0)
  if span == 0.:
    return 0., 0
  scale = atol + y0.abs() * rtol
  d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
  h0 = 1e-6 if d0 < 1e-5 or d1 < 1e-5 else 0.01 * d0 / d1
  h0 = min(h0, span)
  f1 = fun(t0 + h0 * direction, y0 + h0 * direction * f0)
  d2 = _rms((f1 - f0) / scale) / h0
  h1 = max(1e-6, h0 * 1e-3) if d1 <= 1e-15 and d2 <= 1e-15 else (0.01 / max(d1, d2)) ** (1. / 5.)
  return min(100 * h0, h1, span), 1


def solve_ivp_rk45(fun, t_span, y0, rtol=1e-3, atol=1e-6, max_steps=100000):
  """Integrate dy/dt = fun(t, y) from t_span[0] to t_span[1].

  fun(t: float, y: fp64 tensor) -> tensor of y's shape (any float dtype; promoted to fp64).  Returns a namespace with
  `y` (final state, fp64, on y0's device), `t`, `nfev`, `n_accepted`, `n_rejected`, `success`, `ts` (accepted times)."""
  t0, t1 = float(t_span[0]), float(t_span[1])
  y = y0.detach().to(torch.float64).clone()
  f = lambda t, v: fun(t, v).to(torch.float64)      # noqa: E731
  direction = 1. if t1 >= t0 else -1.
  t = t0
  k0 = f(t, y)
  nfev = 1
  h_abs, n = _initial_step(f, t, y, t1, k0, direction, rtol, atol)
  nfev += n
  accepted = rejected = 0
  ts = [t]
  K = [k0, None, None, None, None, None, None]
  while direction * (t - t1) < 0:
    if accepted + rejected >= max_steps:
      return SimpleNamespace(y=y, t=t, nfev=nfev, n_accepted=accepted, n_rejected=rejected, success=False, ts=ts)
    min_step = 10 * abs(math.nextafter(t, direction * math.inf) - t)
    h_abs = max(h_abs, min_step)
    step_rejected = False
    while True:
      if h_abs < min_step:
        return SimpleNamespace(y=y, t=t, nfev=nfev, n_accepted=accepted, n_rejected=rejected, success=False, ts=ts)
      h = h_abs * direction
      t_new = t + h
      if direction * (t_new - t1) > 0:
        t_new = t1
      h = t_new - t
      h_abs = abs(h)
      for s in range(1, 6):
        dy = K[0] * _A[s][0]
        for j in range(1, s):
          dy = dy + K[j] * _A[s][j]
        K[s] = f(t + _C[s] * h, y + dy * h)
      upd = K[0] * _B[0]
      for j in range(2, 6):                       # B[1] == 0
        upd = upd + K[j] * _B[j]
      y_new = y + upd * h
      K[6] = f(t + h, y_new)
      nfev += 6
      err = K[0] * _E[0]
      for j in range(2, 7):                       # E[1] == 0
        err = err + K[j] * _E[j]
      scale = atol + torch.maximum(y.abs(), y_new.abs()) * rtol
      err_norm = _rms(err * h / scale)
      if err_norm < 1.:
        factor = _MAX_FACTOR if err_norm == 0. else min(_MAX_FACTOR, _SAFETY * err_norm ** _ERR_EXP)
        if step_rejected:
          factor = min(1., factor)
        h_abs *= factor
        accepted += 1
        break
      h_abs *= max(_MIN_FACTOR, _SAFETY * err_norm ** _ERR_EXP)
      step_rejected = True
      rejected += 1
    t, y = t_new, y_new
    K[0] = K[6]
    ts.append(t)
  return SimpleNamespace(y=y, t=t, nfev=nfev, n_accepted=accepted, n_rejected=rejected, success=True, ts=ts)
