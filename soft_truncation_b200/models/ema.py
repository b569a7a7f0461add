"""Exponential moving average of the parameters (reference models/ema.py:15-97), same API and
state_dict format.  When the tracked parameters are the views of an NCSNpp flat buffer the shadow
copy is flat as well, so `update` is ONE kernel over 61.8 M elements instead of three per tensor
(and is usually fused away entirely into the optimizer step, losses.py)."""
import torch

from .. import ops
from . import params as _params


class ExponentialMovingAverage:

  def __init__(self, parameters, decay, use_num_updates=True):
    if decay < 0.0 or decay > 1.0:
      raise ValueError('Decay must be between 0 and 1')
    self.decay = decay
    self.num_updates = 0 if use_num_updates else None
    parameters = list(parameters)
    self.owner = _params.flat_owner(parameters)
    self.collected_params = []
    if self.owner is not None:
      m = self.owner
      self.shadow_flat = m._flat.clone().detach()
      self.shadow_params = [_params._logical_view(self.shadow_flat, e) for e in m.store.entries if e.trainable]
      self.mask = m.trainable_mask()
    else:
      self.shadow_flat = None
      self.shadow_params = [p.clone().detach() for p in parameters if p.requires_grad]

  def next_decay(self):
    """Advance the update counter and return the decay of this update (reference :43-46)."""
    decay = self.decay
    if self.num_updates is not None:
      self.num_updates += 1
      decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
    return decay

  def update(self, parameters):
    decay = self.next_decay()
    if self.owner is not None and self.owner._flat.is_cuda:
      # s <- s - (1-d)(s - p) == d*s + (1-d)*p, one pass over the flat buffer
      ops.axpby(self.shadow_flat, self.owner._flat, alpha=decay, beta=1.0 - decay, out=self.shadow_flat)
      return
    one_minus_decay = 1.0 - decay
    with torch.no_grad():
      parameters = [p for p in parameters if p.requires_grad]
      for s_param, param in zip(self.shadow_params, parameters):
        s_param.sub_(one_minus_decay * (s_param - param))

  def copy_to(self, parameters):
    parameters = [p for p in parameters if p.requires_grad]
    for s_param, param in zip(self.shadow_params, parameters):
      if param.requires_grad:
        param.data.copy_(s_param.data)

  def store(self, parameters):
    self.collected_params = [param.clone() for param in parameters]

  def restore(self, parameters):
    for c_param, param in zip(self.collected_params, parameters):
      param.data.copy_(c_param.data)

  def state_dict(self):
    return dict(decay=self.decay, num_updates=self.num_updates, shadow_params=self.shadow_params)

  def load_state_dict(self, state_dict):
    self.decay = state_dict['decay']
    self.num_updates = state_dict['num_updates']
    loaded = state_dict['shadow_params']
    if self.shadow_flat is not None:
      for dst, src in zip(self.shadow_params, loaded):
        dst.copy_(src)
    else:
      self.shadow_params = loaded
