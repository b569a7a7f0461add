"""Flat parameter storage for the score network.

All parameters live in ONE fp32 buffer (`flat`), their gradients in a second one (`grad`) and a
compute-dtype copy (bf16 in fast mode) in a third (`comp`).  Each `nn.Parameter` the model exposes is
a strided VIEW into `flat` whose logical shape is the reference's (OIHW convolution weights,
(in, out) NIN matrices; reference models/layers.py:100-124,546-555) while the bytes underneath are in
the order the kernels want (K-major `[Cout][kh][kw][Cin]`, `[out][in]`).  This gives

  * checkpoint interchange with the reference's `state_dict()` (same names and shapes),
  * one fused clip+Adam+EMA launch over the whole model (losses.py), and
  * weight-gradient GEMMs that accumulate straight into `grad` with no layout change.
"""
import math

import numpy as np
import torch

ALIGN = 64   # elements; keeps every segment 128-byte aligned in bf16 (TMA needs 16)


class Entry:
  """One parameter: logical (reference) shape + physical (padded, kernel-order) extent.

  `pad` = (cout_phys, cin_phys) for 'conv' entries / (n_phys,) for 'vec' entries: 3-channel image
  convolutions are stored zero-padded to 64 channels so that they run on the same 64-wide tensor-core
  K/N blocks as every other convolution; the logical view simply skips the padding."""
  __slots__ = ('name', 'shape', 'kind', 'numel', 'offset', 'region', 'trainable', 'init', 'pad', 'pack')

  def __init__(self, name, shape, kind, region, trainable, init, pad=None, pack=None):
    self.name, self.shape, self.kind, self.region = name, tuple(shape), kind, region
    self.trainable, self.init, self.pack = trainable, init, pack
    if kind == 'conv':
      co, ci, kh, kw = self.shape
      self.pad = tuple(pad) if pad else (co, ci)
      self.numel = self.pad[0] * kh * kw * self.pad[1]
    elif kind == 'vec' and pad:
      self.pad = tuple(pad)
      self.numel = int(self.pad[0])
    else:
      self.pad = None
      self.numel = int(np.prod(shape))
    self.offset = -1


def _phys_view(buf, e):
  """Physical (kernel-order) view of entry `e` inside flat buffer `buf`."""
  seg = buf[e.offset:e.offset + e.numel]
  if e.kind == 'conv':
    co, ci, kh, kw = e.shape
    return seg.view(e.pad[0], kh * kw * e.pad[1])
  if e.kind == 'nin':
    ci, co = e.shape
    return seg.view(co, ci)
  if e.kind == 'vec':
    return seg
  return seg.view(e.shape)


def _logical_view(buf, e):
  """View with the reference's logical shape (what state_dict()/optimizers see)."""
  seg = buf[e.offset:e.offset + e.numel]
  if e.kind == 'conv':
    co, ci, kh, kw = e.shape
    return seg.view(e.pad[0], kh, kw, e.pad[1])[:co, :, :, :ci].permute(0, 3, 1, 2)
  if e.kind == 'nin':
    ci, co = e.shape
    return seg.view(co, ci).t()
  if e.kind == 'vec':
    return seg[:int(np.prod(e.shape))].view(e.shape)
  return seg.view(e.shape)


class ParamStore:
  REGIONS = ('dense_w', 'dense_b', 'conv0_b', 'main')

  def __init__(self):
    self.entries = []
    self.by_name = {}
    self.total = 0

  def add(self, name, shape, kind='vec', region='main', trainable=True, init=None, pad=None, pack=None):
    e = Entry(name, shape, kind, region, trainable, init, pad, pack)
    self.entries.append(e)
    self.by_name[name] = e
    return e

  def layout(self):
    """Assign offsets.  `entries` keeps the reference's registration order (that is the state_dict
    order); the byte order differs: the dense_w / dense_b regions and entries sharing a `pack` key are
    placed back to back because the kernels read them as ONE matrix / vector."""
    off = 0
    for region in self.REGIONS:
      placed = set()
      for e in self.entries:
        if e.region != region or id(e) in placed:
          continue
        group = [e] if e.pack is None else [x for x in self.entries if x.pack == e.pack]
        for x in group:
          x.offset = off
          off += x.numel
          placed.add(id(x))
        if region == 'main':
          off = -(-off // ALIGN) * ALIGN
      off = -(-off // ALIGN) * ALIGN
    self.total = off

  def region_span(self, region):
    es = [e for e in self.entries if e.region == region]
    if not es:
      return 0, 0
    return es[0].offset, sum(e.numel for e in es)


# ------------------------------------------------------------------------------------ initialisers
def variance_scaling_uniform(shape, scale, in_axis, out_axis, gen):
  """fan_avg / uniform variance scaling (reference models/layers.py:54-85); scale 0 -> 1e-10 (:88-91)."""
  scale = 1e-10 if scale == 0 else scale
  rf = np.prod(shape) / shape[in_axis] / shape[out_axis]
  fan_in, fan_out = shape[in_axis] * rf, shape[out_axis] * rf
  variance = scale / ((fan_in + fan_out) / 2)
  return (torch.rand(*shape, generator=gen) * 2. - 1.) * math.sqrt(3 * variance)


def init_conv(scale):
  return lambda shape, gen: variance_scaling_uniform(shape, scale, 1, 0, gen)


def init_zeros(shape, gen):
  return torch.zeros(shape)


def init_ones(shape, gen):
  return torch.ones(shape)


# ------------------------------------------------------------------------------------ flat-buffer owners
_OWNERS = {}    # storage data_ptr -> weakref to the model that owns the flat buffer


def register_owner(model):
  import weakref
  _OWNERS[model._flat.untyped_storage().data_ptr()] = weakref.ref(model)


def flat_owner(parameters):
  """The NCSNpp whose flat buffer backs ALL of `parameters`, else None."""
  parameters = list(parameters)
  if not parameters:
    return None
  ref = _OWNERS.get(parameters[0].untyped_storage().data_ptr())
  model = ref() if ref is not None else None
  if model is None:
    return None
  ptr = model._flat.untyped_storage().data_ptr()
  own = [p for p in model.parameters()]
  if len(own) != len(parameters) or any(p.untyped_storage().data_ptr() != ptr for p in parameters):
    return None
  return model
