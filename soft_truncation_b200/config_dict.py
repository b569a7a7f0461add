"""Minimal attribute-style config tree.

The reference reads its hyper-parameters from `ml_collections.ConfigDict` objects
(reference: configs/default_cifar10_configs.py:5-8, main.py:44).  `ml_collections`
is not installed here, and the hot path only ever needs attribute get/set on a
nested tree plus `.to_dict()`, so this is a small stand-in with that surface.
"""


class ConfigDict:
  """Nested attribute dictionary (`cfg.model.nf`, `cfg.model.nf = 128`)."""

  def __init__(self, initial=None):
    object.__setattr__(self, '_fields', {})
    if initial:
      for k, v in dict(initial).items():
        setattr(self, k, v)

  def __getattr__(self, name):
    fields = object.__getattribute__(self, '_fields')
    if name in fields:
      return fields[name]
    raise AttributeError(name)

  def __setattr__(self, name, value):
    if isinstance(value, dict):
      value = ConfigDict(value)
    self._fields[name] = value

  __getitem__ = __getattr__
  __setitem__ = __setattr__

  def __contains__(self, name):
    return name in self._fields

  def get(self, name, default=None):
    return self._fields.get(name, default)

  def keys(self):
    return self._fields.keys()

  def items(self):
    return self._fields.items()

  def to_dict(self):
    out = {}
    for k, v in self._fields.items():
      out[k] = v.to_dict() if isinstance(v, ConfigDict) else v
    return out

  def copy(self):
    return ConfigDict(self.to_dict())

  def __repr__(self):
    return f'ConfigDict({self.to_dict()!r})'
