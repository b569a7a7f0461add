"""Data normalisers and on-GPU batch preparation - SURVEY 8(f)4.

The reference prepares every training batch on the host in float32 through a TensorFlow input pipeline
(datasets.py:106-128,305-330: convert to [0,1], random left-right flip), moves 4 bytes per value to the GPU, and then
dequantises and rescales it with three more elementwise passes (run_lib.py:73-75).  Here a batch travels as uint8
(1 byte per value, from pinned host memory) and ONE kernel (`st_prep_batch`) produces the fp32 NCHW network input:
/255, flip, uniform dequantisation, scaler.  `get_data_scaler` / `get_data_inverse_scaler` keep the reference's
signatures (datasets.py:56-71).  The TensorFlow / tfds loaders themselves are out of scope; `U8Loader`
is the minimal non-TF iterator over an in-memory uint8 image array that `get_batch` (datasets.py:106-113) needs.
"""
import numpy as np
import torch

from . import ops
from ._lib import check, lib


def get_data_scaler(config):
  """Data normalizer; data are assumed to be in [0, 1] (reference datasets.py:56-62)."""
  if config.data.centered:
    return lambda x: x * 2. - 1.
  return lambda x: x


def get_data_inverse_scaler(config):
  """Inverse data normalizer (reference datasets.py:65-71)."""
  if config.data.centered:
    return lambda x: (x + 1.) / 2.
  return lambda x: x


def prepare_batch(config, batch_u8, train=True, injected=None, seed=None, out=None):
  """uint8 images (B, H, W, C) - a CUDA tensor, or a (pinned) host tensor that is copied asynchronously - to the scaled
  fp32 (B, C, H, W) batch `step_fn` takes.

  Applies, in the reference's order: /255; a left-right flip of each image with probability 1/2 when
  `data.random_flip` and `train` (datasets.py:311,322); `(255 x + u)/256` when `data.dequantization == 'uniform'`
  (run_lib.py:73-74); the data scaler (run_lib.py:75).  `injected` = dict(flip=(B,) bool, u=(B, C, H, W) fp32)
  replaces the draws; otherwise the flips are drawn with torch.rand on the device and u by the kernel's counter-based
  generator (`seed`, default: a fresh torch draw)."""
  if batch_u8.dtype != torch.uint8 or batch_u8.dim() != 4:
    raise ValueError(f'prepare_batch expects a uint8 (B, H, W, C) tensor, got {batch_u8.dtype} {tuple(batch_u8.shape)}')
  dev = config.device if not batch_u8.is_cuda else batch_u8.device
  if torch.device(dev).type != 'cuda':
    raise RuntimeError('prepare_batch runs on the GPU (no CPU fallback)')
  x = batch_u8.to(dev, non_blocking=True).contiguous()
  B, H, W, C = x.shape
  if C != config.data.num_channels or H != config.data.image_size or W != config.data.image_size:
    raise ValueError(f'batch shape {tuple(x.shape)} does not match the config ({config.data.image_size}x'
                     f'{config.data.image_size}x{config.data.num_channels})')
  flip = None
  if config.data.random_flip and train:
    if injected is not None and 'flip' in injected:
      flip = injected['flip'].to(dev).to(torch.uint8).contiguous()
    else:
      flip = (torch.rand(B, device=dev) < 0.5).to(torch.uint8)
  dequant = config.data.dequantization == 'uniform'
  u = None
  if dequant and injected is not None and 'u' in injected:
    u = injected['u'].to(dev).float().contiguous()
  if seed is None:
    seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if dequant and u is None else 0
  a, b = (2., -1.) if config.data.centered else (1., 0.)
  if out is None:
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
  check(lib.st_prep_batch(ops.ptr(x), ops.ptr(u), ops.ptr(flip), ops.ptr(out), B, C, H, W, int(dequant), int(seed),
                          a, b, ops.stream()))
  return out


class U8Loader:
  """Epoch iterator over an in-memory uint8 image array (N, H, W, C): shuffled each epoch (NumPy generator), drops the
  last partial batch, yields PINNED uint8 host tensors so that the copy in `prepare_batch` is asynchronous."""

  def __init__(self, images, batch_size, shuffle=True, seed=0):
    images = np.ascontiguousarray(images)
    if images.dtype != np.uint8 or images.ndim != 4:
      raise ValueError('U8Loader expects a uint8 (N, H, W, C) array')
    if batch_size <= 0 or batch_size > images.shape[0]:
      raise ValueError(f'batch size {batch_size} does not fit {images.shape[0]} images')
    self.images, self.batch_size, self.shuffle = torch.from_numpy(images), batch_size, shuffle
    self.rng = np.random.default_rng(seed)
    self._pin = torch.cuda.is_available()

  def __len__(self):
    return self.images.shape[0] // self.batch_size

  def __iter__(self):
    n = self.images.shape[0]
    order = self.rng.permutation(n) if self.shuffle else np.arange(n)
    for i in range(len(self)):
      idx = torch.from_numpy(np.sort(order[i * self.batch_size:(i + 1) * self.batch_size]) if not self.shuffle
                             else order[i * self.batch_size:(i + 1) * self.batch_size])
      out = torch.empty((self.batch_size,) + tuple(self.images.shape[1:]), dtype=torch.uint8, pin_memory=self._pin)
      torch.index_select(self.images, 0, idx, out=out)
      yield out


def get_batch(config, data_iter, data):
  """Next batch, restarting the epoch when the iterator is exhausted (reference datasets.py:106-113).
  Returns (uint8 batch, iterator)."""
  try:
    batch = next(data_iter)
  except StopIteration:
    data_iter = iter(data)
    batch = next(data_iter)
  return batch, data_iter
