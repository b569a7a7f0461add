"""ORACLE (test infrastructure, never shipped or timed as the product).

Plain fp32 PyTorch-on-CPU restatement of the reference score network, written as one
functional walk over a reference-format state_dict (keys `all_modules.{i}.<leaf>`, NCHW,
OIHW) instead of a module tree:

* NCSNpp.forward                    - reference models/ncsnpp.py:258-432
* ResnetBlockBigGANpp               - reference models/layerspp.py:255-287
* AttnBlockpp                       - reference models/layerspp.py:88-104
* NIN                               - reference models/layers.py:546-555
* get_timestep_embedding            - reference models/layers.py:515-529
* GaussianFourierProjection         - reference models/layerspp.py:45-54
* Combine / Upsample / Downsample   - reference models/layerspp.py:57-72,107-176
* FIR up/down-sampling              - reference models/up_or_down_sampling.py:144-257
* parameter initialisation          - reference models/layers.py:54-91,547-549; ncsnpp.py:74-256

Parity pin: tests/golden/*.npz were produced by importing the untouched reference from
/root/reference in this container (tests/golden/make_golden.py); tests/test_oracle.py
checks this restatement against them (score, per-block activations, gradients).
Only the configurations BASELINE.json names are covered (resblock_type='biggan').
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.)


# ----------------------------------------------------------------------------- resampling
def fir_kernel_2d(taps, gain=1.):
  k = np.asarray(taps, dtype=np.float32)
  if k.ndim == 1:
    k = np.outer(k, k)
  k = k / np.sum(k)
  return torch.tensor(k * gain)


def upfirdn(x, k, up=1, down=1, pad=(0, 0)):
  """NCHW zero-insert upsample -> pad/crop -> convolve with k -> decimate."""
  n, c, h, w = x.shape
  y = x.reshape(n * c, 1, h, w)
  if up > 1:
    z = y.new_zeros(n * c, 1, h * up, w * up)
    z[:, :, ::up, ::up] = y
    y = z
  p0, p1 = pad
  y = F.pad(y, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
  y = y[:, :, max(-p0, 0):y.shape[2] - max(-p1, 0), max(-p0, 0):y.shape[3] - max(-p1, 0)]
  y = F.conv2d(y, torch.flip(k.to(y), [0, 1])[None, None])
  y = y[:, :, ::down, ::down]
  return y.reshape(n, c, y.shape[2], y.shape[3])


def fir_upsample(x, taps, factor=2):
  k = fir_kernel_2d(taps, gain=factor ** 2)
  p = k.shape[0] - factor
  return upfirdn(x, k, up=factor, pad=((p + 1) // 2 + factor - 1, p // 2))


def fir_downsample(x, taps, factor=2):
  k = fir_kernel_2d(taps)
  p = k.shape[0] - factor
  return upfirdn(x, k, down=factor, pad=((p + 1) // 2, p // 2))


def fir_conv_downsample(x, w, taps, factor=2):
  k = fir_kernel_2d(taps)
  p = (k.shape[0] - factor) + (w.shape[-1] - 1)
  return F.conv2d(upfirdn(x, k, pad=((p + 1) // 2, p // 2)), w, stride=factor)


def nearest_up(x):
  n, c, h, w = x.shape
  return x[:, :, :, None, :, None].expand(n, c, h, 2, w, 2).reshape(n, c, 2 * h, 2 * w)


def mean_down(x):
  n, c, h, w = x.shape
  return x.reshape(n, c, h // 2, 2, w // 2, 2).mean(dim=(3, 5))


# ----------------------------------------------------------------------------- small pieces
def swish(x):
  return x * torch.sigmoid(x)


def group_norm(x, w, b):
  c = x.shape[1]
  return F.group_norm(x, min(c // 4, 32), w, b, eps=1e-6)


def positional_embedding(labels, dim, max_positions=10000):
  half = dim // 2
  freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(max_positions) / (half - 1)))
  arg = labels.float()[:, None] * freq[None, :]
  return torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)


def fourier_embedding(log_sigma, W):
  proj = log_sigma[:, None] * W[None, :] * 2 * np.pi
  return torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1)


def nin(x, W, b):
  return torch.einsum('bchw,co->bohw', x, W) + b[None, :, None, None]


class _Params:
  """Cursor over `all_modules.{i}.*`."""

  def __init__(self, sd):
    self.sd = {k[len('module.'):] if k.startswith('module.') else k: v for k, v in sd.items()}
    self.i = 0

  def take(self):
    pre = f'all_modules.{self.i}.'
    self.i += 1
    return {k[len(pre):]: v for k, v in self.sd.items() if k.startswith(pre)}

  def count(self):
    return 1 + max(int(k.split('.')[1]) for k in self.sd if k.startswith('all_modules.'))


def res_block(p, x, temb, cfg, up=False, down=False, train=False, drop_mask=None, taps=None):
  fir = cfg.model.fir
  h = swish(group_norm(x, p['GroupNorm_0.weight'], p['GroupNorm_0.bias']))
  if up:
    h, x = (fir_upsample(h, taps), fir_upsample(x, taps)) if fir else (nearest_up(h), nearest_up(x))
  elif down:
    h, x = (fir_downsample(h, taps), fir_downsample(x, taps)) if fir else (mean_down(h), mean_down(x))
  h = F.conv2d(h, p['Conv_0.weight'], p['Conv_0.bias'], padding=1)
  if temb is not None:
    h = h + F.linear(swish(temb), p['Dense_0.weight'], p['Dense_0.bias'])[:, :, None, None]
  h = swish(group_norm(h, p['GroupNorm_1.weight'], p['GroupNorm_1.bias']))
  if train and cfg.model.dropout > 0:
    if drop_mask is not None:
      h = h * drop_mask
    else:
      h = F.dropout(h, cfg.model.dropout, True)
  h = F.conv2d(h, p['Conv_1.weight'], p['Conv_1.bias'], padding=1)
  if 'Conv_2.weight' in p:
    x = F.conv2d(x, p['Conv_2.weight'], p['Conv_2.bias'])
  return (x + h) / SQRT2 if cfg.model.skip_rescale else x + h


def attn_block(p, x, cfg):
  b, c, hh, ww = x.shape
  h = group_norm(x, p['GroupNorm_0.weight'], p['GroupNorm_0.bias'])
  q = nin(h, p['NIN_0.W'], p['NIN_0.b']).reshape(b, c, hh * ww)
  k = nin(h, p['NIN_1.W'], p['NIN_1.b']).reshape(b, c, hh * ww)
  v = nin(h, p['NIN_2.W'], p['NIN_2.b']).reshape(b, c, hh * ww)
  w = torch.einsum('bcq,bck->bqk', q, k) * (int(c) ** (-0.5))
  w = F.softmax(w, dim=-1)
  h = torch.einsum('bqk,bck->bcq', w, v).reshape(b, c, hh, ww)
  h = nin(h, p['NIN_3.W'], p['NIN_3.b'])
  return (x + h) / SQRT2 if cfg.model.skip_rescale else x + h


def unet_forward(sd, cfg, x, time_cond, train=False, drop_masks=None, taps_out=None):
  """Score-network output for NCHW `x` and per-sample conditioning `time_cond`.

  `drop_masks`: optional {module_index: NCHW keep-mask already scaled by 1/(1-p)} to inject
  dropout; `taps_out`: optional dict that receives the output of every module by index.
  """
  m = cfg.model
  P = _Params(sd)
  taps = m.fir_kernel
  n_res = len(m.ch_mult)
  aux = m.auxiliary_resblock
  assert m.resblock_type.lower() == 'biggan'

  def tap(idx, val):
    if taps_out is not None:
      taps_out[idx] = val.detach()
    return val

  def mask_for(idx):
    return None if drop_masks is None else drop_masks.get(idx)

  if m.embedding_type.lower() == 'fourier':
    used_sigmas = time_cond
    temb = fourier_embedding(torch.log(used_sigmas), P.take()['W'])
  else:
    used_sigmas = sd.get('sigmas', sd.get('module.sigmas'))
    used_sigmas = used_sigmas[time_cond.long()] if used_sigmas is not None else None
    # model.lsgm: `embedding_dim`-wide sinusoidal embedding (reference models/ncsnpp.py:279-283)
    temb = positional_embedding(time_cond, m.embedding_dim if getattr(m, 'lsgm', False) else m.nf)
  if m.conditional:
    p = P.take()
    temb = F.linear(temb, p['weight'], p['bias'])
    p = P.take()
    temb = F.linear(swish(temb), p['weight'], p['bias'])
  else:
    temb = None

  if not cfg.data.centered:
    x = 2 * x - 1.
  pyr_in = x if m.progressive_input != 'none' else None

  p = P.take()
  hs = [tap(P.i - 1, F.conv2d(x, p['weight'], p['bias'], padding=1))]
  for lvl in range(n_res):
    for _ in range(m.num_res_blocks):
      idx = P.i
      h = tap(idx, res_block(P.take(), hs[-1], temb, cfg, train=train, drop_mask=mask_for(idx), taps=taps))
      if h.shape[-1] in m.attn_resolutions and m.attention:
        h = tap(P.i, attn_block(P.take(), h, cfg))
      hs.append(h)
    if lvl != n_res - 1:
      if aux:
        idx = P.i
        h = tap(idx, res_block(P.take(), hs[-1], temb, cfg, down=True, train=train,
                               drop_mask=mask_for(idx), taps=taps))
      if m.progressive_input == 'input_skip':
        pyr_in = fir_downsample(pyr_in, taps) if m.fir else mean_down(pyr_in)
        p = P.take()
        h = tap(P.i - 1, F.conv2d(pyr_in, p['Conv_0.weight'], p['Conv_0.bias']) + h)
      elif m.progressive_input == 'residual':
        p = P.take()
        if m.fir:
          pyr_in = fir_conv_downsample(pyr_in, p['Conv2d_0.weight'], taps) + p['Conv2d_0.bias'][None, :, None, None]
        else:
          pyr_in = F.conv2d(F.pad(pyr_in, (0, 1, 0, 1)), p['Conv_0.weight'], p['Conv_0.bias'], stride=2)
        pyr_in = (pyr_in + h) / SQRT2 if m.skip_rescale else pyr_in + h
        h = tap(P.i - 1, pyr_in)
      if aux:
        hs.append(h)

  h = hs[-1]
  if not aux:
    hs.pop()
  idx = P.i
  h = tap(idx, res_block(P.take(), h, temb, cfg, train=train, drop_mask=mask_for(idx), taps=taps))
  h = tap(P.i, attn_block(P.take(), h, cfg))
  idx = P.i
  h = tap(idx, res_block(P.take(), h, temb, cfg, train=train, drop_mask=mask_for(idx), taps=taps))

  pyramid = None
  for lvl in reversed(range(n_res)):
    for _ in range(m.num_res_blocks + (1 if aux else 0)):
      idx = P.i
      h = tap(idx, res_block(P.take(), torch.cat([h, hs.pop()], dim=1), temb, cfg, train=train,
                             drop_mask=mask_for(idx), taps=taps))
    if h.shape[-1] in m.attn_resolutions and m.attention:
      h = tap(P.i, attn_block(P.take(), h, cfg))
    if m.progressive == 'output_skip':
      if lvl != n_res - 1:
        pyramid = fir_upsample(pyramid, taps) if m.fir else nearest_up(pyramid)
      gn, cv = P.take(), P.take()
      ph = F.conv2d(swish(group_norm(h, gn['weight'], gn['bias'])), cv['weight'], cv['bias'], padding=1)
      pyramid = ph if pyramid is None else pyramid + ph
      tap(P.i - 1, pyramid)
    elif m.progressive != 'none':
      raise NotImplementedError("progressive='residual' hits the reference's broken upsample_conv_2d")
    if lvl != 0 and aux:
      idx = P.i
      h = tap(idx, res_block(P.take(), h, temb, cfg, up=True, train=train, drop_mask=mask_for(idx), taps=taps))
  assert not hs

  if m.progressive == 'output_skip':
    h = pyramid
  else:
    gn, cv = P.take(), P.take()
    h = F.conv2d(swish(group_norm(h, gn['weight'], gn['bias'])), cv['weight'], cv['bias'], padding=1)
  assert P.i == P.count(), (P.i, P.count())
  if m.scale_by_sigma:
    h = h / used_sigmas[:, None, None, None]
  return h


# ----------------------------------------------------------------------------- initialisation
def fan_avg_uniform(shape, scale=1., in_axis=1, out_axis=0, gen=None):
  scale = 1e-10 if scale == 0 else scale
  rf = np.prod(shape) / shape[in_axis] / shape[out_axis]
  var = scale / ((shape[in_axis] * rf + shape[out_axis] * rf) / 2)
  return (torch.rand(*shape, generator=gen) * 2. - 1.) * np.sqrt(3 * var)


def make_state_dict(cfg, seed=0, rezero=True):
  """Random reference-format state_dict for `cfg` (same shapes/keys as the reference's
  NCSNpp(config).state_dict()).  `rezero=False` keeps the reference's ~0 initialisation of
  Conv_1 / NIN_3 / output convs; `rezero=True` gives them init_scale 1 so outputs and
  gradients are non-degenerate."""
  m = cfg.model
  gen = torch.Generator().manual_seed(seed)
  zs = 1. if rezero else m.init_scale
  sd = {}
  mods = []

  def add(entries):
    i = len(mods)
    mods.append(entries)
    for k, v in entries.items():
      sd[f'all_modules.{i}.{k}'] = v

  def conv(cin, cout, k=3, scale=1., name=''):
    pre = name + '.' if name else ''
    return {pre + 'weight': fan_avg_uniform((cout, cin, k, k), scale, gen=gen), pre + 'bias': torch.zeros(cout)}

  def gn(c, name=''):
    pre = name + '.' if name else ''
    return {pre + 'weight': torch.ones(c), pre + 'bias': torch.zeros(c)}

  def lin(cin, cout, name=''):
    pre = name + '.' if name else ''
    return {pre + 'weight': fan_avg_uniform((cout, cin), 1., gen=gen), pre + 'bias': torch.zeros(cout)}

  def nin_p(cin, cout, scale, name):
    return {name + '.W': fan_avg_uniform((cin, cout), scale, gen=gen), name + '.b': torch.zeros(cout)}

  nf = m.nf
  # Dense_0 input width = 4 * embed_dim_2 (reference models/ncsnpp.py:76-91,135): nf for Fourier features, else the
  # sinusoidal width (embedding_dim when model.lsgm, nf otherwise)
  lsgm = m.embedding_type.lower() == 'positional' and getattr(m, 'lsgm', False)
  temb_dim = 4 * (m.embedding_dim if lsgm else nf)

  def resblock(cin, cout=None, up=False, down=False):
    cout = cout or cin
    e = {}
    e.update(gn(cin, 'GroupNorm_0'))
    e.update(conv(cin, cout, name='Conv_0'))
    e.update(lin(temb_dim, cout, 'Dense_0'))
    e.update(gn(cout, 'GroupNorm_1'))
    e.update(conv(cout, cout, scale=zs, name='Conv_1'))
    if cin != cout or up or down:
      e.update(conv(cin, cout, k=1, name='Conv_2'))
    add(e)

  def attn(c):
    e = {}
    e.update(gn(c, 'GroupNorm_0'))
    for j in range(3):
      e.update(nin_p(c, c, 0.1, f'NIN_{j}'))
    e.update(nin_p(c, c, 0.1 if rezero else (m.init_scale or 1e-10), 'NIN_3'))
    add(e)

  if m.embedding_type.lower() == 'fourier':
    add({'W': torch.randn(nf, generator=gen) * m.fourier_scale})
    embed_dim = 2 * nf
  else:
    embed_dim = m.embedding_dim if lsgm else nf
  if m.conditional:
    add(lin(embed_dim, temb_dim))
    add(lin(temb_dim, temb_dim))
  ch = cfg.data.num_channels
  add(conv(ch, nf))
  n_res = len(m.ch_mult)
  res = [cfg.data.image_size // 2 ** i for i in range(n_res)]
  aux = m.auxiliary_resblock
  hs_c = [nf]
  cin = nf
  pyr_ch = ch
  for lvl in range(n_res):
    for _ in range(m.num_res_blocks):
      cout = nf * m.ch_mult[lvl]
      resblock(cin, cout)
      cin = cout
      if res[lvl] in m.attn_resolutions and m.attention:
        attn(cin)
      hs_c.append(cin)
    if lvl != n_res - 1:
      if aux:
        resblock(cin, down=True)
      if m.progressive_input == 'input_skip':
        add(conv(pyr_ch, cin, k=1, name='Conv_0'))
      elif m.progressive_input == 'residual':
        add(conv(pyr_ch, cin, name='Conv2d_0' if m.fir else 'Conv_0'))
        pyr_ch = cin
      if aux:
        hs_c.append(cin)
  cin = hs_c[-1]
  if not aux:
    hs_c.pop()
  resblock(cin)
  attn(cin)
  resblock(cin)
  for lvl in reversed(range(n_res)):
    for _ in range(m.num_res_blocks + (1 if aux else 0)):
      cout = nf * m.ch_mult[lvl]
      resblock(cin + hs_c.pop(), cout)
      cin = cout
    if res[lvl] in m.attn_resolutions and m.attention:
      attn(cin)
    if m.progressive == 'output_skip':
      add(gn(cin))
      add(conv(cin, ch, scale=zs))
    if lvl != 0 and aux:
      resblock(cin, up=True)
  assert not hs_c
  if m.progressive != 'output_skip':
    add(gn(cin))
    add(conv(cin, ch, scale=zs))
  sd['sigmas'] = torch.tensor(np.exp(np.linspace(np.log(m.sigma_max), np.log(m.sigma_min), m.num_scales)))
  return sd
