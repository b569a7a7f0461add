"""NCSN++ / DDPM++ score network on the B200 kernels.

Same constructor, call signature, parameter names and shapes as the reference
`models.ncsnpp.NCSNpp(config, sde)` (reference models/ncsnpp.py:34-432), but the network is executed by
hand-written CUDA (libst_b200) on NHWC activations, and its backward pass is written out explicitly
(one autograd node for the whole network) instead of being recorded op by op:

  * ResnetBlockBigGANpp  (reference models/layerspp.py:225-287)   -> ResBlock.fwd / .bwd
  * AttnBlockpp + NIN    (models/layerspp.py:75-104, layers.py:546-555) -> AttnBlock.fwd / .bwd
  * time embedding MLP   (models/ncsnpp.py:262-292)                -> TimeEmbedding.fwd / .bwd
  * torch.cat skip connections (models/ncsnpp.py:368) are never materialised: GroupNorm, the 1x1
    shortcut and the 3x3 convolutions read the two tensors as one channel axis.

Numeric modes: `compute_dtype=torch.bfloat16` (tensor-core path, fp32 accumulation) or
`torch.float32` (parity path).  Parameters, gradients, optimizer state and all statistics stay fp32.
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .._lib import check, lib
from . import utils
from .params import (ParamStore, _logical_view, _phys_view, init_conv, init_ones, init_zeros, register_owner)

SQRT2 = math.sqrt(2.)
# Emit bias / time-embedding gradient partial sums from the GroupNorm backward kernel instead of re-reading the
# gradient tensors (ST_FUSE_CSUM=0 switches back to explicit column-sum passes; both paths are parity-tested).
FUSE_CSUM = os.environ.get('ST_FUSE_CSUM', '1') != '0'
CPAD = 64     # physical channel count of the 3-channel image-side tensors


def _phys_channels(c):
  """Image-side channel counts (3) are stored padded to CPAD; feature channel counts are kept."""
  return c if (c >= 16 and c % 4 == 0) else CPAD


class _Holder(nn.Module):
  """Parameter container that reproduces the reference's module/parameter names."""


def _attach(root, dotted, param):
  parts = dotted.split('.')
  mod = root
  for p in parts[:-1]:
    if not hasattr(mod, p):
      mod.add_module(p, _Holder())
    mod = getattr(mod, p)
  mod.register_parameter(parts[-1], param)


def _fir_kernel(taps, gain, device):
  k = np.asarray(taps, dtype=np.float32)
  k = np.outer(k, k)
  k = k / k.sum() * gain
  return torch.tensor(k, dtype=torch.float32, device=device)


# ===================================================================================== blocks
class Tape:
  """Minimal reverse-mode tape over activation ids (one entry per block)."""

  def __init__(self, enabled):
    self.enabled = enabled
    self.ops = []
    self.next_id = 0

  def new_id(self):
    self.next_id += 1
    return self.next_id

  def record(self, bwd, in_ids, out_id):
    if self.enabled:
      self.ops.append((bwd, in_ids, out_id))


class Act:
  """An activation tensor plus its tape id and, when the GEMM that produced it emitted them, its GroupNorm partial
  sums (ops.Quads): the next GroupNorm then needs no statistics pass over the tensor."""
  __slots__ = ('t', 'id', 'q')

  def __init__(self, t, tape, q=None):
    self.t, self.id, self.q = t, tape.new_id(), q


class NetCtx:
  """Per-call state shared by all blocks."""

  def __init__(self, model, train, tape):
    self.m, self.train, self.tape = model, train, tape
    self.dense = None          # fp32 [B][sum Cout]: Dense_0(act(temb)) of every res-block
    self.d_dense = None        # its gradient, filled slice by slice during backward
    self.taps = model._taps
    self.drop_calls = 0


USE_WT = os.environ.get('ST_DGRAD_WT', '1') != '0'    # data gradients as forward convs over transposed weight copies


def conv_dgrad_w(P, dy, name, cin, kh=3, kw=3, alpha=1.0, dz=None):
  """Data gradient of the convolution whose weight is `name`.  dz (ops.DzRequest): returns (gradient, qpart | None);
  with qpart the gradient already is dz of the GroupNorm that fed the convolution (ops.conv_fwd)."""
  wt = P.ct(name)
  if wt is not None and dy.shape[3] % 64 == 0:
    return ops.conv_fwd(dy, wt, cin, kh, kw, alpha=alpha, dz=dz)
  out = ops.conv_dgrad(dy, P.c(name), cin, kh, kw, alpha=alpha)
  return (out, None) if dz is not None else out


def gn_backward_after(P, pre_gn, dy, qpart, x, x2, G, stats, act, **kw):
  """GroupNorm backward of `pre_gn` behind a data-gradient convolution: the one-pass form when the convolution's
  epilogue produced dz (qpart), the two-phase form otherwise.  kw: p_drop / seed / mask / keepbits (two-phase form only)
  and the common extra / destination / csum / queue arguments."""
  gamma, beta = P.f(pre_gn + '.weight'), P.f(pre_gn + '.bias')
  dgamma, dbeta = P.g(pre_gn + '.weight'), P.g(pre_gn + '.bias')
  if qpart is not None:
    for k in ('p_drop', 'seed', 'mask', 'keepbits'):
      kw.pop(k, None)
    return ops.gn_backward_dz(x, x2, dy, qpart, G, gamma, stats, dgamma, dbeta, **kw)
  return ops.gn_backward(x, x2, dy, G, gamma, beta, stats, act, dgamma, dbeta, **kw)


CSQ = ops.ColsumQueue()      # reductions deferred to the end of NCSNpp._backward (destinations are distinct parameters)


def bias_grad(dst, g2, gs, scale=1.0):
  """dst[c] += scale * sum_rows g2[row][c].  `gs` = list of fp32 (rows, C)-shaped column-sum partials that the
  producers of `g2` emitted as a by-product of their GroupNorm backward kernels (then no pass over g2 is
  needed), or None."""
  C = g2.shape[-1]
  if not ops.PARAM_GRADS:
    return
  if gs:
    CSQ.add_reduce(dst, list(gs), scale)       # one batched launch at the end of the backward pass
  else:
    ops.colsum(g2, 1, g2.shape[0], C, dst, scale=scale, accumulate=True)


def _split_csum(cs, C1, C2):
  """(B, chunks, C1+C2) partial column sums -> 2-D views for the two concatenated inputs."""
  rows = cs.shape[0] * cs.shape[1]
  flat = cs.view(rows, C1 + C2)
  return flat[:, :C1], (flat[:, C1:] if C2 else None)


class ResBlock:
  """ResnetBlockBigGANpp (reference models/layerspp.py:225-287)."""

  def __init__(self, model, idx, cin, cout, up=False, down=False, hw=None):
    self.idx, self.cin, self.cout, self.up, self.down = idx, cin, cout, up, down
    self.out_hw = (hw, hw)      # spatial size of the block's output
    model._resblocks.append(self)
    m = model.config.model
    add = model._add_param
    pre = f'all_modules.{idx}.'
    add(pre + 'GroupNorm_0.weight', (cin,), init=init_ones)
    add(pre + 'GroupNorm_0.bias', (cin,), init=init_zeros)
    add(pre + 'Conv_0.weight', (cout, cin, 3, 3), 'conv', init=init_conv(1.))
    add(pre + 'Conv_0.bias', (cout,), region='conv0_b', init=init_zeros)    # same order as the Dense_0 columns
    self.dense_off = model._dense_cols
    model._dense_cols += cout
    add(pre + 'Dense_0.weight', (cout, model.temb_dim), 'linear', region='dense_w', init=init_conv(1.))
    add(pre + 'Dense_0.bias', (cout,), region='dense_b', init=init_zeros)
    add(pre + 'GroupNorm_1.weight', (cout,), init=init_ones)
    add(pre + 'GroupNorm_1.bias', (cout,), init=init_zeros)
    add(pre + 'Conv_1.weight', (cout, cout, 3, 3), 'conv', init=init_conv(m.init_scale))
    add(pre + 'Conv_1.bias', (cout,), init=init_zeros)
    self.shortcut = cin != cout or up or down
    if self.shortcut:
      add(pre + 'Conv_2.weight', (cout, cin, 1, 1), 'conv', init=init_conv(1.))
      add(pre + 'Conv_2.bias', (cout,), init=init_zeros)
    self.pre = pre
    self.scale = 1. / SQRT2 if m.skip_rescale else 1.
    self.G0, self.G1 = min(cin // 4, 32), min(cout // 4, 32)
    self.fir = m.fir

  # ---- resampling of (h, x): nearest / box when fir=False, FIR [1,3,3,1] otherwise
  def _resample(self, net, t, t2):
    m = net.m
    if not self.fir:
      return ops.resample2x(t, t2, +1, 1.0) if self.up else ops.resample2x(t, t2, -1, 0.25)
    assert t2 is None
    if self.up:      # upsample_2d: up=2, pad=((p+1)//2+1, p//2) with p = 4-2 (up_or_down_sampling.py:195-224)
      return ops.upfirdn2d_nhwc(t, m._fir_up, up=2, pad=(2, 1))
    return ops.upfirdn2d_nhwc(t, m._fir_down, down=2, pad=(1, 1))

  def _resample_adj(self, net, g):
    m = net.m
    if not self.fir:
      return ops.resample2x(g, None, -1, 1.0) if self.up else ops.resample2x(g, None, +1, 0.25)
    # adjoint of upfirdn2d = upfirdn2d with the flipped FIR and up/down swapped (op/upfirdn2d.py:101-116);
    # [1,3,3,1] (x) [1,3,3,1] is symmetric, so the flip is the identity.
    if self.up:      # forward up=2,pad=(2,1): g_pad0 = kw-pad0-1 = 1, g_pad1 = in*2 - out + pad0 - 2 + 1 = 1
      return ops.upfirdn2d_nhwc(g, m._fir_up, down=2, pad=(1, 1))
    return ops.upfirdn2d_nhwc(g, m._fir_down, up=2, pad=(2, 1))   # forward down=2,pad=(1,1)

  def fwd(self, net, xa, xb=None):
    """xa: Act (B,H,W,C1); xb: optional skip Act (B,H,W,C2) concatenated on channels."""
    P, m = net.m.P, net.m
    pre = self.pre
    x1, x2 = xa.t, (xb.t if xb is not None else None)
    if x2 is not None and (self.up or self.down) and self.fir:
      raise NotImplementedError('FIR resampling of a concatenated input')
    a0, st0 = ops.gn_norm_act(x1, x2, self.G0, P.f(pre + 'GroupNorm_0.weight'), P.f(pre + 'GroupNorm_0.bias'), act=1,
                              quads=(xa.q, xb.q if xb is not None else None))
    xr = None
    if self.up or self.down:
      a0 = self._resample(net, a0, None)
      xr = self._resample(net, x1, x2)
    B, H, W, _ = a0.shape
    h1, q_h1 = ops.conv_fwd(a0, P.c(pre + 'Conv_0.weight'), self.cout, bias=P.f(pre + 'Conv_0.bias'),
                            rowbias=net.dense[:, self.dense_off:], rowbias_ld=net.dense.shape[1], want_quads=True)
    p_drop, seed, mask, keepbits = 0., 0, None, None
    if net.train and m.dropout > 0:
      mask = m._mask_for(self.idx, h1)
      if mask is None:
        p_drop, seed = m.dropout, m._next_seed(self.idx)
        if net.tape.enabled:      # keep flags (1 bit per element) for the two backward passes
          keepbits = torch.empty(h1.numel() // 8, dtype=torch.uint8, device=h1.device)
    a1, st1 = ops.gn_norm_act(h1, None, self.G1, P.f(pre + 'GroupNorm_1.weight'), P.f(pre + 'GroupNorm_1.bias'), act=1,
                              p_drop=p_drop, seed=seed, mask=mask, keepbits=keepbits, quads=(q_h1, None))
    if self.shortcut:
      if xr is not None:
        sc = ops.conv_fwd(xr, P.c(pre + 'Conv_2.weight'), self.cout, 1, 1, bias=P.f(pre + 'Conv_2.bias'))
      else:
        sc = ops.conv_fwd(x1, P.c(pre + 'Conv_2.weight'), self.cout, 1, 1, x2=x2, bias=P.f(pre + 'Conv_2.bias'))
    else:
      sc = x1
    out, q_out = ops.conv_fwd(a1, P.c(pre + 'Conv_1.weight'), self.cout, bias=P.f(pre + 'Conv_1.bias'), residual=sc,
                              alpha=self.scale, want_quads=True)
    y = Act(out, net.tape, q_out)
    if net.tape.enabled:
      saved = (x1, x2, st0, a0, xr, h1, st1, a1, p_drop, seed, mask, keepbits)
      net.tape.record(lambda g, acc, gs: self.bwd(net, saved, g, acc, gs),
                      (xa.id,) + ((xb.id,) if xb is not None else ()), y.id)
    if net.taps is not None:
      net.taps[self.idx] = out
    return y

  def bwd(self, net, saved, g, acc, gs=None):
    """g: d(out); acc: existing gradient tensors of (xa[, xb]) to accumulate into, or None; gs: column-sum
    partials of g (see bias_grad).  Returns (input gradients, their column-sum partials)."""
    P = net.m.P
    pre, s = self.pre, self.scale
    x1, x2, st0, a0, xr, h1, st1, a1, p_drop, seed, mask, keepbits = saved
    B, H, W, Co = g.shape
    npix = B * H * W
    g2 = g.view(npix, Co)
    # ---- Conv_1 (and the 1/sqrt2 output scale)
    bias_grad(P.g(pre + 'Conv_1.bias'), g2, gs, s)
    ops.conv_wgrad(g, a1, P.g(pre + 'Conv_1.weight'), alpha=s)
    # ---- GroupNorm_1 + SiLU + dropout: dz and its group sums come out of the data-gradient GEMM's epilogue where the
    # shape allows (ops.dz_applicable), leaving one streaming pass
    req1 = None
    if ops.dz_applicable(h1, None, mask, p_drop, keepbits):
      req1 = ops.DzRequest(h1, self.G1, P.f(pre + 'GroupNorm_1.weight'), P.f(pre + 'GroupNorm_1.bias'), st1, 1, p_drop,
                           keepbits)
    da1 = conv_dgrad_w(P, g, pre + 'Conv_1.weight', Co, alpha=s, dz=req1)
    da1, qp1 = da1 if req1 is not None else (da1, None)
    r1 = gn_backward_after(P, pre + 'GroupNorm_1', da1, qp1, h1, None, self.G1, st1, 1, p_drop=p_drop, seed=seed,
                           mask=mask, keepbits=keepbits, want_csum=FUSE_CSUM, queue=CSQ)
    dh1 = r1[0]
    del da1
    # ---- temb projection gradient = per-image column sums of dh1 (by-product of the kernel above, or an explicit
    # pass); their sum over images is the Conv_0.bias / Dense_0.bias gradient (TimeEmbedding.bwd)
    if not ops.PARAM_GRADS:
      pass                                    # input-gradient-only pass (likelihood divergence): temb is a constant
    elif FUSE_CSUM:
      CSQ.add_groups(net.d_dense[:, self.dense_off:self.dense_off + Co], r1[2])
    else:
      dd = torch.empty((B, Co), dtype=torch.float32, device=g.device)
      ops.colsum(dh1.view(npix, Co), B, H * W, Co, dd)
      net.d_dense[:, self.dense_off:self.dense_off + Co].copy_(dd)
    ops.conv_wgrad(dh1, a0, P.g(pre + 'Conv_0.weight'))
    req0 = None
    if not (self.up or self.down) and ops.dz_applicable(x1, x2, None, 0., None):
      req0 = ops.DzRequest(x1, self.G0, P.f(pre + 'GroupNorm_0.weight'), P.f(pre + 'GroupNorm_0.bias'), st0, 1)
    da0 = conv_dgrad_w(P, dh1, pre + 'Conv_0.weight', self.cin, dz=req0)
    da0, qp0 = da0 if req0 is not None else (da0, None)
    del dh1
    # ---- shortcut
    extra, extra_scale = None, 1.0
    if self.shortcut:
      bias_grad(P.g(pre + 'Conv_2.bias'), g2, gs, s)
      if xr is not None:
        ops.conv_wgrad(g, xr, P.g(pre + 'Conv_2.weight'), 1, 1, alpha=s)
      else:
        ops.conv_wgrad(g, x1, P.g(pre + 'Conv_2.weight'), 1, 1, x2=x2, alpha=s)
      extra = conv_dgrad_w(P, g, pre + 'Conv_2.weight', self.cin, 1, 1, alpha=s)
    else:
      extra, extra_scale = g, s
    if self.up or self.down:
      da0 = self._resample_adj(net, da0)
      extra = self._resample_adj(net, extra)
    # ---- GroupNorm_0 + SiLU, plus the shortcut gradient, split over the two inputs
    a1_acc = acc[0]
    a2_acc = acc[1] if x2 is not None else None
    r0 = gn_backward_after(P, pre + 'GroupNorm_0', da0, qp0, x1, x2, self.G0, st0, 1, extra=extra,
                           extra_scale=extra_scale, dx1=a1_acc, accum1=a1_acc is not None, dx2=a2_acc,
                           accum2=a2_acc is not None, want_csum=FUSE_CSUM, queue=CSQ)
    dx1, dx2 = r0[0], r0[1]
    c1, c2 = _split_csum(r0[2], x1.shape[3], 0 if x2 is None else x2.shape[3]) if FUSE_CSUM else (None, None)
    return ((dx1,), (c1,)) if x2 is None else ((dx1, dx2), (c1, c2))


class AttnBlock:
  """AttnBlockpp with its four NIN projections (reference models/layerspp.py:75-104)."""

  def __init__(self, model, idx, c):
    self.idx, self.c = idx, c
    m = model.config.model
    add = model._add_param
    pre = f'all_modules.{idx}.'
    add(pre + 'GroupNorm_0.weight', (c,), init=init_ones)
    add(pre + 'GroupNorm_0.bias', (c,), init=init_zeros)
    # NIN_0..2 W / b are laid out back to back so that q,k,v come from ONE (3C x C) GEMM
    self.names_w = [pre + f'NIN_{j}.W' for j in range(3)]
    self.names_b = [pre + f'NIN_{j}.b' for j in range(3)]
    for j in range(3):
      add(pre + f'NIN_{j}.W', (c, c), 'nin', init=init_conv(0.1), pack=pre + 'qkv.W')
      add(pre + f'NIN_{j}.b', (c,), init=init_zeros, pack=pre + 'qkv.b')
    add(pre + 'NIN_3.W', (c, c), 'nin', init=init_conv(m.init_scale))
    add(pre + 'NIN_3.b', (c,), init=init_zeros)
    self.pre = pre
    self.scale = 1. / SQRT2 if m.skip_rescale else 1.
    self.G = min(c // 4, 32)

  def fwd(self, net, xa):
    P = net.m.P
    pre, C = self.pre, self.c
    x = xa.t
    B, H, W, _ = x.shape
    L, npix = H * W, B * H * W
    h, st = ops.gn_norm_act(x, None, self.G, P.f(pre + 'GroupNorm_0.weight'), P.f(pre + 'GroupNorm_0.bias'), act=0,
                            quads=(xa.q, None))
    wqkv, bqkv = P.c_group(self.names_w), P.f_group(self.names_b)        # (3C, C), (3C,)
    qkv = ops.gemm_nt(h.view(npix, C), wqkv, bias=bqkv)                 # (npix, 3C)
    if ops.attn_fused_ok(L, C, qkv.dtype):
      # logits in tensor memory, probabilities in shared memory: one kernel (csrc/attn_tc.cu); p is only written
      # when the backward pass will need it
      o, p = ops.attn_fwd(qkv, B, L, C, float(C) ** -0.5, save_p=net.tape.enabled)
    else:
      # logits[b][i][j] = sum_c q[b,i,c] k[b,j,c]   (einsum 'bchw,bcij->bhwij')
      logits = ops.gemm_nt(qkv, qkv[:, C:], out_dtype=torch.float32, M=L, N=L, K=C, lda=3 * C, ldb=3 * C, batch=B,
                           sAb=L * 3 * C, sBb=L * 3 * C, sCb=L * L)
      p = ops.softmax_fwd(logits, L, float(C) ** -0.5, x.dtype)           # (B, L, L)
      del logits
      o = ops.gemm_nn(p, qkv[:, 2 * C:], C, M=L, K=L, lda=L, ldb=3 * C, batch=B, sAb=L * L, sBb=L * 3 * C,
                      sCb=L * C)                                          # (B, L, C)
    out, q_out = ops.gemm_nt(o.view(npix, C), P.c(pre + 'NIN_3.W'), bias=P.f(pre + 'NIN_3.b'), residual=x.view(npix, C),
                             alpha=self.scale, quads_hw=L)
    out = out.view(B, H, W, C)
    y = Act(out, net.tape, q_out)
    if net.tape.enabled:
      saved = (x, st, h, qkv, p, o)
      net.tape.record(lambda g, acc, gs: self.bwd(net, saved, g, acc, gs), (xa.id,), y.id)
    if net.taps is not None:
      net.taps[self.idx] = out
    return y

  def bwd(self, net, saved, g, acc, gs=None):
    P = net.m.P
    pre, C, s = self.pre, self.c, self.scale
    x, st, h, qkv, p, o = saved
    B, H, W, _ = x.shape
    L, npix = H * W, B * H * W
    g2 = g.view(npix, C)
    # ---- NIN_3
    bias_grad(P.g(pre + 'NIN_3.b'), g2, gs, s)
    if ops.PARAM_GRADS:
      ops.gemm_tn(g2, o.view(npix, C), C, C, npix, out=P.g(pre + 'NIN_3.W'), alpha=s, accumulate=True)
    do = ops.gemm_nn(g2, P.c(pre + 'NIN_3.W'), C, alpha=s)              # (npix, C): g W3 (W3 is [out][in])
    # ---- attention core
    dqkv = torch.empty_like(qkv)
    bs = dict(batch=B)
    # dV[j][c] = sum_i p[i][j] do[i][c]
    ops.gemm_tn(p, do, L, C, L, out=dqkv[:, 2 * C:], lda=L, ldb=C, ldc=3 * C, sAb=L * L, sBb=L * C, sCb=L * 3 * C, **bs)
    # dP[i][j] = sum_c do[i][c] v[j][c]
    dp = ops.gemm_nt(do, qkv[:, 2 * C:], out_dtype=torch.float32, M=L, N=L, K=C, lda=C, ldb=3 * C, sAb=L * C,
                     sBb=L * 3 * C, sCb=L * L, **bs)
    ds = ops.softmax_bwd(p, dp, L, float(C) ** -0.5)
    del dp
    # dQ[i][c] = sum_j ds[i][j] k[j][c];  dK[j][c] = sum_i ds[i][j] q[i][c]
    ops.gemm_nn(ds, qkv[:, C:], C, out=dqkv, M=L, K=L, lda=L, ldb=3 * C, ldc=3 * C, sAb=L * L, sBb=L * 3 * C,
                sCb=L * 3 * C, **bs)
    ops.gemm_tn(ds, qkv, L, C, L, out=dqkv[:, C:], lda=L, ldb=3 * C, ldc=3 * C, sAb=L * L, sBb=L * 3 * C,
                sCb=L * 3 * C, **bs)
    del ds
    # ---- q,k,v projections
    if ops.PARAM_GRADS:
      ops.colsum(dqkv, 1, npix, 3 * C, P.g_group(self.names_b), accumulate=True)
      ops.gemm_tn(dqkv, h.view(npix, C), 3 * C, C, npix, out=P.g_group(self.names_w), accumulate=True)
    dh = ops.gemm_nn(dqkv, P.c_group(self.names_w), C).view(B, H, W, C)
    # ---- GroupNorm (no activation) + residual
    r = ops.gn_backward(x, None, dh, self.G, P.f(pre + 'GroupNorm_0.weight'), P.f(pre + 'GroupNorm_0.bias'), st, 0,
                        P.g(pre + 'GroupNorm_0.weight'), P.g(pre + 'GroupNorm_0.bias'), extra=g, extra_scale=s,
                        dx1=acc[0], accum1=acc[0] is not None, want_csum=FUSE_CSUM, queue=CSQ)
    return (r[0],), ((_split_csum(r[2], C, 0)[0] if FUSE_CSUM else None),)


class ConvBlock:
  """Plain 3x3 / 1x1 convolution (ddpm_conv3x3 / conv1x1, reference models/layers.py:100-124)."""

  def __init__(self, model, idx, cin, cout, k=3, init_scale=1., name='', is_input=False):
    """cin / cout are the LOGICAL channel counts; image-side (3-channel) axes are stored padded to
    CPAD so that the convolution runs on the common 64-wide K/N blocks."""
    self.idx, self.k, self.is_input = idx, k, is_input
    self.cin_l, self.cout_l = cin, cout
    self.cin, self.cout = _phys_channels(cin), _phys_channels(cout)
    pre = f'all_modules.{idx}.' + (name + '.' if name else '')
    model._add_param(pre + 'weight', (cout, cin, k, k), 'conv', init=init_conv(init_scale), pad=(self.cout, self.cin))
    model._add_param(pre + 'bias', (cout,), init=init_zeros, pad=(self.cout,))
    self.pre = pre

  def fwd(self, net, xa, record_tap=True):
    P = net.m.P
    out, q_out = ops.conv_fwd(xa.t, P.c(self.pre + 'weight'), self.cout, self.k, self.k, bias=P.f(self.pre + 'bias'),
                              want_quads=True)
    y = Act(out, net.tape, q_out)
    if net.tape.enabled:
      x = xa.t
      net.tape.record(lambda g, acc, gs: self.bwd(net, x, g, acc, need_dx=net.need_dx or not self.is_input, gs=gs),
                      (xa.id,), y.id)
    if net.taps is not None and record_tap:
      net.taps[self.idx] = out
    return y

  def bwd(self, net, x, g, acc, need_dx=True, gs=None):
    P = net.m.P
    B, H, W, Co = g.shape
    bias_grad(P.g(self.pre + 'bias'), g.view(-1, Co), gs)
    ops.conv_wgrad(g, x, P.g(self.pre + 'weight'), self.k, self.k)
    if not need_dx:
      return (None,), (None,)
    dx = conv_dgrad_w(P, g, self.pre + 'weight', self.cin, self.k, self.k)
    if acc[0] is not None:
      ops.axpby(acc[0], dx, out=acc[0])
      dx = acc[0]
    return (dx,), (None,)


class NormActConv:
  """Output head: GroupNorm -> SiLU -> conv3x3 (reference models/ncsnpp.py:250-254, 422-425)."""

  def __init__(self, model, idx_gn, idx_conv, cin, cout, init_scale):
    self.idx_gn, self.idx_conv, self.cin, self.cout = idx_gn, idx_conv, cin, cout
    self.pg = f'all_modules.{idx_gn}.'
    model._add_param(self.pg + 'weight', (cin,), init=init_ones)
    model._add_param(self.pg + 'bias', (cin,), init=init_zeros)
    self.conv = ConvBlock(model, idx_conv, cin, cout, 3, init_scale)
    self.G = min(cin // 4, 32)

  def fwd(self, net, xa, res=None):
    """`res`: optional Act added to the output (the up-sampled output pyramid, ncsnpp.py:391-396)."""
    P = net.m.P
    x = xa.t
    a, st = ops.gn_norm_act(x, None, self.G, P.f(self.pg + 'weight'), P.f(self.pg + 'bias'), act=1, quads=(xa.q, None))
    out = ops.conv_fwd(a, P.c(self.conv.pre + 'weight'), self.conv.cout, bias=P.f(self.conv.pre + 'bias'),
                       residual=res.t if res is not None else None)
    y = Act(out, net.tape)
    if net.tape.enabled:
      ids = (xa.id,) + ((res.id,) if res is not None else ())
      net.tape.record(lambda g, acc, gs: self.bwd(net, (x, st, a), g, acc, gs), ids, y.id)
    if net.taps is not None:
      net.taps[self.idx_conv] = out
    return y

  def bwd(self, net, saved, g, acc, gs=None):
    P = net.m.P
    x, st, a = saved
    (da,), _ = self.conv.bwd(net, a, g, (None,), gs=gs)
    r = ops.gn_backward(x, None, da, self.G, P.f(self.pg + 'weight'), P.f(self.pg + 'bias'), st, 1,
                        P.g(self.pg + 'weight'), P.g(self.pg + 'bias'), dx1=acc[0], accum1=acc[0] is not None,
                        want_csum=FUSE_CSUM, queue=CSQ)
    dx = r[0]
    c = _split_csum(r[2], x.shape[3], 0)[0] if FUSE_CSUM else None
    if len(acc) == 1:
      return (dx,), (c,)
    if acc[1] is not None:
      ops.axpby(acc[1], g, out=acc[1])
      return (dx, acc[1]), (c, None)
    return (dx, g), (c, None)


class ImageResample:
  """2x FIR / nearest-box resampling of an image-side tensor (pyramid_upsample / pyramid_downsample without
  convolution, reference models/ncsnpp.py:112-124, layerspp.py:107-176 with with_conv=False)."""

  def __init__(self, model, up):
    self.up, self.fir = up, model.config.model.fir

  def _run(self, net, t, up):
    m = net.m
    if up:
      return ops.upfirdn2d_nhwc(t, m._fir_up, up=2, pad=(2, 1)) if self.fir else ops.resample2x(t, None, +1, 1.0)
    return ops.upfirdn2d_nhwc(t, m._fir_down, down=2, pad=(1, 1)) if self.fir else ops.resample2x(t, None, -1, 0.25)

  def _adj(self, net, g, up):
    m = net.m
    if up:
      return ops.upfirdn2d_nhwc(g, m._fir_up, down=2, pad=(1, 1)) if self.fir else ops.resample2x(g, None, -1, 1.0)
    return ops.upfirdn2d_nhwc(g, m._fir_down, up=2, pad=(2, 1)) if self.fir else ops.resample2x(g, None, +1, 0.25)

  def fwd(self, net, xa, image_side_input=False):
    y = Act(self._run(net, xa.t, self.up), net.tape)
    if net.tape.enabled:
      def bwd(g, acc, gs=None):
        if image_side_input and not net.need_dx:
          return (None,), (None,)
        d = self._adj(net, g, self.up)
        if acc[0] is not None:
          ops.axpby(acc[0], d, out=acc[0])
          d = acc[0]
        return (d,), (None,)
      net.tape.record(bwd, (xa.id,), y.id)
    return y


class CombineBlock:
  """Combine(method='sum'): h + conv1x1(image pyramid) (reference models/layerspp.py:57-72)."""

  def __init__(self, model, idx, cin_img, cout):
    if model.config.model.progressive_combine != 'sum':
      raise NotImplementedError("progressive_combine='cat' is not built")
    self.idx = idx
    self.conv = ConvBlock(model, idx, cin_img, cout, 1, name='Conv_0')

  def fwd(self, net, pyr, ha):
    P, c = net.m.P, self.conv
    out, q_out = ops.conv_fwd(pyr.t, P.c(c.pre + 'weight'), c.cout, 1, 1, bias=P.f(c.pre + 'bias'), residual=ha.t,
                              want_quads=True)
    y = Act(out, net.tape, q_out)
    if net.tape.enabled:
      x = pyr.t
      net.tape.record(lambda g, acc, gs: self.bwd(net, x, g, acc, gs), (pyr.id, ha.id), y.id)
    if net.taps is not None:
      net.taps[self.idx] = out
    return y

  def bwd(self, net, x, g, acc, gs=None):
    (dp,), _ = self.conv.bwd(net, x, g, (acc[0],), need_dx=net.need_dx, gs=gs)
    if acc[1] is not None:
      ops.axpby(acc[1], g, out=acc[1])
      return (dp, acc[1]), (None, None)
    return (dp, g), (None, None)


class PyramidDownConv:
  """Downsample(with_conv=True) of the input pyramid followed by the residual merge
  (pyr + h)/sqrt2 (reference models/layerspp.py:142-176, up_or_down_sampling.py:144-178,
  models/ncsnpp.py:337-344): FIR pre-filter (pad 2,2) -> 3x3 stride-2 convolution as im2col + GEMM."""

  def __init__(self, model, idx, cin, cout, image_side):
    m = model.config.model
    self.idx, self.fir, self.image_side = idx, m.fir, image_side
    self.cin = _phys_channels(cin)
    self.cout = cout
    self.pre = f'all_modules.{idx}.' + ('Conv2d_0.' if m.fir else 'Conv_0.')
    model._add_param(self.pre + 'weight', (cout, cin, 3, 3), 'conv', init=init_conv(1.), pad=(cout, self.cin))
    model._add_param(self.pre + 'bias', (cout,), init=init_zeros)
    self.scale = 1. / SQRT2 if m.skip_rescale else 1.

  def fwd(self, net, pyr, ha):
    P, m = net.m.P, net.m
    x = pyr.t
    B, H, W, C = x.shape
    y = ops.upfirdn2d_nhwc(x, m._fir_down, pad=(2, 2)) if self.fir else x
    cols = ops.im2col(y, 3, 3, 2, 0, H // 2, W // 2)
    out, q_out = ops.gemm_nt(cols, P.c(self.pre + 'weight'), bias=P.f(self.pre + 'bias'),
                             residual=ha.t.view(-1, self.cout), alpha=self.scale, quads_hw=(H // 2) * (W // 2))
    o = Act(out.view(B, H // 2, W // 2, self.cout), net.tape, q_out)
    out = o.t
    if net.tape.enabled:
      net.tape.record(lambda g, acc, gs: self.bwd(net, (cols, x.shape, y.shape), g, acc, gs), (pyr.id, ha.id), o.id)
    if net.taps is not None:
      net.taps[self.idx] = out
    return o

  def bwd(self, net, saved, g, acc, gs=None):
    P, m, s = net.m.P, net.m, self.scale
    cols, xshape, yshape = saved
    B, H, W, C = xshape
    g2 = g.view(-1, self.cout)
    rows = g2.shape[0]
    bias_grad(P.g(self.pre + 'bias'), g2, gs, s)
    if ops.PARAM_GRADS:
      ops.gemm_tn(g2, cols, self.cout, cols.shape[1], rows, out=P.g(self.pre + 'weight'), alpha=s, accumulate=True)
    if acc[1] is not None:
      dh = ops.axpby(acc[1], g, 1.0, s, out=acc[1])
    else:
      dh = ops.axpby(g, None, s)
    if self.image_side and not net.need_dx:
      return (None, dh), (None, None)
    dcols = ops.gemm_nn(g2, P.c(self.pre + 'weight'), cols.shape[1], alpha=s)
    dy = ops.col2im(dcols, yshape, 3, 3, 2, 0, H // 2, W // 2)
    dp = ops.upfirdn2d_nhwc(dy, m._fir_down, pad=(1, 1)) if self.fir else dy
    if acc[0] is not None:
      ops.axpby(acc[0], dp, out=acc[0])
      dp = acc[0]
    return (dp, dh), (None, None)


class TimeEmbedding:
  """Embedding -> Linear -> SiLU -> Linear, then SiLU -> all Dense_0 projections at once
  (reference models/ncsnpp.py:262-292 and layerspp.py:272-274)."""

  def __init__(self, model, first_idx):
    m = model.config.model
    nf = m.nf
    self.fourier = m.embedding_type.lower() == 'fourier'
    if not self.fourier and m.embedding_type.lower() != 'positional':
      raise ValueError(f'embedding type {m.embedding_type} unknown.')
    i = first_idx
    if self.fourier:
      model._add_param(f'all_modules.{i}.W', (nf,), trainable=False,
                       init=lambda shape, gen: torch.randn(*shape, generator=gen) * m.fourier_scale)
      self.w_name = f'all_modules.{i}.W'
      i += 1
      self.embed_dim = 2 * nf
    else:
      # `model.lsgm`: the sinusoidal embedding has `embedding_dim` entries instead of nf and the whole time-embedding
      # MLP (and every Dense_0 input) is 4*embedding_dim wide (reference models/ncsnpp.py:86-91,135,279-283)
      self.embed_dim = m.embedding_dim if getattr(m, 'lsgm', False) else nf
    td = self.td = model.temb_dim
    self.fp32 = bool(getattr(m, 'temb_fp32', False))
    self.l0, self.l1 = f'all_modules.{i}.', f'all_modules.{i + 1}.'
    model._add_param(self.l0 + 'weight', (td, self.embed_dim), 'linear', init=init_conv(1.))
    model._add_param(self.l0 + 'bias', (td,), init=init_zeros)
    model._add_param(self.l1 + 'weight', (td, td), 'linear', init=init_conv(1.))
    model._add_param(self.l1 + 'bias', (td,), init=init_zeros)
    self.next_idx = i + 2

  def fwd(self, net, time_cond):
    """Embedding features (sin / cos of arguments up to 999) are always computed in fp32.  The two Linear
    layers of the MLP run in the compute dtype on the tensor cores by default; `model.temb_fp32 = True` keeps them in
    fp32 on the master weights instead (five B x 512 x 512 fp32-FMA GEMMs per step, +0.3 ms at B = 512: measured to make
    no difference to the bf16 gradient error, which is dominated by the 55 blocks behind them)."""
    m, P = net.m, net.m.P
    cd = m.compute_dtype
    f32 = cd == torch.float32 or self.fp32
    if self.fourier:
      emb = ops.fourier_embedding(time_cond, P.f(self.w_name))
    else:
      emb = ops.timestep_embedding(time_cond, self.embed_dim)
    W = P.f if f32 else P.c
    emb_c = emb if f32 else ops.cast(emb, cd)
    e0 = ops.gemm_nt(emb_c, W(self.l0 + 'weight'), out_dtype=torch.float32, bias=P.f(self.l0 + 'bias'))
    a0 = ops.silu(e0)
    a0_c = a0 if f32 else ops.cast(a0, cd)
    temb = ops.gemm_nt(a0_c, W(self.l1 + 'weight'), out_dtype=torch.float32, bias=P.f(self.l1 + 'bias'))
    at = ops.silu(temb)
    at_c = ops.cast(at, cd) if cd != torch.float32 else at
    wd, bd = P.c_region('dense_w').view(-1, self.td), P.f_region('dense_b')
    net.dense = ops.gemm_nt(at_c, wd, out_dtype=torch.float32, bias=bd)      # (B, sum Cout)
    if net.tape.enabled:
      net.d_dense = torch.zeros_like(net.dense)
      net.temb_saved = (emb_c, e0, a0_c, temb, at_c)

  def bwd(self, net):
    m, P = net.m, net.m.P
    cd = m.compute_dtype
    f32 = cd == torch.float32 or self.fp32
    emb_c, e0, a0_c, temb, at_c = net.temb_saved
    B = emb_c.shape[0]
    nd = net.d_dense.shape[1]
    td = self.td
    dd = net.d_dense
    # column sums of d_dense are the gradient of every Dense_0.bias AND of every Conv_0.bias (both are added to
    # the same pre-GroupNorm_1 activation); the two bias regions are laid out in the same block order
    colsum = torch.empty(nd, dtype=torch.float32, device=dd.device)
    ops.colsum(dd, 1, B, nd, colsum)
    ops.axpby(P.g_region('dense_b'), colsum, out=P.g_region('dense_b'))
    ops.axpby(P.g_region('conv0_b'), colsum, out=P.g_region('conv0_b'))
    dd_c = ops.cast(dd, cd) if cd != torch.float32 else dd
    ops.gemm_tn(dd_c, at_c, nd, td, B, out=P.g_region('dense_w').view(nd, td), accumulate=True, split_k=1)
    d_at = ops.gemm_nn(dd_c, P.c_region('dense_w').view(nd, td), td, out_dtype=torch.float32)
    d_temb = ops.silu_bwd(temb, d_at)
    ops.colsum(d_temb, 1, B, td, P.g(self.l1 + 'bias'), accumulate=True)
    d_temb_c = d_temb if f32 else ops.cast(d_temb, cd)
    W = P.f if f32 else P.c
    ops.gemm_tn(d_temb_c, a0_c, td, td, B, out=P.g(self.l1 + 'weight'), accumulate=True, split_k=1)
    d_a0 = ops.gemm_nn(d_temb_c, W(self.l1 + 'weight'), td, out_dtype=torch.float32)
    d_e0 = ops.silu_bwd(e0, d_a0)
    ops.colsum(d_e0, 1, B, td, P.g(self.l0 + 'bias'), accumulate=True)
    d_e0_c = d_e0 if f32 else ops.cast(d_e0, cd)
    ops.gemm_tn(d_e0_c, emb_c, td, self.embed_dim, B, out=P.g(self.l0 + 'weight'), accumulate=True, split_k=1)


# ===================================================================================== parameter access
class _ParamAccess:
  """Physical-layout views of parameters (compute dtype / fp32 master / fp32 gradient)."""

  def __init__(self, model):
    self.model = model
    self._cache = {}

  def reset(self):
    self._cache.clear()

  def _view(self, buf_name, name):
    key = (buf_name, name)
    v = self._cache.get(key)
    if v is None:
      v = _phys_view(getattr(self.model, buf_name), self.model.store.by_name[name])
      self._cache[key] = v
    return v

  def c(self, name):
    return self._view('_comp', name)

  def ct(self, name):
    """[Ci][reversed taps][Co] copy of a convolution weight (see NCSNpp._sync_transposed), or None in fp32 mode."""
    if self.model._compT is None:
      return None
    key = ('_compT', name)
    v = self._cache.get(key)
    if v is None:
      e = self.model.store.by_name[name]
      co, ci, kh, kw = e.shape
      v = self.model._compT[e.offset:e.offset + e.numel].view(e.pad[1], kh * kw * e.pad[0])
      self._cache[key] = v
    return v

  def f(self, name):
    return self._view('_flat', name)

  def g(self, name):
    return self._view('_grad', name)

  def _group(self, buf_name, names):
    key = (buf_name, tuple(names))
    v = self._cache.get(key)
    if v is None:
      es = [self.model.store.by_name[n] for n in names]
      off, n = es[0].offset, sum(e.numel for e in es)
      for a, b in zip(es[:-1], es[1:]):
        assert b.offset == a.offset + a.numel, 'grouped parameters must be contiguous'
      v = getattr(self.model, buf_name)[off:off + n]
      if es[0].kind == 'nin':
        v = v.view(-1, es[0].shape[0])
      self._cache[key] = v
    return v

  def c_group(self, names):
    return self._group('_comp', names)

  def f_group(self, names):
    return self._group('_flat', names)

  def g_group(self, names):
    return self._group('_grad', names)

  def _region(self, buf_name, region):
    off, n = self.model.store.region_span(region)
    return getattr(self.model, buf_name)[off:off + n]

  def c_region(self, region):
    return self._region('_comp', region)

  def f_region(self, region):
    return self._region('_flat', region)

  def g_region(self, region):
    return self._region('_grad', region)


# ===================================================================================== the network
class _UNetFn(torch.autograd.Function):
  """One autograd node for the whole network.  Parameter gradients are accumulated directly into the
  model's flat gradient buffer (which `param.grad` views), so backward returns only d(input)."""

  @staticmethod
  def forward(ctx, x, time_cond, anchor, model):
    out, net = model._execute(x, time_cond, record=True)
    ctx.net = net
    ctx.model = model
    return out

  @staticmethod
  def backward(ctx, dout):
    dx = ctx.model._backward(ctx.net, dout, need_dx=ctx.needs_input_grad[0])
    ctx.net = None
    return dx, None, None, None


@utils.register_model(name='ncsnpp')
class NCSNpp(nn.Module):
  """NCSN++ model (same surface as reference models/ncsnpp.py:34-432)."""

  def __init__(self, config, sde=None, compute_dtype=None, seed=None):
    super().__init__()
    self.config = config
    self.sde = config.training.sde
    m = config.model
    if m.resblock_type.lower() != 'biggan':
      raise NotImplementedError("only resblock_type='biggan' (every BASELINE config) is built")
    if m.nonlinearity.lower() != 'swish':
      raise NotImplementedError("only the 'swish' nonlinearity (every shipped NCSN++ config) is built")
    self.progressive, self.progressive_input = m.progressive.lower(), m.progressive_input.lower()
    if self.progressive not in ('none', 'output_skip'):
      raise NotImplementedError("progressive='residual' runs into the reference's broken upsample_conv_2d "
                                "(up_or_down_sampling.py:126) and is not built")
    if self.progressive_input not in ('none', 'input_skip', 'residual'):
      raise ValueError(f'progressive input method {m.progressive_input!r} not recognized.')
    cd = compute_dtype or getattr(m, 'compute_dtype', None) or torch.bfloat16
    self.compute_dtype = {'bf16': torch.bfloat16, 'fp32': torch.float32}.get(cd, cd) if isinstance(cd, str) else cd
    self.nf = nf = m.nf
    positional_lsgm = m.embedding_type.lower() == 'positional' and getattr(m, 'lsgm', False)
    self.temb_dim = 4 * (m.embedding_dim if positional_lsgm else nf)
    self.dropout = m.dropout
    self.scale_by_sigma = m.scale_by_sigma
    self.conditional = m.conditional
    if not m.conditional:
      raise NotImplementedError('unconditional NCSN++ is not built')
    self.centered = config.data.centered
    self.num_scales = m.num_scales
    self.store = ParamStore()
    self._dense_cols = 0
    self._resblocks = []
    self._taps = None
    self.drop_masks = None
    # in-kernel dropout streams are keyed by (seed base, forward-call counter, block index).  The base follows the
    # process's torch seed and its data-parallel rank, so that `torch.manual_seed` selects the masks and ranks draw
    # different ones (the reference draws dropout from torch's per-device CUDA generator)
    self._seed_base = (int(torch.initial_seed()) % (2 ** 31)) ^ 0x5eed
    self._seed_base += int(os.environ.get('RANK', '0'))
    self._calls = 0

    ch = config.data.num_channels
    n_res = len(m.ch_mult)
    res = [config.data.image_size // 2 ** i for i in range(n_res)]
    aux = m.auxiliary_resblock
    attn_on = m.attention

    # ---- build the block list in the reference's module order (models/ncsnpp.py:74-256)
    self.temb = TimeEmbedding(self, 0)
    i = self.temb.next_idx
    self.conv_in = ConvBlock(self, i, ch, nf, 3, is_input=True)
    i += 1
    self.down = []          # per level: [(ResBlock, AttnBlock|None)], down block, input-pyramid block
    hs_c = [nf]
    cin = nf
    pyr_ch = ch
    self.img_down = ImageResample(self, up=False)
    self.img_up = ImageResample(self, up=True)
    for lvl in range(n_res):
      blocks = []
      for _ in range(m.num_res_blocks):
        cout = nf * m.ch_mult[lvl]
        rb = ResBlock(self, i, cin, cout, hw=res[lvl])
        i += 1
        cin = cout
        ab = None
        if res[lvl] in m.attn_resolutions and attn_on:
          ab = AttnBlock(self, i, cin)
          i += 1
        blocks.append((rb, ab))
        hs_c.append(cin)
      dn = pyr = None
      if lvl != n_res - 1:
        if not aux:
          raise NotImplementedError('auxiliary_resblock=False is not built')
        dn = ResBlock(self, i, cin, cin, down=True, hw=res[lvl + 1])
        i += 1
        if self.progressive_input == 'input_skip':
          pyr = CombineBlock(self, i, pyr_ch, cin)
          i += 1
        elif self.progressive_input == 'residual':
          pyr = PyramidDownConv(self, i, pyr_ch, cin, image_side=(lvl == 0))
          i += 1
          pyr_ch = cin
        hs_c.append(cin)
      self.down.append((blocks, dn, pyr))
    self.mid = (ResBlock(self, i, cin, cin, hw=res[-1]), AttnBlock(self, i + 1, cin),
                ResBlock(self, i + 2, cin, cin, hw=res[-1]))
    i += 3
    self.up = []
    for lvl in reversed(range(n_res)):
      blocks = []
      for _ in range(m.num_res_blocks + 1):
        cout = nf * m.ch_mult[lvl]
        skip_c = hs_c.pop()
        blocks.append((ResBlock(self, i, cin + skip_c, cout, hw=res[lvl]), cin, skip_c))
        i += 1
        cin = cout
      ab = None
      if res[lvl] in m.attn_resolutions and attn_on:
        ab = AttnBlock(self, i, cin)
        i += 1
      pout = None
      if self.progressive == 'output_skip':
        pout = NormActConv(self, i, i + 1, cin, ch, m.init_scale)
        i += 2
      upb = None
      if lvl != 0:
        upb = ResBlock(self, i, cin, cin, up=True, hw=res[lvl - 1])
        i += 1
      self.up.append((blocks, ab, pout, upb))
    assert not hs_c
    if self.progressive == 'output_skip':
      self.head = self.up[-1][2]
      self.n_modules = i
    else:
      self.head = NormActConv(self, i, i + 1, cin, ch, m.init_scale)
      self.n_modules = i + 2

    # ---- storage
    self.store.layout()
    gen = torch.Generator().manual_seed(int(seed if seed is not None else torch.initial_seed() % (2 ** 31)))
    flat = torch.zeros(self.store.total, dtype=torch.float32)
    for e in self.store.entries:
      _logical_view(flat, e).copy_(e.init(e.shape, gen))
    self._flat = flat
    self._grad = torch.zeros_like(flat)
    self._comp = flat
    self.P = _ParamAccess(self)
    self.all_modules = nn.ModuleList([_Holder() for _ in range(self.n_modules)])
    self._params = {}
    for e in self.store.entries:
      p = nn.Parameter(_logical_view(self._flat, e), requires_grad=e.trainable)
      self._params[e.name] = p
      _attach(self, e.name, p)
    self.register_buffer('sigmas', torch.tensor(utils.get_sigmas(config)))
    self._fir_up = self._fir_down = None
    self._rebind()

  # ---------------------------------------------------------------- construction helpers
  def _add_param(self, name, shape, kind='vec', region='main', trainable=True, init=None, pad=None, pack=None):
    return self.store.add(name, shape, kind, region, trainable, init, pad, pack)

  def _rebind(self):
    """Re-point every Parameter (and its .grad) at the flat buffers after they moved."""
    for e in self.store.entries:
      p = self._params[e.name]
      p.data = _logical_view(self._flat, e)
      p.grad = _logical_view(self._grad, e) if e.trainable else None
    if self.compute_dtype == torch.float32:
      self._comp = self._flat
    else:
      self._comp = torch.empty(self._flat.shape, dtype=self.compute_dtype, device=self._flat.device)
    # transposed copies of the convolution weights for the data-gradient GEMMs (tensor-core path only)
    self._compT = None
    if self.compute_dtype != torch.float32 and self._flat.is_cuda and USE_WT:
      convs = [e for e in self.store.entries if e.kind == 'conv']
      rows, prefix = [], [0]
      for e in convs:
        co, ci, kh, kw = e.shape
        rows.append([e.offset, e.pad[0], kh * kw, e.pad[1]])
        prefix.append(prefix[-1] + kh * kw * ((e.pad[0] + 31) // 32) * ((e.pad[1] + 31) // 32))
      self._compT = torch.zeros(self._flat.shape, dtype=self.compute_dtype, device=self._flat.device)
      self._wt_table = torch.tensor(rows, dtype=torch.int64, device=self._flat.device)
      self._wt_prefix = torch.tensor(prefix, dtype=torch.int64, device=self._flat.device)
      self._wt_tiles = prefix[-1]
    taps = self.config.model.fir_kernel
    self._fir_up = _fir_kernel(taps, 4., self._flat.device)
    self._fir_down = _fir_kernel(taps, 1., self._flat.device)
    self.P.reset()
    register_owner(self)

  def _apply(self, fn, recurse=True):
    # Move the flat buffers as a whole and rebuild the views (nn.Module._apply would break the sharing).
    new_flat = fn(self._flat)
    if new_flat.dtype != torch.float32:
      raise TypeError('NCSNpp master parameters are fp32; pick the compute dtype with compute_dtype=')
    self._flat = new_flat.contiguous()
    self._grad = fn(self._grad).contiguous()
    for k, b in list(self._buffers.items()):
      if b is not None:
        self._buffers[k] = fn(b)
    self._rebind()
    return self

  def _all_resblocks(self):
    return list(self._resblocks)

  def flat_parameters(self):
    """(flat fp32 params, flat fp32 grads, bool mask of trainable elements)."""
    return self._flat, self._grad

  def trainable_mask(self):
    mask = torch.zeros(self.store.total, dtype=torch.uint8)
    for e in self.store.entries:
      if e.trainable:
        mask[e.offset:e.offset + e.numel] = 1
    return mask.to(self._flat.device)

  def zero_grad(self, set_to_none=False):
    self._grad.zero_()
    for e in self.store.entries:        # restore views if someone set them to None
      p = self._params[e.name]
      if e.trainable and p.grad is None:
        p.grad = _logical_view(self._grad, e)

  def _sync_transposed(self):
    """Refresh the [Ci][reversed taps][Co] weight copies from the compute-dtype weights (one launch)."""
    if self._compT is not None:
      check(lib.st_transpose_conv_weights(ops.ptr(self._comp), ops.ptr(self._compT), ops.dt(self._comp),
                                          ops.ptr(self._wt_table), ops.ptr(self._wt_prefix), self._wt_table.shape[0],
                                          self._wt_tiles, ops.stream()))

  def buffers_for_graph_key(self):
    """Tensors whose addresses a captured CUDA graph of this network depends on."""
    return [self._flat, self._comp]

  def sync_compute_weights(self):
    """Refresh the compute-dtype copy of the parameters (one cast kernel over the flat buffer)."""
    if self._comp is not self._flat:
      ops.cast(self._flat, self.compute_dtype, out=self._comp)

  def _mask_for(self, idx, like):
    if self.drop_masks is None or idx not in self.drop_masks:
      return None
    mk = self.drop_masks[idx]            # NCHW fp32 keep-mask already scaled by 1/(1-p)
    return ops.nchw_to_nhwc(mk.to(like.device).float().contiguous(), like.dtype)

  SEED_STRIDE = 7919      # seed advance per forward call (a captured training graph adds it through device memory)

  def _next_seed(self, idx):
    return self._seed_base * 1000003 + self._calls * self.SEED_STRIDE + idx

  def seed_dropout(self, seed):
    self._seed_base, self._calls = int(seed), 0

  # ---------------------------------------------------------------- execution
  fused_out_scale = True      # forward(..., out_scale=v) multiplies image n of the output by v[n] in the layout kernel

  def forward(self, x, time_cond, out_scale=None):
    """`out_scale` (ours, optional, inference only): per-image factor folded into the final NHWC->NCHW kernel, e.g.
    -1/std(t) for the VP score (models/utils.py:169-170) - one elementwise pass less per sampler step."""
    if not x.is_cuda:
      raise RuntimeError('NCSNpp (B200 build) runs on CUDA tensors only; there is no CPU fallback')
    x = x.float().contiguous()
    time_cond = time_cond.float().contiguous()
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self._params.values())):
      if out_scale is not None:
        raise ValueError('out_scale is an inference-only option')
      return _UNetFn.apply(x, time_cond, self._params[self.head.conv.pre + 'weight'], self)
    out, _ = self._execute(x, time_cond, record=False, out_scale=out_scale)
    return out

  def _execute(self, x, time_cond, record, out_scale=None):
    m = self.config.model
    self._calls += 1
    self.sync_compute_weights()
    tape = Tape(record)
    net = NetCtx(self, self.training, tape)
    self.temb.fwd(net, time_cond)
    if m.embedding_type.lower() == 'fourier':
      used_sigmas = time_cond
    else:
      used_sigmas = self.sigmas[time_cond.long()].float()
    # data in [0,1] is re-centred to [-1,1] here when the loader has not done it (ncsnpp.py:303-305)
    h_in = Act(ops.nchw_to_nhwc(x, self.compute_dtype, CPAD, *((1., 0.) if self.centered else (2., -1.))), tape)
    net.x_id = h_in.id
    h = self.conv_in.fwd(net, h_in)
    hs = [h]
    pyr_in = h_in                      # input pyramid (progressive_input != 'none')
    for lvl, (blocks, dn, pyr) in enumerate(self.down):
      for rb, ab in blocks:
        h = rb.fwd(net, hs[-1])
        if ab is not None:
          h = ab.fwd(net, h)
        hs.append(h)
      if dn is not None:
        h = dn.fwd(net, hs[-1])
        if self.progressive_input == 'input_skip':
          pyr_in = self.img_down.fwd(net, pyr_in, image_side_input=True)
          h = pyr.fwd(net, pyr_in, h)
        elif self.progressive_input == 'residual':
          pyr_in = pyr.fwd(net, pyr_in, h)
          h = pyr_in
        hs.append(h)
    h = hs[-1]
    h = self.mid[0].fwd(net, h)
    h = self.mid[1].fwd(net, h)
    h = self.mid[2].fwd(net, h)
    pyramid = None                     # output pyramid (progressive == 'output_skip')
    for blocks, ab, pout, upb in self.up:
      for rb, _, _ in blocks:
        h = rb.fwd(net, h, hs.pop())
      if ab is not None:
        h = ab.fwd(net, h)
      if pout is not None:
        if pyramid is not None:
          pyramid = self.img_up.fwd(net, pyramid)
        pyramid = pout.fwd(net, h, res=pyramid)
      if upb is not None:
        h = upb.fwd(net, h)
    assert not hs
    h = pyramid if self.progressive == 'output_skip' else self.head.fwd(net, h)
    net.out_id = h.id
    net.out_scale = (1. / used_sigmas).contiguous() if m.scale_by_sigma else None
    if out_scale is not None:
      out_scale = out_scale.float().reshape(-1)
      net.out_scale = (out_scale if net.out_scale is None else net.out_scale * out_scale).contiguous()
    out = ops.nhwc_to_nchw(h.t, x.shape[1], net.out_scale)
    return out, net

  def _backward(self, net, dout, need_dx=False):
    dout = dout.float().contiguous()
    if net.out_scale is not None:
      dout = dout * net.out_scale[:, None, None, None]
    net.need_dx = need_dx
    CSQ.jobs, CSQ.keep, CSQ.blocks_y = [], [], 1       # nothing may survive an aborted backward pass
    self._sync_transposed()
    grads = {net.out_id: ops.nchw_to_nhwc(dout, self.compute_dtype, CPAD)}
    # gsum[id]: column-sum partials of grads[id] emitted by the kernels that produced it (a list), or False once a
    # contribution arrived without partials (then the consumer reduces the tensor itself)
    gsum = {}
    for bwd, in_ids, out_id in reversed(net.tape.ops):
      g = grads.pop(out_id, None)
      if g is None:
        continue
      gs = gsum.pop(out_id, None)
      res, css = bwd(g, tuple(grads.get(i) for i in in_ids), gs if gs else None)
      for i, r, c in zip(in_ids, res, css):
        if r is not None:
          grads[i] = r
          if c is None or gsum.get(i) is False:
            gsum[i] = False
          else:
            gsum.setdefault(i, []).append(c)
    CSQ.flush()
    if ops.PARAM_GRADS:
      self.temb.bwd(net)
    net.tape.ops = []
    if need_dx:
      dx = ops.nhwc_to_nchw(grads[net.x_id], dout.shape[1])
      return dx if self.centered else dx * 2.
    return None
