"""Tensor-level wrappers over the C ABI (no autograd here; the blocks in models/ncsnpp.py write
their own backward).  Activations are NHWC torch tensors (B, H, W, C), fp32 or bf16, contiguous.
Every wrapper enqueues on torch's current CUDA stream and never synchronises.
"""
import ctypes
import math

import numpy as np
import torch

from ._lib import GemmArgs, check, lib

F32, BF16 = 0, 1
_DT = {torch.float32: F32, torch.bfloat16: BF16}
OP_STRIDED, OP_GATHER, OP_DGRADW = 0, 1, 2
BACKEND = {'auto': 0, 'simt': 1, 'tcgen05': 2}

import os

# default backend for st_gemm ('auto' = tcgen05 whenever the problem is expressible, else SIMT); tests flip
# this (or set ST_GEMM_BACKEND) to pin a path
gemm_backend = os.environ.get('ST_GEMM_BACKEND', 'auto')


def dt(t):
  return _DT[t.dtype]


def ptr(t):
  if t is None:
    return None
  assert t.is_cuda, 'libst_b200 works on CUDA tensors only'
  return ctypes.c_void_p(t.data_ptr())


def stream():
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def tc_available():
  return bool(lib.st_tc_available())


# ------------------------------------------------------------------------------------ GEMM
def _gemm(**kw):
  a = GemmArgs()
  a.backend = BACKEND[gemm_backend]
  a.alpha = 1.0
  a.batch = 1
  a.split_k = 1
  keep = []
  for k, v in kw.items():
    if isinstance(v, torch.Tensor):
      keep.append(v)
      v = v.data_ptr()
    setattr(a, k, v)
  check(lib.st_gemm(ctypes.byref(a), stream()))


# GroupNorm statistics as a by-product of the GEMM that produces a tensor (st_gemm_args.gn_part): `Quads` travels with
# the activation; gn_norm_act consumes it instead of reading the tensor for its statistics.  ST_GN_QUADS=0 turns it off.
GN_QUADS = os.environ.get('ST_GN_QUADS', '1') != '0'


class Quads:
  """Per-(rows block, 4 channels) sum / sum of squares of an NHWC tensor: `t` fp32 (M // rows, C // 4, 2)."""
  __slots__ = ('t', 'rows')

  def __init__(self, t, rows):
    self.t, self.rows = t, rows


def _quads_request(kw, out, hw):
  """Adds the gn_part request to a GEMM call; returns a closure that yields the Quads (or None) after the call."""
  if not (GN_QUADS and out.dtype == torch.bfloat16 and hw >= 16):
    return lambda: None
  M, N = kw['M'], kw['N']
  rows = min(hw, 128)
  part = torch.empty((M // rows, max(N // 4, 1), 2), dtype=torch.float32, device=out.device)
  got = ctypes.c_int32(0)
  kw.update(gn_part=part, gn_hw=hw, gn_rows_out=ctypes.pointer(got))
  return lambda: Quads(part, got.value) if got.value == rows else None


def _split_k(M, N, K):
  """Split the reduction so that a weight-gradient GEMM (few output tiles, huge K) fills the GPU."""
  tiles = ((M + 127) // 128) * ((N + 127) // 128)
  want = max(1, (148 * 2) // max(tiles, 1))
  return int(max(1, min(want, K // 1024 if K >= 2048 else 1, 64)))


# GroupNorm backward phase 1 inside the data-gradient GEMM (st_gemm_args.dz_x).  Parity-tested but OFF by default
# (ST_GN_DZ=1 turns it on): measured on B200 at B=512 it removes 0.8 ms of GroupNorm kernels per step and adds 3.3 ms to
# the GEMMs - ~20 ALU instructions per element in 8 epilogue warps outlast the 9216 tensor-core cycles of a 128-channel
# tile.
GN_DZ = os.environ.get('ST_GN_DZ', '0') != '0'


class DzRequest:
  """What the dz epilogue needs: the GroupNorm input `x` (B,H,W,C) whose output the convolution consumed, the
  statistics / parameters of that GroupNorm, the activation flag and the dropout keep bits (or None)."""
  __slots__ = ('x', 'G', 'gamma', 'beta', 'stats', 'act', 'p_drop', 'keepbits')

  def __init__(self, x, G, gamma, beta, stats, act, p_drop=0., keepbits=None):
    self.x, self.G, self.gamma, self.beta, self.stats, self.act = x, G, gamma, beta, stats, act
    self.p_drop, self.keepbits = p_drop, keepbits


def dz_applicable(x, x2, mask, p_drop, keepbits):
  """Shapes / modes the dz epilogue covers (everything else keeps the two-phase GroupNorm backward)."""
  if not GN_DZ or x2 is not None or mask is not None or x.dtype != torch.bfloat16 or gemm_backend == 'simt':
    return False
  B, H, W, C = x.shape
  hw = H * W
  if p_drop > 0. and keepbits is None:
    return False
  return hw >= 32 and (hw & (hw - 1)) == 0 and C % 128 == 0 and C <= 1024 and (B * hw) % 256 == 0


def _dz_request(kw, req, out):
  """Adds the dz request to a (data-gradient) convolution call; returns a closure that yields the quad sums the
  epilogue emitted (fp32 (M/32, N/4, 2)) or None when the launch could not take the request (out is then plain dy)."""
  B, H, W, C = req.x.shape
  M, N = kw['M'], kw['N']
  assert N == C and M == B * H * W
  cst = torch.empty((B, C, 4), dtype=torch.float32, device=out.device)
  check(lib.st_gn_bwd_consts(ptr(req.gamma), ptr(req.beta), ptr(req.stats[0]), ptr(req.stats[1]), B, C, req.G, ptr(cst),
                             stream()))
  part = torch.empty((M // 32, N // 4, 2), dtype=torch.float32, device=out.device)
  got = ctypes.c_int32(0)
  kw.update(gn_part=part, gn_hw=H * W, gn_rows_out=ctypes.pointer(got), dz_x=req.x, dz_ldx=C, dz_cst=cst,
            dz_keep=req.keepbits if req.p_drop > 0. else None, dz_inv_keep=1.0 / (1.0 - req.p_drop), dz_act=int(req.act))
  return lambda: part if got.value == 32 else None


def conv_fwd(x, w, cout, kh=3, kw=3, x2=None, bias=None, rowbias=None, rowbias_ld=0, residual=None, alpha=1.0,
             out_dtype=None, out=None, want_quads=False, dz=None):
  """'same' convolution of NHWC x (optionally channel-concatenated with x2) with packed weights
  w[cout][kh*kw][cin] (contiguous, same dtype as x).  rowbias: fp32 [B][rowbias_ld] slice added per image.
  want_quads: returns (out, Quads | None) - the GroupNorm partial sums of `out` emitted by the epilogue.
  dz (DzRequest; data gradients run as forward convolutions): returns (out, qpart | None) - with qpart, `out` holds
  dz of the GroupNorm whose output this convolution's forward consumed and gn_backward_dz finishes the job."""
  B, H, W, C1 = x.shape
  C2 = 0 if x2 is None else x2.shape[3]
  K = kh * kw * (C1 + C2)
  od = out_dtype or x.dtype
  if out is None:
    out = torch.empty((B, H, W, cout), dtype=od, device=x.device)
  kwargs = dict(a_mode=OP_GATHER, b_mode=OP_STRIDED, in_dtype=dt(x), out_dtype=_DT[od], M=B * H * W, N=cout, K=K,
                A=x, A2=x2, B=w, C=out, sBn=K, sBk=1, sCm=cout, n_img=B, H=H, W=W, C1=C1, C2=C2, kh=kh, kw=kw,
                bias=bias, rowbias=rowbias, rows_per_rb=H * W, ld_rb=rowbias_ld, residual=residual, sRm=cout,
                alpha=alpha)
  q = _quads_request(kwargs, out, H * W) if want_quads else None
  if dz is not None:
    q = _dz_request(kwargs, dz, out)
  _gemm(**kwargs)
  return (out, q()) if (want_quads or dz is not None) else out


def conv_dgrad(dy, w, cin, kh=3, kw=3, alpha=1.0, out=None):
  """Data gradient of conv_fwd: dy (B,H,W,cout) x w[cout][taps][cin] -> (B,H,W,cin)."""
  B, H, W, Co = dy.shape
  if out is None:
    out = torch.empty((B, H, W, cin), dtype=dy.dtype, device=dy.device)
  _gemm(a_mode=OP_GATHER, b_mode=OP_DGRADW, in_dtype=dt(dy), out_dtype=dt(out), M=B * H * W, N=cin,
        K=kh * kw * Co, A=dy, B=w, C=out, sCm=cin, n_img=B, H=H, W=W, C1=Co, C2=0, kh=kh, kw=kw, alpha=alpha)
  return out


def conv_wgrad(dy, x, dw, kh=3, kw=3, x2=None, alpha=1.0):
  """dw[cout][taps][cin] (fp32, contiguous view into the flat gradient buffer) += alpha * dy^T * im2col(x)."""
  if not PARAM_GRADS:
    return
  B, H, W, Co = dy.shape
  C1 = x.shape[3]
  C2 = 0 if x2 is None else x2.shape[3]
  N = kh * kw * (C1 + C2)
  K = B * H * W
  _gemm(a_mode=OP_STRIDED, b_mode=OP_GATHER, in_dtype=dt(dy), out_dtype=F32, M=Co, N=N, K=K, A=dy, B=x, B2=x2,
        C=dw, sAm=1, sAk=Co, sCm=N, n_img=B, H=H, W=W, C1=C1, C2=C2, kh=kh, kw=kw, accumulate=1,
        split_k=_split_k(Co, N, K), alpha=alpha)


def gemm_nt(a, b, out=None, out_dtype=None, bias=None, residual=None, alpha=1.0, lda=None, ldb=None, ldc=None,
            M=None, N=None, K=None, batch=1, sAb=0, sBb=0, sCb=0, sRb=0, ldr=None, quads_hw=0):
  """C[m][n] = alpha*(sum_k a[m][k] b[n][k] + bias[n] + residual[m][n]); a, b row-major with leading
  dimensions lda/ldb (K contiguous).  quads_hw > 0 (rows are pixels of images of quads_hw pixels): returns
  (C, Quads | None) with the GroupNorm partial sums of C emitted by the epilogue."""
  M = M if M is not None else a.shape[-2]
  K = K if K is not None else a.shape[-1]
  N = N if N is not None else b.shape[-2]
  od = out_dtype or a.dtype
  if out is None:
    out = torch.empty((M, N) if batch == 1 else (batch, M, N), dtype=od, device=a.device)
  ldc = ldc or N
  kwargs = dict(a_mode=OP_STRIDED, b_mode=OP_STRIDED, in_dtype=dt(a), out_dtype=dt(out), M=M, N=N, K=K, batch=batch, A=a,
                B=b, C=out, sAm=lda or K, sAk=1, sAb=sAb, sBn=ldb or K, sBk=1, sBb=sBb, sCm=ldc, sCb=sCb, bias=bias,
                residual=residual, sRm=ldr or ldc, sRb=sRb, alpha=alpha)
  q = _quads_request(kwargs, out, quads_hw) if (quads_hw and batch == 1 and ldc == N) else None
  _gemm(**kwargs)
  if quads_hw:
    return out, (q() if q is not None else None)
  return out


def gemm_nn(a, b, N, out=None, out_dtype=None, alpha=1.0, lda=None, ldb=None, ldc=None, M=None, K=None, batch=1,
            sAb=0, sBb=0, sCb=0):
  """C[m][n] = alpha * sum_k a[m][k] b[k][n]; b row-major [K][ldb] (N contiguous)."""
  M = M if M is not None else a.shape[-2]
  K = K if K is not None else a.shape[-1]
  od = out_dtype or a.dtype
  if out is None:
    out = torch.empty((M, N) if batch == 1 else (batch, M, N), dtype=od, device=a.device)
  _gemm(a_mode=OP_STRIDED, b_mode=OP_STRIDED, in_dtype=dt(a), out_dtype=dt(out), M=M, N=N, K=K, batch=batch, A=a,
        B=b, C=out, sAm=lda or K, sAk=1, sAb=sAb, sBn=1, sBk=ldb or N, sBb=sBb, sCm=ldc or N, sCb=sCb, alpha=alpha)
  return out


def gemm_tn(a, b, M, N, K, out=None, out_dtype=None, alpha=1.0, lda=None, ldb=None, ldc=None, batch=1, sAb=0,
            sBb=0, sCb=0, accumulate=False, split_k=None):
  """C[m][n] (+)= alpha * sum_k a[k][m] b[k][n]; a row-major [K][lda], b row-major [K][ldb]."""
  od = out_dtype or a.dtype
  if out is None:
    out = torch.empty((M, N) if batch == 1 else (batch, M, N), dtype=od, device=a.device)
  if split_k is None:
    split_k = _split_k(M, N, K) if accumulate else 1
  _gemm(a_mode=OP_STRIDED, b_mode=OP_STRIDED, in_dtype=dt(a), out_dtype=dt(out), M=M, N=N, K=K, batch=batch, A=a,
        B=b, C=out, sAm=1, sAk=lda or M, sAb=sAb, sBn=1, sBk=ldb or N, sBb=sBb, sCm=ldc or N, sCb=sCb,
        accumulate=int(accumulate), split_k=split_k, alpha=alpha)
  return out


# Backward passes normally produce parameter AND input gradients.  `input_grads_only()` switches the weight / bias /
# GroupNorm-parameter gradient work off for passes that only want d(out)/d(input): the Hutchinson divergence of the
# likelihood ODE (reference likelihood.py:27-37) differentiates the score with respect to x with the weights fixed.
PARAM_GRADS = True


class input_grads_only:
  def __enter__(self):
    global PARAM_GRADS
    self.prev, PARAM_GRADS = PARAM_GRADS, False

  def __exit__(self, *exc):
    global PARAM_GRADS
    PARAM_GRADS = self.prev


# ------------------------------------------------------------------------------------ GroupNorm
# resident forward for the in-kernel-dropout passes: measured level with / behind the pipelined pair (94 vs 85 us at
# 32x32x128 even with the generator moved under the loads: it is ALU-bound), so off by default
_GN_FWD_FUSED_DROP = os.environ.get('ST_GN_FWD_FUSED_DROP', '0') != '0'
_GN_SPLIT_W = int(os.environ.get('ST_GN_SPLIT_W', '4'))    # tuning knob: blocks per SM the split reductions aim for


def _gn_splits(n_img, hw):
  s = max(1, min((148 * _GN_SPLIT_W + n_img - 1) // n_img, hw // 64, 64))
  return int(s)


class GnStats:
  """mean / rstd of one GroupNorm call: `t` is the (2, B, G) tensor; while `part` is set the statistics still live
  as partial sums and the next gn_apply finalises them inside its own kernel (no st_gn_finalize launch)."""
  __slots__ = ('t', 'part', 'splits', 'count', 'eps')

  def __init__(self, t, part=None, splits=0, count=0, eps=0.):
    self.t, self.part, self.splits, self.count, self.eps = t, part, splits, count, eps

  def __getitem__(self, i):
    assert self.part is None, 'statistics not finalised yet (run gn_apply first)'
    return self.t[i]


def gn_stats(x, x2, G, eps=1e-6, finalize=True):
  B, H, W, C1 = x.shape
  C2 = 0 if x2 is None else x2.shape[3]
  hw = H * W
  splits = _gn_splits(B, hw)
  part = torch.empty((B, splits, G, 2), dtype=torch.float32, device=x.device)
  stats = torch.empty((2, B, G), dtype=torch.float32, device=x.device)
  check(lib.st_gn_stats(ptr(x), ptr(x2), dt(x), B, hw, C1, C2, G, splits, ptr(part), stream()))
  count = hw * ((C1 + C2) // G)
  if not finalize:
    return GnStats(stats, part, splits, count, eps)
  check(lib.st_gn_finalize(ptr(part), B, splits, G, count, eps, ptr(stats[0]), ptr(stats[1]), stream()))
  return GnStats(stats)


def gn_apply(x, x2, G, gamma, beta, stats, act, p_drop=0., seed=0, mask=None, keepbits=None, quads=None, eps=1e-6):
  """`keepbits`: optional uint8 tensor (B*H*W*C/8 bytes) that receives the dropout keep flags drawn by the kernel.
  `quads` = (Quads of x, Quads of x2 | None): the statistics are finalised inside the kernel from the sums the
  producing GEMMs emitted (`stats` is then the (2, B, G) tensor that receives mean / rstd)."""
  B, H, W, C1 = x.shape
  C2 = 0 if x2 is None else x2.shape[3]
  y = torch.empty((B, H, W, C1 + C2), dtype=x.dtype, device=x.device)
  pending = isinstance(stats, GnStats) and stats.part is not None
  t = stats.t if isinstance(stats, GnStats) else stats
  q1 = q2 = None
  qrows, count = 0, stats.count if pending else 0
  if quads is not None:
    q1, q2, qrows = quads[0].t, (quads[1].t if x2 is not None else None), quads[0].rows
    count = H * W * ((C1 + C2) // G)
  check(lib.st_gn_apply(ptr(x), ptr(x2), dt(x), B, H * W, C1, C2, G, ptr(gamma), ptr(beta), ptr(t[0]),
                        ptr(t[1]), int(act), float(p_drop), int(seed), ptr(mask), ptr(keepbits), ptr(y),
                        ptr(stats.part) if pending else None, stats.splits if pending else 0,
                        count, float(stats.eps) if pending else float(eps), ptr(q1), ptr(q2), int(qrows), stream()))
  if pending:
    stats.part = None          # finalised by the kernel
  return y


def gn_norm_act(x, x2, G, gamma, beta, act, p_drop=0., seed=0, mask=None, keepbits=None, eps=1e-6, fused_chunks=None,
                quads=None):
  """GroupNorm (+SiLU, +dropout) forward: returns (y, GnStats).  `quads` = (Quads | None of x, of x2): when every source
  carries the partial sums its producing GEMM emitted, ONE streaming apply launch (1 read + 1 write at copy speed, no
  statistics pass).  Otherwise one cluster launch with the image resident in shared memory when it fits
  (st_gn_fwd_fused), else st_gn_stats followed by st_gn_apply (which finalises the statistics)."""
  B, H, W, C1 = x.shape
  C2 = 0 if x2 is None else x2.shape[3]
  if (quads is not None and fused_chunks is None and quads[0] is not None and (x2 is None or quads[1] is not None)
      and (x2 is None or quads[1].rows == quads[0].rows) and (H * W) % quads[0].rows == 0):
    stats = torch.empty((2, B, G), dtype=torch.float32, device=x.device)
    y = gn_apply(x, x2, G, gamma, beta, stats, act, p_drop=p_drop, seed=seed, mask=mask, keepbits=keepbits,
                 quads=quads, eps=eps)
    return y, GnStats(stats)
  if fused_chunks is not None:
    fc = int(fused_chunks)
  elif p_drop > 0. and mask is None and not _GN_FWD_FUSED_DROP:
    fc = 0      # in-kernel dropout passes keep the pipelined stats + apply pair (ST_GN_FWD_FUSED_DROP=1 to fuse)
  else:
    fc = lib.st_gn_fwd_fused_chunks(B, H * W, C1 + C2)
  if fc <= 0:
    st = gn_stats(x, x2, G, eps=eps, finalize=False)
    return gn_apply(x, x2, G, gamma, beta, st, act, p_drop=p_drop, seed=seed, mask=mask, keepbits=keepbits), st
  y = torch.empty((B, H, W, C1 + C2), dtype=x.dtype, device=x.device)
  stats = torch.empty((2, B, G), dtype=torch.float32, device=x.device)
  check(lib.st_gn_fwd_fused(ptr(x), ptr(x2), dt(x), B, H * W, C1, C2, G, ptr(gamma), ptr(beta), float(eps), int(act),
                            float(p_drop), int(seed), ptr(mask), ptr(keepbits), ptr(y), ptr(stats[0]), ptr(stats[1]),
                            fc, stream()))
  return y, GnStats(stats)


_WAVE_WORK = {}


def _wave_work(device, n_img):
  """Zeroed int32 counters of st_gn_bwd_wave (the kernel hands them back zeroed), one buffer per device."""
  w = _WAVE_WORK.get(device)
  if w is None or w.numel() < n_img + 2:
    w = torch.zeros(max(n_img + 2, 1024), dtype=torch.int32, device=device)
    _WAVE_WORK[device] = w
  return w


def gn_backward(x, x2, dy, G, gamma, beta, stats, act, dgamma, dbeta, p_drop=0., seed=0, mask=None, extra=None,
                extra_scale=1.0, dx1=None, accum1=False, dx2=None, accum2=False, keepbits=None, want_csum=False,
                queue=None, fused_chunks=None, resident=None, wave=None):
  """Returns (dx1, dx2[, csum]); accumulates dgamma/dbeta (fp32 views into the flat gradient buffer).
  csum (want_csum): fp32 (B, chunks, C1+C2) column sums of the gradient this call contributed.
  `queue`: a ColsumQueue that takes the dgamma/dbeta reduction of the fused (single-launch) form; `fused_chunks`:
  cluster size override (0 = two-kernel form, None = the library decides); `resident`: cluster size of the form that
  keeps x / dy in shared memory between the two phases (2-stream calls only; None = the library decides, 0 = off);
  `wave`: the persistent two-phase form (st_gn_bwd_wave; None = whenever no other form is forced and the tensor is
  large enough to gain, False = off)."""
  B, H, W, C1 = x.shape
  C2 = 0 if x2 is None else x2.shape[3]
  hw, Ct = H * W, C1 + C2
  if dx1 is None:
    dx1 = torch.empty_like(x)
    accum1 = False
  if x2 is not None and dx2 is None:
    dx2 = torch.empty_like(x2)
    accum2 = False
  if not PARAM_GRADS:
    dgamma = dbeta = None
  head = (ptr(x), ptr(x2), ptr(dy), dt(x), B, hw, C1, C2, G, ptr(gamma), ptr(beta), ptr(stats[0]), ptr(stats[1]),
          int(act), float(p_drop), int(seed), ptr(mask), ptr(keepbits))
  n_streams = 2 + (extra is not None) + bool(accum1 or accum2)
  if wave is None:
    wave = fused_chunks is None and resident is None
  if wave:
    # one persistent launch, reduction items of the next group of images overlapping the apply items of this one
    group = ctypes.c_int32(0)
    wc = lib.st_gn_bwd_wave_plan(B, hw, Ct, dt(x), ctypes.byref(group))
    if wc > 0:
      red = torch.empty((B, wc, Ct, 2), dtype=torch.float32, device=x.device)
      csum = torch.empty((B, wc, Ct), dtype=torch.float32, device=x.device) if want_csum else None
      check(lib.st_gn_bwd_wave(*head, int(wc), int(group.value), ptr(red), ptr(extra), float(extra_scale), ptr(dx1), int(accum1),
                               ptr(dx2), int(accum2), ptr(csum), ptr(_wave_work(x.device, B)), 0, stream()))
      if dgamma is None:
        pass
      elif queue is not None:
        queue.add_gn_params(dgamma, dbeta, red.view(B * wc, Ct, 2))
      else:
        check(lib.st_gn_bwd_params(ptr(red), B * wc, Ct, ptr(dgamma), ptr(dbeta), stream()))
      return (dx1, dx2, csum) if want_csum else (dx1, dx2)
  if resident is None:
    ok = fused_chunks is None and not (accum1 or accum2)
    resident = lib.st_gn_bwd_resident_chunks(B, hw, Ct, n_streams) if ok else 0
  if resident > 0:
    # x and dy read once: the cluster keeps the image in shared memory between the reduction and the apply phase
    assert not (accum1 or accum2), 'the resident backward form takes no accumulated destination'
    red = torch.empty((B, resident, Ct, 2), dtype=torch.float32, device=x.device)
    csum = torch.empty((B, resident, Ct), dtype=torch.float32, device=x.device) if want_csum else None
    check(lib.st_gn_bwd_resident(*head, int(resident), ptr(red), ptr(extra), float(extra_scale), ptr(dx1), ptr(dx2),
                                 ptr(csum), stream()))
    if dgamma is None:
      pass
    elif queue is not None:
      queue.add_gn_params(dgamma, dbeta, red.view(B * resident, Ct, 2))
    else:
      check(lib.st_gn_bwd_params(ptr(red), B * resident, Ct, ptr(dgamma), ptr(dbeta), stream()))
    return (dx1, dx2, csum) if want_csum else (dx1, dx2)
  fc = lib.st_gn_bwd_fused_chunks(B, hw, Ct, dt(x), n_streams) if fused_chunks is None else int(fused_chunks)
  if fc > 0:
    # one launch: a cluster of `fc` CTAs per image reduces, synchronises and applies; the parameter gradients are
    # column sums of `red` over all images (deferred to the batched reduction when a queue is given)
    red = torch.empty((B, fc, Ct, 2), dtype=torch.float32, device=x.device)
    csum = torch.empty((B, fc, Ct), dtype=torch.float32, device=x.device) if want_csum else None
    check(lib.st_gn_bwd_fused(*head, fc, ptr(red), ptr(extra), float(extra_scale), ptr(dx1), int(accum1), ptr(dx2),
                              int(accum2), ptr(csum), stream()))
    if dgamma is None:
      pass
    elif queue is not None:
      queue.add_gn_params(dgamma, dbeta, red.view(B * fc, Ct, 2))
    else:
      check(lib.st_gn_bwd_params(ptr(red), B * fc, Ct, ptr(dgamma), ptr(dbeta), stream()))
    return (dx1, dx2, csum) if want_csum else (dx1, dx2)
  splits = _gn_splits(B, hw)
  red = torch.empty((B, splits, Ct, 2), dtype=torch.float32, device=x.device)
  common = head + (splits, ptr(red))
  check(lib.st_gn_bwd_reduce(*common, stream()))
  chunks, csum = 0, None
  if want_csum:
    chunks = lib.st_gn_chunks(B, hw, Ct)
    csum = torch.empty((B, chunks, Ct), dtype=torch.float32, device=x.device)
  check(lib.st_gn_bwd_apply(*common, ptr(extra), float(extra_scale), ptr(dx1), int(accum1), ptr(dx2), int(accum2),
                            chunks, ptr(csum), ptr(dgamma), ptr(dbeta), stream()))
  return (dx1, dx2, csum) if want_csum else (dx1, dx2)


def gn_backward_dz(x, x2, dz, qpart, G, gamma, stats, dgamma, dbeta, extra=None, extra_scale=1.0, dx1=None,
                   accum1=False, dx2=None, accum2=False, want_csum=False, queue=None):
  """The GroupNorm backward that is left when the data-gradient GEMM already produced dz and its quad sums
  (conv_fwd(dz=...)): one streaming pass.  Same returns / parameter-gradient handling as gn_backward."""
  B, H, W, C1 = x.shape
  C2 = 0 if x2 is None else x2.shape[3]
  hw, Ct = H * W, C1 + C2
  if dx1 is None:
    dx1 = torch.empty_like(x)
    accum1 = False
  if x2 is not None and dx2 is None:
    dx2 = torch.empty_like(x2)
    accum2 = False
  if not PARAM_GRADS:
    dgamma = dbeta = None
  chunks = lib.st_gn_chunks(B, hw, Ct)
  red = torch.empty((B, chunks, Ct, 2), dtype=torch.float32, device=x.device) if dgamma is not None else None
  csum = torch.empty((B, chunks, Ct), dtype=torch.float32, device=x.device) if want_csum else None
  check(lib.st_gn_bwd_dz_apply(ptr(x), ptr(x2), ptr(dz), dt(x), B, hw, C1, C2, G, ptr(gamma), ptr(stats[0]), ptr(stats[1]),
                               ptr(qpart), ptr(extra), float(extra_scale), ptr(dx1), int(accum1), ptr(dx2), int(accum2),
                               chunks, ptr(red), ptr(csum), stream()))
  if dgamma is None:
    pass
  elif queue is not None:
    queue.add_gn_params(dgamma, dbeta, red.view(B * chunks, Ct, 2))
  else:
    check(lib.st_gn_bwd_params(ptr(red), B * chunks, Ct, ptr(dgamma), ptr(dbeta), stream()))
  return (dx1, dx2, csum) if want_csum else (dx1, dx2)


# ------------------------------------------------------------------------------------ elementwise
def cast(src, dtype, out=None):
  if out is None:
    out = torch.empty(src.shape, dtype=dtype, device=src.device)
  check(lib.st_cast(ptr(src), dt(src), ptr(out), dt(out), src.numel(), stream()))
  return out


def axpby(a, b=None, alpha=1.0, beta=1.0, out=None):
  if out is None:
    out = torch.empty_like(a)
  check(lib.st_axpby(ptr(a), ptr(b), ptr(out), dt(a), float(alpha), float(beta), a.numel(), stream()))
  return out


def silu(x):
  y = torch.empty_like(x)
  check(lib.st_silu(ptr(x), ptr(y), dt(x), x.numel(), stream()))
  return y


def silu_bwd(x, dy):
  dx = torch.empty_like(x)
  check(lib.st_silu_bwd(ptr(x), ptr(dy), ptr(dx), dt(x), x.numel(), stream()))
  return dx


def resample2x(x, x2, direction, scale):
  """direction=+1: nearest replicate x2 (times scale); -1: 2x2 box sum (times scale)."""
  B, H, W, C1 = x.shape
  C2 = 0 if x2 is None else x2.shape[3]
  oh, ow = (2 * H, 2 * W) if direction > 0 else (H // 2, W // 2)
  y = torch.empty((B, oh, ow, C1 + C2), dtype=x.dtype, device=x.device)
  check(lib.st_resample2x(ptr(x), ptr(x2), ptr(y), dt(x), B, H, W, C1, C2, int(direction), float(scale), stream()))
  return y


def colsum(x, groups, rows_per_group, C, out, scale=1.0, accumulate=False, ld=None):
  """out[g][c] (+)= scale * sum of x[g*rows_per_group + r][c] (rows `ld` elements apart).  A single long group
  (bias gradients over every pixel) is reduced in two deterministic passes so that the first one fills the GPU."""
  ld = ld or C
  if groups == 1 and rows_per_group >= 4096 and ld == C:
    chunks = 1
    while chunks < 256 and rows_per_group % (chunks * 2) == 0 and rows_per_group // (chunks * 2) >= 32:
      chunks *= 2
    if chunks > 1:
      part = torch.empty((chunks, C), dtype=torch.float32, device=x.device)
      check(lib.st_colsum(ptr(x), dt(x), chunks, rows_per_group // chunks, C, C, 1.0, ptr(part), 0, stream()))
      check(lib.st_colsum(ptr(part), F32, 1, chunks, C, C, float(scale), ptr(out), int(accumulate), stream()))
      return out
  check(lib.st_colsum(ptr(x), dt(x), groups, rows_per_group, C, ld, float(scale), ptr(out), int(accumulate), stream()))
  return out


class ColsumQueue:
  """Small fp32 column sums deferred to ONE st_colsum_batched launch (bias and time-embedding gradients built from the
  column-sum partials of the GroupNorm backward kernels).  Destinations of queued jobs must be distinct."""

  def __init__(self):
    self.jobs, self.keep, self.blocks_y = [], [], 1

  def _check(self, t):
    assert t.dtype == torch.float32 and t.is_cuda and t.stride(-1) == 1 and t.data_ptr() % 16 == 0

  def add_reduce(self, dst, parts, scale=1.0, accumulate=True):
    """dst[c] (+)= scale * sum_p sum_rows parts[p][row][c]; parts: 2-D fp32 views (row stride arbitrary)."""
    C = dst.numel()
    for i in range(0, len(parts), 4):
      chunk = parts[i:i + 4]
      rec = [0] * 21
      for k, p in enumerate(chunk):
        self._check(p)
        assert p.shape[1] == C and p.stride(0) % 4 == 0
        rec[k], rec[4 + k], rec[8 + k] = p.data_ptr(), p.shape[0], p.stride(0)
      rec[12], rec[13], rec[14], rec[15] = dst.data_ptr(), len(chunk), 0, C
      rec[19] = np.float64(scale).view(np.int64).item()
      rec[20] = int(accumulate or i > 0)
      if i > 0:                       # same destination twice: keep stream order by flushing the first record
        self.flush()
      self.jobs.append(rec)
      self.keep.extend(chunk)
      self.keep.append(dst)
      self.blocks_y = max(self.blocks_y, (C + 127) // 128)

  def add_groups(self, dst, part, scale=1.0, accumulate=False):
    """dst[g][c] (+)= scale * sum_k part[g][k][c]; part contiguous fp32 (G, K, C), dst a (G, C) view (row stride free)."""
    self._check(part)
    self._check(dst)
    G, K, C = part.shape
    assert part.is_contiguous() and dst.shape == (G, C) and dst.stride(0) % 4 == 0 and C % 4 == 0
    rec = [0] * 21
    rec[0], rec[4], rec[8] = part.data_ptr(), G * K, C
    rec[12], rec[13], rec[14], rec[15] = dst.data_ptr(), 1, 1, C
    rec[16], rec[17], rec[18] = G, K, dst.stride(0)
    rec[19] = np.float64(scale).view(np.int64).item()
    rec[20] = int(accumulate)
    self.jobs.append(rec)
    self.keep.extend((part, dst))
    self.blocks_y = max(self.blocks_y, (G * (C // 4) + 1023) // 1024)

  def add_gn_params(self, dgamma, dbeta, red):
    """dgamma[c] += sum_rows red[row][c][1]; dbeta[c] += sum_rows red[row][c][0]; red contiguous fp32 (rows, C, 2)."""
    self._check(red)
    rows, C, two = red.shape
    assert two == 2 and red.is_contiguous() and C % 2 == 0 and dgamma.numel() == C and dbeta.numel() == C
    assert dgamma.dtype == torch.float32 and dbeta.dtype == torch.float32 and dgamma.is_contiguous() and dbeta.is_contiguous()
    rec = [0] * 21
    rec[0], rec[1], rec[4], rec[8] = red.data_ptr(), dbeta.data_ptr(), rows, 2 * C
    rec[12], rec[13], rec[14], rec[15] = dgamma.data_ptr(), 1, 2, C
    rec[20] = 1
    self.jobs.append(rec)
    self.keep.extend((red, dgamma, dbeta))
    self.blocks_y = max(self.blocks_y, (2 * C + 127) // 128)

  # Pinned staging buffers for job tables built while a CUDA graph is being captured: pinned memory must not be
  # allocated under capture, and the captured host->device copy re-reads the buffer at every replay, so the buffers
  # are reserved before the capture starts and handed to the graph's owner afterwards (losses._StepGraph).
  PINNED_ROWS = 4096
  _pinned_free, _pinned_used = [], []

  @classmethod
  def reserve_pinned(cls, n):
    while len(cls._pinned_free) < n:
      cls._pinned_free.append(torch.zeros((cls.PINNED_ROWS, 21), dtype=torch.int64).pin_memory())

  @classmethod
  def take_pinned(cls):
    used, cls._pinned_used = cls._pinned_used, []
    return used

  def flush(self):
    if not self.jobs:
      return
    dev = self.keep[0].device
    host = torch.tensor(self.jobs, dtype=torch.int64)
    if torch.cuda.is_current_stream_capturing():
      assert ColsumQueue._pinned_free and len(self.jobs) <= self.PINNED_ROWS, 'reserve_pinned() before capturing'
      buf = ColsumQueue._pinned_free.pop()
      ColsumQueue._pinned_used.append(buf)
      buf[:len(self.jobs)].copy_(host)
      table = torch.empty((len(self.jobs), 21), dtype=torch.int64, device=dev)
      table.copy_(buf[:len(self.jobs)], non_blocking=True)
    else:
      table = host.pin_memory().to(dev, non_blocking=True)
    check(lib.st_colsum_batched(ptr(table), len(self.jobs), self.blocks_y, stream()))
    self.jobs, self.keep, self.blocks_y = [], [], 1


def softmax_fwd(logits, L, scale, dtype):
  rows = logits.numel() // L
  p = torch.empty(logits.shape, dtype=dtype, device=logits.device)
  check(lib.st_softmax_fwd(ptr(logits), ptr(p), dt(p), rows, L, float(scale), stream()))
  return p


def softmax_bwd(p, dp, L, scale):
  rows = p.numel() // L
  ds = torch.empty_like(p)
  check(lib.st_softmax_bwd(ptr(p), ptr(dp), ptr(ds), dt(p), rows, L, float(scale), stream()))
  return ds


_ATTN_FUSED = os.environ.get('ST_ATTN_FUSED', '1') != '0'


def attn_fused_ok(L, C, dtype):
  """True when the one-kernel attention core (st_attn_fwd) runs this shape."""
  return _ATTN_FUSED and dtype == torch.bfloat16 and bool(lib.st_attn_fwd_supported(int(L), int(C), BF16))


def attn_fwd(qkv, B, L, C, scale, save_p):
  """o = softmax(q k^T * scale) v from the packed (B*L, 3C) projections; returns (o (B, L, C), p (B, L, L) or None)."""
  o = torch.empty((B, L, C), dtype=qkv.dtype, device=qkv.device)
  p = torch.empty((B, L, L), dtype=qkv.dtype, device=qkv.device) if save_p else None
  check(lib.st_attn_fwd(ptr(qkv), ptr(o), ptr(p), B, L, C, float(scale), stream()))
  return o, p


_FREQS = {}


def timestep_frequencies(dim, device, max_positions=10000):
  """exp(-log(max_positions)/(half-1) * arange(half)) exactly as the reference evaluates it
  (models/layers.py:518-521: float32 tensor times a Python double, torch.exp on the host)."""
  key = (dim, str(device), max_positions)
  if key not in _FREQS:
    half = dim // 2
    emb = math.log(max_positions) / (half - 1)
    _FREQS[key] = torch.exp(torch.arange(half, dtype=torch.float32) * -emb).to(device)
  return _FREQS[key]


def timestep_embedding(labels, dim, max_positions=10000):
  out = torch.empty((labels.shape[0], dim), dtype=torch.float32, device=labels.device)
  check(lib.st_timestep_embedding(ptr(labels), ptr(timestep_frequencies(dim, labels.device, max_positions)), ptr(out),
                                  labels.shape[0], dim, stream()))
  return out


def fourier_embedding(sigma, W):
  out = torch.empty((sigma.shape[0], 2 * W.numel()), dtype=torch.float32, device=sigma.device)
  check(lib.st_fourier_embedding(ptr(sigma), ptr(W), ptr(out), sigma.shape[0], W.numel(), stream()))
  return out


def nchw_to_nhwc(x, dtype, cpad=None, alpha=1.0, beta=0.0):
  """NCHW fp32 -> NHWC `dtype` with the channel axis zero-padded to `cpad`; y = alpha*x + beta."""
  B, C, H, W = x.shape
  cpad = cpad or C
  y = torch.empty((B, H, W, cpad), dtype=dtype, device=x.device)
  check(lib.st_nchw_to_nhwc(ptr(x), ptr(y), dt(y), B, C, H, W, cpad, float(alpha), float(beta), stream()))
  return y


def nhwc_to_nchw(x, C=None, row_scale=None):
  """First C channels of NHWC x -> NCHW fp32 (optionally times row_scale[n])."""
  B, H, W, cpad = x.shape
  C = C or cpad
  y = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
  check(lib.st_nhwc_to_nchw(ptr(x), dt(x), ptr(y), B, C, H, W, cpad, ptr(row_scale), stream()))
  return y


def im2col_small(x, kh, kw, kpad):
  B, H, W, C = x.shape
  out = torch.empty((B * H * W, kpad), dtype=torch.bfloat16, device=x.device)
  check(lib.st_im2col_small(ptr(x), dt(x), ptr(out), B, H, W, C, kh, kw, kpad, stream()))
  return out


def upfirdn2d_nhwc(x, k, up=1, down=1, pad=(0, 0)):
  """x (major, H, W, minor) -> FIR-resampled tensor; k fp32 (kh, kw) on the device."""
  mj, H, W, mn = x.shape
  kh, kw = k.shape
  oh = (H * up + pad[0] + pad[1] - kh + down) // down
  ow = (W * up + pad[0] + pad[1] - kw + down) // down
  y = torch.empty((mj, oh, ow, mn), dtype=x.dtype, device=x.device)
  check(lib.st_upfirdn2d(ptr(x), ptr(y), dt(x), ptr(k), mj, H, W, mn, kh, kw, up, up, down, down, pad[0], pad[1],
                         pad[0], pad[1], stream()))
  return y


def im2col(x, kh, kw, stride, pad, oh, ow):
  """NHWC x -> (B*oh*ow, kh*kw*C) patch matrix of a strided convolution (zero outside the image)."""
  B, H, W, C = x.shape
  cols = torch.empty((B * oh * ow, kh * kw * C), dtype=x.dtype, device=x.device)
  check(lib.st_im2col(ptr(x), ptr(cols), dt(x), B, H, W, C, kh, kw, stride, pad, oh, ow, stream()))
  return cols


def col2im(dcols, shape, kh, kw, stride, pad, oh, ow):
  """Adjoint of im2col: (B*oh*ow, kh*kw*C) -> NHWC gradient of shape `shape`."""
  B, H, W, C = shape
  dx = torch.empty(shape, dtype=dcols.dtype, device=dcols.device)
  check(lib.st_col2im(ptr(dcols), ptr(dx), dt(dcols), B, H, W, C, kh, kw, stride, pad, oh, ow, stream()))
  return dx
