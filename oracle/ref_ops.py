"""ORACLE (test infrastructure, never shipped or timed as the product).

CPU restatements of the two native ops the reference ships:

* upfirdn2d      - reference op/upfirdn2d_kernel.cu:107-207 (index arithmetic :182-203),
                   CPU fallback op/upfirdn2d.py:159-200.
* fused_bias_act - reference op/fused_bias_act_kernel.cu:18-49.

Parity pin: tests/golden/ops_golden.npz holds outputs of the reference's own
`upfirdn2d_native` and of `F.leaky_relu`-based `fused_leaky_relu` generated in this
container by tests/golden/make_golden.py; tests/test_oracle.py checks these
restatements against them.
"""
import numpy as np


def upfirdn2d_out_size(n, up, down, pad0, pad1, k):
  # op/upfirdn2d_kernel.cu:237-240
  return (n * up + pad0 + pad1 - k + down) // down


def upfirdn2d_ref(x, k, up=(1, 1), down=(1, 1), pad=(0, 0, 0, 0)):
  """x: [major, in_h, in_w] float array, k: [kh, kw]. pad = (x0, x1, y0, y1).

  Direct evaluation of the reference kernel's per-output-pixel formula
  (op/upfirdn2d_kernel.cu:182-203): zero-insert upsample, pad/crop, true convolution
  with k (the kernel is flipped on load, :137), stride-`down` decimation.
  """
  x = np.asarray(x)
  k = np.asarray(k, dtype=x.dtype)
  up_x, up_y = up
  down_x, down_y = down
  px0, px1, py0, py1 = pad
  major, in_h, in_w = x.shape
  kh, kw = k.shape
  out_h = upfirdn2d_out_size(in_h, up_y, down_y, py0, py1, kh)
  out_w = upfirdn2d_out_size(in_w, up_x, down_x, px0, px1, kw)
  kf = k[::-1, ::-1]
  out = np.zeros((major, out_h, out_w), dtype=np.float64)
  for oy in range(out_h):
    mid_y = oy * down_y + up_y - 1 - py0
    in_y = mid_y // up_y            # floor division, also for negatives
    ky0 = (in_y + 1) * up_y - mid_y - 1
    for ox in range(out_w):
      mid_x = ox * down_x + up_x - 1 - px0
      in_x = mid_x // up_x
      kx0 = (in_x + 1) * up_x - mid_x - 1
      acc = np.zeros(major, dtype=np.float64)
      for y in range(kh // up_y):
        iy = in_y + y
        if iy < 0 or iy >= in_h:
          continue
        for xx in range(kw // up_x):
          ix = in_x + xx
          if ix < 0 or ix >= in_w:
            continue
          acc += x[:, iy, ix].astype(np.float64) * float(kf[ky0 + y * up_y, kx0 + xx * up_x])
      out[:, oy, ox] = acc
  return out.astype(x.dtype)


def upfirdn2d_backward_args(in_h, in_w, out_h, out_w, kh, kw, up, down, pad):
  """up/down/pad of the op that maps grad_out -> grad_in (op/upfirdn2d.py:101-116);
  it runs on flip(k)."""
  up_x, up_y = up
  down_x, down_y = down
  px0, px1, py0, py1 = pad
  g_px0 = kw - px0 - 1
  g_py0 = kh - py0 - 1
  g_px1 = in_w * up_x - out_w * down_x + px0 - up_x + 1
  g_py1 = in_h * up_y - out_h * down_y + py0 - up_y + 1
  return (down_x, down_y), (up_x, up_y), (g_px0, g_px1, g_py0, g_py1)


def fused_bias_act_ref(x, bias, ref, act, grad, alpha, scale, step_b=None):
  """y = act(x + b[(i / step_b) % size_b]) * scale (op/fused_bias_act_kernel.cu:18-49).

  act: 1 linear, 3 leaky-relu(alpha); grad: 0 forward, 1 first derivative gated on the
  sign of `ref`, 2 -> zeros.  x is [N, C, ...]; step_b is the stride of the channel
  axis (prod of trailing dims), as the reference's launcher computes it (:66-71).
  """
  x = np.asarray(x)
  flat = x.reshape(-1).astype(np.float64)
  if bias is not None and np.size(bias):
    if step_b is None:
      step_b = int(np.prod(x.shape[2:])) if x.ndim > 2 else 1
    idx = (np.arange(flat.size) // step_b) % np.size(bias)
    flat = flat + np.asarray(bias, dtype=np.float64)[idx]
  r = np.asarray(ref).reshape(-1).astype(np.float64) if ref is not None and np.size(ref) else np.zeros_like(flat)
  if act == 3:
    if grad == 0:
      y = np.where(flat > 0, flat, flat * alpha)
    elif grad == 1:
      y = np.where(r > 0, flat, flat * alpha)
    else:
      y = np.zeros_like(flat)
  else:
    y = flat if grad < 2 else np.zeros_like(flat)
  return (y * scale).astype(x.dtype).reshape(x.shape)
