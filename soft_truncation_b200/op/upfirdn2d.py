"""upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)) on NCHW tensors, same signature and autograd
behaviour as reference op/upfirdn2d.py:19-156 (forward, backward and double backward all run the same
kernel with swapped factors / flipped FIR), executed by st_upfirdn2d."""
import torch
from torch.autograd import Function

from .. import ops
from .._lib import check, lib


def _run(x, kernel, up, down, pad):
  """x: (major, in_h, in_w, minor) contiguous CUDA tensor; up/down/pad = (x, y) pairs / (x0, x1, y0, y1)."""
  if not x.is_cuda:
    raise RuntimeError('upfirdn2d (B200 build) needs a CUDA tensor')
  if x.dtype not in (torch.float32, torch.bfloat16):
    raise RuntimeError(f'upfirdn2d: unsupported dtype {x.dtype}')
  major, in_h, in_w, minor = x.shape
  kh, kw = kernel.shape
  out_h = (in_h * up[1] + pad[2] + pad[3] - kh + down[1]) // down[1]
  out_w = (in_w * up[0] + pad[0] + pad[1] - kw + down[0]) // down[0]
  y = torch.empty((major, out_h, out_w, minor), dtype=x.dtype, device=x.device)
  k = kernel.to(device=x.device, dtype=torch.float32).contiguous()
  check(lib.st_upfirdn2d(ops.ptr(x), ops.ptr(y), ops.dt(x), ops.ptr(k), major, in_h, in_w, minor, kh, kw, up[0], up[1],
                         down[0], down[1], pad[0], pad[1], pad[2], pad[3], ops.stream()))
  return y


class UpFirDn2dBackward(Function):

  @staticmethod
  def forward(ctx, grad_output, kernel, grad_kernel, up, down, pad, g_pad, in_size, out_size):
    up_x, up_y = up
    down_x, down_y = down
    grad_output = grad_output.reshape(-1, out_size[0], out_size[1], 1).contiguous()
    grad_input = _run(grad_output, grad_kernel, (down_x, down_y), (up_x, up_y), g_pad)
    grad_input = grad_input.view(in_size[0], in_size[1], in_size[2], in_size[3])
    ctx.save_for_backward(kernel)
    ctx.up, ctx.down, ctx.pad, ctx.in_size, ctx.out_size = up, down, pad, in_size, out_size
    return grad_input

  @staticmethod
  def backward(ctx, gradgrad_input):
    kernel, = ctx.saved_tensors
    gradgrad_input = gradgrad_input.reshape(-1, ctx.in_size[2], ctx.in_size[3], 1).contiguous()
    gradgrad_out = _run(gradgrad_input, kernel, ctx.up, ctx.down, ctx.pad)
    gradgrad_out = gradgrad_out.view(ctx.in_size[0], ctx.in_size[1], ctx.out_size[0], ctx.out_size[1])
    return gradgrad_out, None, None, None, None, None, None, None, None


class UpFirDn2d(Function):

  @staticmethod
  def forward(ctx, input, kernel, up, down, pad):
    up_x, up_y = up
    down_x, down_y = down
    pad_x0, pad_x1, pad_y0, pad_y1 = pad
    kernel_h, kernel_w = kernel.shape
    batch, channel, in_h, in_w = input.shape
    ctx.in_size = input.shape
    input = input.reshape(-1, in_h, in_w, 1).contiguous()
    ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]))
    out = _run(input, kernel, up, down, pad)
    out_h, out_w = out.shape[1], out.shape[2]
    ctx.out_size = (out_h, out_w)
    ctx.up, ctx.down, ctx.pad = (up_x, up_y), (down_x, down_y), (pad_x0, pad_x1, pad_y0, pad_y1)
    g_pad_x0 = kernel_w - pad_x0 - 1
    g_pad_y0 = kernel_h - pad_y0 - 1
    g_pad_x1 = in_w * up_x - out_w * down_x + pad_x0 - up_x + 1
    g_pad_y1 = in_h * up_y - out_h * down_y + pad_y0 - up_y + 1
    ctx.g_pad = (g_pad_x0, g_pad_x1, g_pad_y0, g_pad_y1)
    return out.view(-1, channel, out_h, out_w)

  @staticmethod
  def backward(ctx, grad_output):
    kernel, grad_kernel = ctx.saved_tensors
    grad_input = UpFirDn2dBackward.apply(grad_output, kernel, grad_kernel, ctx.up, ctx.down, ctx.pad, ctx.g_pad,
                                         ctx.in_size, ctx.out_size)
    return grad_input, None, None, None, None


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
  """Same call as the reference (op/upfirdn2d.py:145-156); CUDA tensors only in this build."""
  return UpFirDn2d.apply(input, kernel, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
