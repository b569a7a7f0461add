"""fused_leaky_relu / FusedLeakyReLU with the reference's signatures (op/fused_act.py:20-97), executed by
st_fused_bias_act: y = leaky_relu(x + bias[c], slope) * scale, with the same first/second-order backward."""
import torch
from torch import nn
from torch.autograd import Function

from .. import ops
from .._lib import check, lib


def _bias_act(x, bias, ref, act, grad, alpha, scale):
  if not x.is_cuda:
    raise RuntimeError('fused_leaky_relu (B200 build) needs a CUDA tensor')
  x = x.contiguous()
  y = torch.empty_like(x)
  size_b = bias.numel() if bias is not None else 0
  step_b = 1
  for d in x.shape[2:]:
    step_b *= d
  check(lib.st_fused_bias_act(ops.ptr(x), ops.ptr(bias.to(x.dtype).contiguous()) if bias is not None else None,
                              ops.ptr(ref.contiguous()) if ref is not None else None, ops.ptr(y), ops.dt(x), x.numel(),
                              size_b, step_b, act, grad, float(alpha), float(scale), ops.stream()))
  return y


class FusedLeakyReLUFunctionBackward(Function):

  @staticmethod
  def forward(ctx, grad_output, out, negative_slope, scale):
    ctx.save_for_backward(out)
    ctx.negative_slope, ctx.scale = negative_slope, scale
    grad_input = _bias_act(grad_output, None, out, 3, 1, negative_slope, scale)
    dim = [0]
    if grad_input.ndim > 2:
      dim += list(range(2, grad_input.ndim))
    grad_bias = grad_input.sum(dim).detach()
    return grad_input, grad_bias

  @staticmethod
  def backward(ctx, gradgrad_input, gradgrad_bias):
    out, = ctx.saved_tensors
    gradgrad_out = _bias_act(gradgrad_input, gradgrad_bias, out, 3, 1, ctx.negative_slope, ctx.scale)
    return gradgrad_out, None, None, None


class FusedLeakyReLUFunction(Function):

  @staticmethod
  def forward(ctx, input, bias, negative_slope, scale):
    out = _bias_act(input, bias, None, 3, 0, negative_slope, scale)
    ctx.save_for_backward(out)
    ctx.negative_slope, ctx.scale = negative_slope, scale
    return out

  @staticmethod
  def backward(ctx, grad_output):
    out, = ctx.saved_tensors
    grad_input, grad_bias = FusedLeakyReLUFunctionBackward.apply(grad_output, out, ctx.negative_slope, ctx.scale)
    return grad_input, grad_bias, None, None


class FusedLeakyReLU(nn.Module):

  def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
    super().__init__()
    self.bias = nn.Parameter(torch.zeros(channel))
    self.negative_slope = negative_slope
    self.scale = scale

  def forward(self, input):
    return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
  return FusedLeakyReLUFunction.apply(input, bias, negative_slope, scale)
