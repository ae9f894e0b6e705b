"""Dense layers of the model shell on the native library: `NativeLinear` is a drop-in for the `nn.Linear` layers the
reference's BaseGNN creates (m1, m11, m12, mx, mp, fc, m2 -- reference src/base_classes.py:110-127, used at
src/GNN.py:23-31 and 60-71): same constructor, parameter names (`weight`, `bias`), initialisation and state_dict keys.

  forward   y = x W^T + b          grand_linear: tcgen05 / TMEM / TMA kernel (csrc/project_tc.cu) when the contraction
                                   is short (d_in <= 128), exact fp32 tiles otherwise
  backward  dx = dy W              grand_linear on the transposed weight
            dW = dy^T x, db = 1^T dy   the two reductions over the node dimension stay cuBLAS GEMMs (torch.matmul)

There is no CPU path: CPU tensors raise GrandError, as everywhere in this package.
"""
import torch
from torch import nn

from . import _native as N


def linear_forward(x, weight, bias):
  """x [n, d_in] (CUDA fp32) -> [n, d_out] through grand_linear."""
  N.require_cuda(x, 'x')
  N.require_cuda(weight, 'weight')
  if x.dim() != 2 or x.dtype != torch.float32 or weight.dtype != torch.float32:
    raise N.GrandError('grand_linear: expected a 2-D fp32 input and fp32 weights, got %s %s' % (tuple(x.shape), x.dtype))
  x = x.contiguous()
  weight = weight.contiguous()
  n, d_in = x.shape
  d_out = weight.shape[0]
  if weight.shape[1] != d_in:
    raise N.GrandError('grand_linear: input width %d != weight width %d' % (d_in, weight.shape[1]))
  y = torch.empty((n, d_out), dtype=torch.float32, device=x.device)
  b = bias.contiguous() if bias is not None else None
  N.check(N.lib().grand_linear(x.data_ptr(), n, d_in, weight.data_ptr(), N.ptr(b), d_out, y.data_ptr(), d_out,
                               N.current_stream(x.device)))
  return y


class LinearFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x, weight, bias):
    ctx.save_for_backward(x, weight)
    ctx.has_bias = bias is not None
    return linear_forward(x, weight, bias)

  @staticmethod
  def backward(ctx, g):
    x, weight = ctx.saved_tensors
    g = g.contiguous()
    gx = gw = gb = None
    if ctx.needs_input_grad[0]:
      gx = linear_forward(g, weight.detach().t().contiguous(), None)
    if ctx.needs_input_grad[1]:
      gw = g.t() @ x
    if ctx.has_bias and ctx.needs_input_grad[2]:
      gb = g.sum(0)
    return gx, gw, gb


class NativeLinear(nn.Linear):
  """nn.Linear whose forward / input gradient run on the native library."""

  def forward(self, x):
    return LinearFn.apply(x, self.weight, self.bias)
