"""Differentiable restatement of SpGraphTransAttentionLayer's head-mean attention (reference
src/function_transformer_attention.py:128-214, src/utils.py:179-208) in torch ops ON THE DEVICE -- used for ONE thing:
the backward of the attention FROZEN at x(0) (AttODEblock, reference src/block_transformer_attention.py:28-39) in the
configurations the fused backward kernel does not cover (cosine / pearson / exp-kernel scorers, squareplus,
attention_norm_idx = 1, reweighting: what the reference's tuned Cora / Citeseer configurations use).  That backward
happens once per training step, not once per RHS evaluation: the forward of the attention stays the native kernel, and
so does every evaluation of the Laplacian function and its per-edge weight gradient (grand_edge_dot); here the
scores are recomputed from x(0) with autograd enabled and differentiated -- a dozen edge-sized torch kernels per step,
in the same class as the cuBLAS reductions of the dense tails.  CPU tensors are refused like everywhere else.
"""
import torch

from . import _native as N


def _heads(t, h):
  return t.view(t.shape[0], h, -1).transpose(1, 2)                    # [N, d_k, H]  (:176-188)


def _cosine(a, b, eps=1e-5):
  return torch.nn.functional.cosine_similarity(a, b, dim=1, eps=eps)  # torch.nn.CosineSimilarity(dim=1, eps=1e-5) (:95)


def _segment_softmax(src, index, n):
  """torch_geometric.utils.softmax: exp(s - max_seg) / (sum_seg + 1e-16), the maximum taken without gradient."""
  idx = index.unsqueeze(1).expand_as(src)
  m = torch.full((n, src.shape[1]), float('-inf'), dtype=src.dtype, device=src.device)
  m = m.scatter_reduce(0, idx, src.detach(), 'amax', include_self=True)
  out = (src - m[index]).exp()
  s = torch.zeros((n, src.shape[1]), dtype=src.dtype, device=src.device).index_add_(0, index, out)
  return out / (s[index] + 1e-16)


def _squareplus(src, index, n):
  """reference src/utils.py:179-208: GLOBAL max subtraction (with gradient, as autograd sees it there)."""
  out = src - src.max()
  out = (out + torch.sqrt(out ** 2 + 4)) / 2
  s = torch.zeros((n, src.shape[1]), dtype=src.dtype, device=src.device).index_add_(0, index, out)
  return out / (s[index] + 1e-16)


def mean_attention(x, edge, layer, num_nodes):
  """mean over heads of SpGraphTransAttentionLayer.forward(x, edge)[0], [E] in the order of `edge`, differentiable with
  respect to x and the layer's parameters."""
  N.require_cuda(x, 'x')
  opt = layer.opt
  if not hasattr(layer, 'Q'):
    raise NotImplementedError('grand_b200 backward: the beltrami exp_kernel layer (Qx / Qp) is forward-only')
  h = int(opt['heads'])
  q = _heads(torch.nn.functional.linear(x, layer.Q.weight, layer.Q.bias), h)
  k = _heads(torch.nn.functional.linear(x, layer.K.weight, layer.K.bias), h)
  src, dst = q[edge[0]], k[edge[1]]                                   # :190-191
  kind = opt['attention_type']
  if kind == 'scaled_dot':
    prods = torch.sum(src * dst, dim=1) / (src.shape[1] ** 0.5)      # :195-196
  elif kind == 'cosine_sim':
    prods = _cosine(src, dst)                                         # :197-199
  elif kind == 'pearson':
    prods = _cosine(src - src.mean(dim=1, keepdim=True), dst - dst.mean(dim=1, keepdim=True))   # :200-206
  elif kind == 'exp_kernel':
    prods = layer.output_var ** 2 * torch.exp(-torch.sum((src - dst) ** 2, dim=1) / (2 * layer.lengthscale ** 2))   # :193-194
  else:
    raise NotImplementedError('attention_type %r' % kind)
  if opt['reweight_attention'] and layer.edge_weights is not None:    # :208-209
    prods = prods * layer.edge_weights.to(prods).unsqueeze(1)
  index = edge[int(opt['attention_norm_idx'])]
  att = _squareplus(prods, index, num_nodes) if opt['square_plus'] else _segment_softmax(prods, index, num_nodes)
  return att.mean(dim=1)


def transformer_rhs(x, func, x0):
  """ODEFuncTransformerAtt.forward (reference src/function_transformer_attention.py:38-53) in torch ops, differentiable:
  f = alpha (A(x) x - x) + beta x0 with A the head-mean attention.  Materialises [E, d] (as the reference's spmm backward
  does): meant for the graphs the unfused configurations are used on, not for ogbn-arxiv."""
  layer, edge = func.multihead_att_layer, func.edge_index
  mean = mean_attention(x, edge, layer, x.shape[0])
  ax = torch.zeros_like(x).index_add_(0, edge[0], x[edge[1]] * mean.unsqueeze(1))     # torch_sparse.spmm (:35)
  alpha = func.alpha_train if func.opt['no_alpha_sigmoid'] else torch.sigmoid(func.alpha_train)
  f = alpha * (ax - x)
  if func.opt['add_source']:
    f = f + func.beta_train * x0
  return f


def gat_rhs(x, func, x0):
  """ODEFuncAtt.forward (reference src/function_GAT_attention.py:31-64, layer :105-116) in torch ops, differentiable:
  additive scorer LeakyReLU(a . [h_i ; h_j]) per head, segment softmax, then -- by linearity of the per-head spmm's and
  their mean -- ONE aggregation with the head-mean attention of x (or of x W followed by Wout with mix_features)."""
  N.require_cuda(x, 'x')
  layer, edge, opt = func.multihead_att_layer, func.edge_index, func.opt
  h, dk, n = layer.h, layer.d_k, x.shape[0]
  wx = torch.mm(x, layer.W)                                                           # :106
  hh = wx.view(-1, h, dk).transpose(1, 2)                                             # [N, d_k, H]  :107-108
  edge_h = torch.cat((hh[edge[0]], hh[edge[1]]), dim=1).transpose(0, 1)               # [2 d_k, E, H]  :111-112
  scores = torch.nn.functional.leaky_relu(torch.sum(layer.a * edge_h, dim=0), layer.alpha)   # :113
  mean = _segment_softmax(scores, edge[int(opt['attention_norm_idx'])], n).mean(dim=1)        # :114
  if opt['mix_features']:                                                             # :32-37
    agg = torch.zeros_like(wx).index_add_(0, edge[0], wx[edge[1]] * mean.unsqueeze(1))
    ax = torch.mm(agg, layer.Wout)
  else:                                                                               # :38-42
    ax = torch.zeros_like(x).index_add_(0, edge[0], x[edge[1]] * mean.unsqueeze(1))
  alpha = func.alpha_train if opt['no_alpha_sigmoid'] else torch.sigmoid(func.alpha_train)
  f = alpha * (ax - x)
  if opt['add_source']:
    f = f + func.beta_train * x0
  return f
