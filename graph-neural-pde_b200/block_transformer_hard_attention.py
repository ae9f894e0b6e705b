"""HardAttODEblock -- reference src/block_transformer_hard_attention.py:7-104 (GRAND-nl-rw / "hard attention"), one of
the "next" rows of SURVEY.md 8(f).4.

Every forward scores the FULL graph once at x(0).  In training the block keeps only the edges whose head-mean attention
(optionally times the feature "flux" |x_i - x_j|) exceeds the (1 - att_samp_pct) quantile, renormalises the kept
weights per node and hands the smaller graph to the function; in eval mode every edge stays with its head-mean
attention.  Scores, quantile, mask and renormalisation all stay on the device: the only host synchronisation is the
edge count the new graph plan needs (the reference has the same one in its boolean-mask indexing)."""
import torch

from .base_classes import ODEblock
from .function_transformer_attention import SpGraphTransAttentionLayer
from .utils import get_rw_adj


class HardAttODEblock(ODEblock):
  def __init__(self, odefunc, regularization_fns, opt, data, device, t=torch.tensor([0, 1]), gamma=0.5):
    super(HardAttODEblock, self).__init__(odefunc, regularization_fns, opt, data, device, t)
    assert opt['att_samp_pct'] > 0 and opt['att_samp_pct'] <= 1, "attention sampling threshold must be in (0,1]"
    self.rebuild_function(odefunc, data)
    self.num_nodes = data.num_nodes
    self.install_graph(*get_rw_adj(data.edge_index, edge_weight=data.edge_attr, norm_dim=1,
                                   fill_value=opt['self_loop_weight'], num_nodes=data.num_nodes, dtype=data.x.dtype))
    self.data_edge_index = self.odefunc.edge_index          # the function's edge_index is replaced every forward
    self.use_default_integrators()
    if opt['function'] not in {'GAT', 'transformer'}:        # the Laplacian function has no scorer of its own
      width = opt['hidden_dim']
      self.multihead_att_layer = SpGraphTransAttentionLayer(width, width, opt, device,
                                                            edge_weights=self.odefunc.edge_weight).to(device)

  def get_attention_weights(self, x):
    own = self.opt['function'] not in {'GAT', 'transformer'}
    layer = self.multihead_att_layer if own else self.odefunc.multihead_att_layer
    attention, _ = layer(x, self.data_edge_index)
    return attention

  def renormalise_attention(self, attention):
    index = self.odefunc.edge_index[self.opt['attention_norm_idx']]
    sums = torch.zeros(self.num_nodes, dtype=attention.dtype, device=attention.device).index_add_(0, index, attention)
    return attention / (sums[index] + 1e-16)

  def _share(self, edge_index, weights):
    for f in (self.odefunc, self.reg_odefunc.odefunc):
      if f.edge_index is not edge_index:     # assigning drops the function's cached graph plan: only when it changed
        f.edge_index = edge_index
      f.attention_weights = weights
    self.reg_odefunc.odefunc.edge_weight = self.odefunc.edge_weight

  @torch.no_grad()
  def select_edges(self, x):
    """-> (edge_index, attention_weights[E']) as the reference's forward computes them (:49-68)"""
    mean_att = self.get_attention_weights(x).mean(dim=1, keepdim=False)
    if not self.training:
      return self.data_edge_index, mean_att
    if self.opt['use_flux']:
      src, dst = self.data_edge_index[0, :], self.data_edge_index[1, :]
      mean_att = mean_att * torch.linalg.norm(x[src, :] - x[dst, :], dim=1)
    mask = mean_att > torch.quantile(mean_att, 1 - self.opt['att_samp_pct'])
    self.odefunc.edge_index = self.data_edge_index[:, mask]
    return self.odefunc.edge_index, self.renormalise_attention(mean_att[mask])

  def forward(self, x):
    self._share(*self.select_edges(x))
    o = self.opt
    return self.integrate(x, solver_options={'step_size': o['step_size']},
                          adjoint_options=lambda: {'step_size': o['adjoint_step_size']})
