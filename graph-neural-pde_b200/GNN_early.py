"""Default model shell (reference src/GNN_early.py): like GNN, but the block's test integrator is the early-stopping
EarlyStopInt, which needs the decoder weights copied in before every forward (reference :31-33, :78-79)."""
import torch

from .GNN import GNN
from .early_stop_solver import EarlyStopInt


class GNNEarly(GNN):
  def __init__(self, opt, dataset, device=torch.device('cpu')):
    super(GNNEarly, self).__init__(opt, dataset, device)
    with torch.no_grad():
      self.odeblock.test_integrator = EarlyStopInt(self.T, self.opt, self.device)
      self.set_solver_data(dataset.data)

  def set_solver_m2(self):
    self.odeblock.test_integrator.m2_weight = self.m2.weight.data.detach().clone().to(self.device)
    self.odeblock.test_integrator.m2_bias = self.m2.bias.data.detach().clone().to(self.device)

  def set_solver_data(self, data):
    self.odeblock.test_integrator.data = data

  def forward(self, x, pos_encoding=None):
    x = self.encode(x, pos_encoding)
    self.odeblock.set_x0(x)
    with torch.no_grad():
      self.set_solver_m2()
    z = self.odeblock(x)
    return self.decode(z, x.shape[1])

  def forward_encoder(self, x, pos_encoding):
    return self.encode(x, pos_encoding)

  def forward_ODE(self, x, pos_encoding):
    x = self.encode(x, pos_encoding)
    self.odeblock.set_x0(x)
    z = self.odeblock(x)
    if self.opt['augment']:
      z = torch.split(z, x.shape[1] // 2, dim=1)[0]
    return z
