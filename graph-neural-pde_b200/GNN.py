"""Model shell without early stopping (reference src/GNN.py): encoder Linear -> ODE block -> ReLU -> decoder Linear.
The encoder / decoder are ordinary nn.Linear calls either side of the hot path."""
import torch
import torch.nn.functional as F

from .base_classes import BaseGNN
from .model_configurations import set_block, set_function


class GNN(BaseGNN):
  def __init__(self, opt, dataset, device=torch.device('cpu')):
    super(GNN, self).__init__(opt, dataset, device)
    self.f = set_function(opt)
    block = set_block(opt)
    time_tensor = torch.tensor([0, self.T]).to(device)
    self.odeblock = block(self.f, self.regularization_fns, opt, dataset.data, device, t=time_tensor).to(device)

  def encode(self, x, pos_encoding=None):
    """reference src/GNN.py:19-48"""
    opt = self.opt
    if opt['use_labels']:
      y = x[:, -self.num_classes:]
      x = x[:, :-self.num_classes]
    if opt['beltrami']:
      x = self.mx(F.dropout(x, opt['input_dropout'], training=self.training))
      p = self.mp(F.dropout(pos_encoding, opt['input_dropout'], training=self.training))
      x = torch.cat([x, p], dim=1)
    else:
      x = self.m1(F.dropout(x, opt['input_dropout'], training=self.training))
    if opt['use_mlp']:
      x = F.dropout(x, opt['dropout'], training=self.training)
      x = F.dropout(x + self.m11(F.relu(x)), opt['dropout'], training=self.training)
      x = F.dropout(x + self.m12(F.relu(x)), opt['dropout'], training=self.training)
    if opt['use_labels']:
      x = torch.cat([x, y], dim=-1)
    if opt['batch_norm']:
      x = self.bn_in(x)
    if opt['augment']:
      x = torch.cat([x, torch.zeros(x.shape).to(self.device)], dim=1)
    return x

  def decode(self, z, width):
    """reference src/GNN.py:57-72"""
    if self.opt['augment']:
      z = torch.split(z, width // 2, dim=1)[0]
    z = F.relu(z)
    if self.opt['fc_out']:
      z = F.relu(self.fc(z))
    z = F.dropout(z, self.opt['dropout'], training=self.training)
    return self.m2(z)

  def forward(self, x, pos_encoding=None):
    x = self.encode(x, pos_encoding)
    self.odeblock.set_x0(x)
    z = self.odeblock(x)
    return self.decode(z, x.shape[1])
