"""Model shells around the ODE block: `GNN` (reference src/GNN.py) and `GNNEarly` (reference src/GNN_early.py, the
default model of run_GNN.py).  Both are   encoder -> ODE block -> ReLU -> decoder;   they differ only in the eval-time
integrator.  The shells are callers of the hot path, not part of it: the encoder and
decoder are ordinary nn.Linear layers created by BaseGNN under the reference's names, so reference checkpoints load.

Written as two small stages (`embed`, `readout`) instead of the reference's two copies of one long forward.
"""
import torch
import torch.nn.functional as F

from .base_classes import BaseGNN
from .early_stop_solver import EarlyStopInt
from .model_configurations import set_block, set_function


class GNN(BaseGNN):
  """No early stopping: train and eval integrate over [0, T] with the block's own integrators."""

  def __init__(self, opt, dataset, device=torch.device('cpu')):
    super(GNN, self).__init__(opt, dataset, device)
    self.f = set_function(opt)
    span = torch.tensor([0, self.T]).to(device)
    self.odeblock = set_block(opt)(self.f, self.regularization_fns, opt, dataset.data, device, t=span).to(device)

  # -- stage 1: node features (+ positional encodings / labels) -> initial ODE state -------------------------------
  def embed(self, x, pos_encoding=None):
    opt, drop = self.opt, lambda t, p: F.dropout(t, p, training=self.training)
    labels = None
    if opt['use_labels']:
      x, labels = x[:, :-self.num_classes], x[:, -self.num_classes:]
    if opt['beltrami']:      # BLEND: joint feature + positional state (reference src/GNN.py:23-28)
      x = torch.cat([self.mx(drop(x, opt['input_dropout'])), self.mp(drop(pos_encoding, opt['input_dropout']))], dim=1)
    else:
      x = self.m1(drop(x, opt['input_dropout']))
    if opt['use_mlp']:
      x = drop(x, opt['dropout'])
      for layer in (self.m11, self.m12):
        x = drop(x + layer(F.relu(x)), opt['dropout'])
    if labels is not None:
      x = torch.cat([x, labels], dim=-1)
    if opt['batch_norm']:
      x = self.bn_in(x)
    if opt['augment']:
      x = torch.cat([x, torch.zeros(x.shape).to(self.device)], dim=1)
    return x

  # -- stage 3: final ODE state -> class logits ---------------------------------------------------------------------
  def readout(self, z, state_width):
    if self.opt['augment']:
      z = torch.split(z, state_width // 2, dim=1)[0]
    z = F.relu(z)
    if self.opt['fc_out']:
      z = F.relu(self.fc(z))
    return self.m2(F.dropout(z, self.opt['dropout'], training=self.training))

  def before_block(self):
    pass

  def forward(self, x, pos_encoding=None):
    x = self.embed(x, pos_encoding)
    self.odeblock.set_x0(x)
    self.before_block()
    return self.readout(self.odeblock(x), x.shape[1])


class GNNEarly(GNN):
  """Eval passes integrate to earlystopxT * T with EarlyStopInt, which scores the decoder after every solver step;
  it therefore needs the current decoder weights before every forward (reference src/GNN_early.py:27-33, 78-79)."""

  def __init__(self, opt, dataset, device=torch.device('cpu')):
    super(GNNEarly, self).__init__(opt, dataset, device)
    with torch.no_grad():
      self.odeblock.test_integrator = EarlyStopInt(self.T, self.opt, self.device)
      self.set_solver_data(dataset.data)

  def set_solver_m2(self):
    integ = self.odeblock.test_integrator
    integ.m2_weight = self.m2.weight.data.detach().clone().to(self.device)
    integ.m2_bias = self.m2.bias.data.detach().clone().to(self.device)

  def set_solver_data(self, data):
    self.odeblock.test_integrator.data = data

  def before_block(self):
    with torch.no_grad():
      self.set_solver_m2()

  def forward_encoder(self, x, pos_encoding):
    return self.embed(x, pos_encoding)

  def forward_ODE(self, x, pos_encoding):
    x = self.embed(x, pos_encoding)
    self.odeblock.set_x0(x)
    z = self.odeblock(x)
    return torch.split(z, x.shape[1] // 2, dim=1)[0] if self.opt['augment'] else z
