"""Eval-time integrators with early stopping (reference src/early_stop_solver.py): integrate to earlystopxT * T, score
the classifier after every solver step and remember the best validation step.  The attribute chain
`odeblock.test_integrator.solver.best_{val,test,train,time}` read by run_GNN.main (reference src/run_GNN.py:266-271)
is kept.  The ODE steps run on the CUDA engine; the per-step classifier read-out (relu -> m2 -> argmax -> masked
accuracy -> "is this the best validation step so far") stays on the device: the reference's three `.item()` host reads
per solver step (src/early_stop_solver.py:203-218) become ONE read of four numbers when the integration ends
(SURVEY.md 8(f).2).  Accuracies are formed as integer counts divided in float64, i.e. the same values python's
int / int gives the reference."""
import torch
import torch.nn.functional as F

from .odeint import integrate_dopri5, integrate_fixed


def run_evaluator(evaluator, data, y_pred):
  return tuple(evaluator.eval({'y_true': data.y[m], 'y_pred': y_pred[m]})['acc']
               for m in (data.train_mask, data.val_mask, data.test_mask))


class AccuracyEvaluator:
  """Stand-in for ogb.nodeproppred.Evaluator(name='ogbn-arxiv').eval -> {'acc': mean(y_true == y_pred)}."""

  def __init__(self, name=None):
    self.name = name

  def eval(self, d):
    y_true, y_pred = d['y_true'], d['y_pred']
    return {'acc': float((y_true.reshape(-1) == y_pred.reshape(-1)).sum().item()) / max(1, y_true.shape[0])}


class _EarlyStopSolver:
  def __init__(self, func, y0, rtol, atol, opt, **options):
    self.func, self.y0, self.rtol, self.atol, self.opt, self.options = func, y0, rtol, atol, opt, options
    self.m2_weight = None
    self.m2_bias = None
    self.data = None
    self.best_val = 0
    self.best_test = 0
    self.best_train = 0
    self.best_time = 0
    self.dataset = opt['dataset']
    self.max_test_steps = opt['max_test_steps']
    self.evaluator = AccuracyEvaluator(opt['dataset']) if opt['dataset'] == 'ogbn-arxiv' else None

  def set_accs(self, train, val, test, time):
    self.best_train, self.best_val, self.best_test = train, val, test
    self.best_time = float(time)

  @torch.no_grad()
  def test(self, logits):
    accs = []
    for mask in (self.data.train_mask, self.data.val_mask, self.data.test_mask):
      pred = logits[mask].max(1)[1]
      accs.append(pred.eq(self.data.y[mask]).sum().item() / mask.sum().item())
    return accs

  @torch.no_grad()
  def test_OGB(self, logits):
    y_pred = logits.argmax(dim=-1, keepdim=True)
    return list(run_evaluator(self.evaluator, self.data, y_pred))

  @torch.no_grad()
  def evaluate(self, z):
    if not self.m2_weight.shape[1] == z.shape[1]:  # system has been augmented
      z = torch.split(z, self.m2_weight.shape[1], dim=1)[0]
    z = F.linear(F.relu(z), self.m2_weight, self.m2_bias)
    if self.dataset == 'ogbn-arxiv':
      return self.test_OGB(z.log_softmax(dim=-1))
    return self.test(z)

  # ---- device-resident scoring ------------------------------------------------------------------------------------
  def _masks(self, device):
    if getattr(self, '_mask_cache', None) is None:
      d = self.data
      masks = torch.stack([m.to(device).bool() for m in (d.train_mask, d.val_mask, d.test_mask)])
      self._mask_cache = (masks, masks.sum(dim=1).double(), d.y.to(device).reshape(-1))
      self._best = torch.zeros(4, dtype=torch.float64, device=device)   # train, val, test, time
    return self._mask_cache

  @torch.no_grad()
  def _score(self, y1, t1):
    """evaluate() + the `val_acc > best_val` update of the reference (:203-218, :172-176) without host reads."""
    if not self.m2_weight.shape[1] == y1.shape[1]:  # system has been augmented
      y1 = torch.split(y1, self.m2_weight.shape[1], dim=1)[0]
    masks, counts, labels = self._masks(y1.device)
    pred = F.linear(F.relu(y1), self.m2_weight, self.m2_bias).argmax(dim=-1)   # log_softmax keeps the argmax
    hits = (pred.reshape(-1) == labels).unsqueeze(0) & masks
    accs = hits.sum(dim=1).double() / counts
    cand = torch.cat([accs, torch.full((1,), float(t1), dtype=torch.float64, device=y1.device)])
    self._best = torch.where(accs[1] > self._best[1], cand, self._best)

  def _publish(self):
    """the single host read: best_{train,val,test,time} as python floats, as run_GNN.main expects them"""
    if getattr(self, '_best', None) is not None:
      train, val, test, time = self._best.tolist()
      if val > self.best_val:
        self.set_accs(train, val, test, time)


class EarlyStopRK4(_EarlyStopSolver):
  """reference src/early_stop_solver.py:131-225"""
  order = 4

  def integrate(self, t):
    solution = integrate_fixed(self.func, self.y0, t, 'rk4', self.options['step_size'],
                               on_step=lambda t0, t1, y0, y1: self._score(y1, t1))
    self._publish()
    return t[-1], solution


class EarlyStopDopri5(_EarlyStopSolver):
  """reference src/early_stop_solver.py:30-129; `advance` stops after max_test_steps attempted steps"""
  order = 5

  def integrate(self, t):
    def hook(solver, accepted):
      self._score(solver.ycur, solver.t1)   # the reference scores after every attempt; a reject re-scores y0

    solution, _ = integrate_dopri5(self.func, self.y0, t, self.rtol, self.atol, on_step=hook,
                                   max_steps=self.max_test_steps)
    self._publish()
    return t[-1], solution


SOLVERS = {'dopri5': EarlyStopDopri5, 'rk4': EarlyStopRK4}


class EarlyStopInt(torch.nn.Module):
  """odeint-compatible callable (reference src/early_stop_solver.py:234-308): ignores the `t` it is given and
  integrates over [0, earlystopxT * T]; a fresh solver per call, kept as `self.solver`."""

  def __init__(self, t, opt, device=None):
    super(EarlyStopInt, self).__init__()
    self.device = device
    self.solver = None
    self.data = None
    self.max_test_steps = opt['max_test_steps']
    self.m2_weight = None
    self.m2_bias = None
    self.opt = opt
    self.t = torch.tensor([0, opt['earlystopxT'] * t], dtype=torch.float).to(self.device)

  def __call__(self, func, y0, t, method=None, rtol=1e-7, atol=1e-9, adjoint_method="dopri5", adjoint_atol=1e-9,
               adjoint_rtol=1e-7, options=None):
    method = self.opt['method']
    assert method in ['rk4', 'dopri5'], "Only dopri5 and rk4 implemented with early stopping"
    self.solver = SOLVERS[method](func, y0, rtol=rtol, atol=atol, opt=self.opt, **dict(options or {}))
    if self.solver.data is None:
      self.solver.data = self.data
    self.solver.m2_weight = self.m2_weight
    self.solver.m2_bias = self.m2_bias
    t, solution = self.solver.integrate(self.t)
    return solution
