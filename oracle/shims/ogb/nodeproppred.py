class Evaluator:
  def __init__(self, name=None):
    self.name = name

  def eval(self, d):
    y_true, y_pred = d['y_true'], d['y_pred']
    return {'acc': float((y_true == y_pred).sum().item()) / max(1, y_true.shape[0])}


class PygNodePropPredDataset:
  def __init__(self, *a, **k):
    raise RuntimeError('no datasets in this image')
