import os
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(__file__))


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu on the GPU box')


def pytest_collection_modifyitems(config, items):
  if torch.cuda.is_available():
    return
  skip = pytest.mark.skip(reason='no CUDA device')
  for item in items:
    if 'gpu' in item.keywords:
      item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
  cases = {}
  for name in ('golden_v1.pt', 'golden_v2.pt', 'golden_v3.pt', 'golden_v4.pt'):   # v2: GAT + hard attention, v3: model shell, v4: GAT mix_features
    part = torch.load(os.path.join(ROOT, 'tests', 'golden', name), map_location='cpu', weights_only=False)
    assert not set(part) & set(cases)
    cases.update(part)
  return cases
