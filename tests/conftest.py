import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
  config.addinivalue_line('markers', 'slow: long-running CPU test')


def pytest_collection_modifyitems(config, items):
  import torch
  if torch.cuda.is_available():
    return
  skip = pytest.mark.skip(reason='no CUDA device')
  for item in items:
    if 'gpu' in item.keywords:
      item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
  def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
  return load


def rel_l2(a, b):
  import torch
  a = torch.as_tensor(a).detach().cpu().double().reshape(-1)
  b = torch.as_tensor(b).detach().cpu().double().reshape(-1)
  return float((a - b).norm() / (b.norm() + 1e-30))
