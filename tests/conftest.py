import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
  """gpu-marked tests are skipped (not failed with a NewBatch error) where no CUDA device
  exists; the engine itself has no CPU fallback."""
  try:
    import torch
    have_gpu = torch.cuda.is_available()
  except Exception:  # pylint: disable=broad-except
    have_gpu = False
  if have_gpu:
    return
  skip = pytest.mark.skip(reason="needs a CUDA device")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
  from hanabi_b200 import build
  return build.build()
