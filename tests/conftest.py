import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `pytest -m gpu` on the GPU box")


@pytest.fixture(scope="session")
def ctx():
  """One libdd_b200 context for the whole GPU session (fails loudly when the library or the GPU is missing)."""
  from deepdenoiser_b200 import _lib
  return _lib.Context(0)
