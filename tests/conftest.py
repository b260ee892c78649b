import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
  import numpy as np
  return np.load(os.path.join(ROOT, "tests", "golden", "corenet_reference.npz"), allow_pickle=False)


def cube_mesh(d: float):
  """12-triangle cube [d, 3-d]^3: the mesh of the reference's known-answer tests
  (src/corenet/test/voxelization_test.py:29-48)."""
  import numpy as np
  m, x = d, 3 - d
  return np.array([
      [[m, m, m], [m, x, m], [m, m, x]], [[m, x, x], [m, x, m], [m, m, x]],
      [[x, m, m], [x, x, m], [x, m, x]], [[x, x, x], [x, x, m], [x, m, x]],
      [[m, m, m], [m, m, x], [x, m, m]], [[x, m, x], [m, m, x], [x, m, m]],
      [[m, x, m], [m, x, x], [x, x, m]], [[x, x, x], [m, x, x], [x, x, m]],
      [[m, m, m], [m, x, m], [x, m, m]], [[x, x, m], [m, x, m], [x, m, m]],
      [[m, m, x], [m, x, x], [x, m, x]], [[x, x, x], [m, x, x], [x, m, x]]], np.float32)


def fill_test_grids():
  """The two 4^3 grids of EmptyRegionFillTests (voxelization_test.py:152-195) and their expected fills."""
  import numpy as np
  g1 = np.ones((4, 4, 4), np.float32)
  g1[1:3, 1:3, 1:3] = 0
  g2 = np.zeros((4, 4, 4), np.float32)
  g2[0:3, 0:3, 0:3] = 1
  g2[1, 1, 1] = 0
  e1 = np.ones((4, 4, 4), np.float32)
  e2 = g2.copy()
  e2[1, 1, 1] = 1
  return np.stack([g1, g2]), np.stack([e1, e2])
