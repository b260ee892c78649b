"""Pins oracle/fill_voxels_oracle.c: reference known-answer grids, scipy restatement, compiled reference."""
import numpy as np
import pytest

from oracle import build_ref
from oracle import fill_voxels_oracle as F
from tests.conftest import fill_test_grids


def test_reference_known_answers():
  # src/corenet/test/voxelization_test.py:197-244 (float32 and uint8 variants)
  grids, expected = fill_test_grids()
  np.testing.assert_array_equal(F.fill_inside_voxels_oracle(grids), expected)
  np.testing.assert_array_equal(F.fill_inside_voxels_oracle(grids.astype(np.uint8)), expected.astype(np.uint8))


def test_near_face_rule():
  # a pocket open only towards a FAR face is filled; towards a near face it is not
  g = np.ones((1, 5, 5, 5), np.float32)
  g[0, 2, 2, 2:] = 0          # tunnel to x = W-1 (far face)
  assert F.fill_inside_voxels_oracle(g)[0, 2, 2, 2] == 1
  g = np.ones((1, 5, 5, 5), np.float32)
  g[0, 2, 2, :3] = 0          # tunnel to x = 0 (near face)
  assert F.fill_inside_voxels_oracle(g)[0, 2, 2, 2] == 0


@pytest.mark.parametrize("shape,p", [((3, 16, 16, 16), 0.3), ((2, 9, 20, 45), 0.45), ((2, 7, 7, 7), 0.5),
                                     ((1, 40, 33, 70), 0.38), ((1, 1, 1, 1), 0.5), ((2, 1, 8, 3), 0.4)])
def test_vs_scipy_and_compiled_reference(shape, p):
  rng = np.random.default_rng(sum(shape))
  g = (rng.random(shape) < p).astype(np.float32)
  a = F.fill_inside_voxels_oracle(g)
  np.testing.assert_array_equal(a, F.fill_inside_scipy(g))
  ref = build_ref.load()
  if ref is not None:       # oracle/_ref exists when the reference was compiled in the build container
    import torch
    np.testing.assert_array_equal(a, ref.fill_inside_voxels_cpu(torch.from_numpy(g)).numpy())


def test_empty_batch():
  assert F.fill_inside_voxels_oracle(np.zeros((0, 4, 4, 4), np.float32)).shape == (0, 4, 4, 4)
