"""Pins oracle/voxelize_oracle.py with the reference's three rasteriser known-answer tests
(src/corenet/test/voxelization_test.py:53-147).  Beyond these vectors rasteriser parity is unpinned."""
import numpy as np
import pytest

from oracle import corenet_oracle as O
from oracle import fill_voxels_oracle as F
from oracle import voxelize_oracle as V
from tests.conftest import cube_mesh


def test_diagonal_quad():
  quad = np.array([[[0, 0, 0], [1, 0, 1], [0, 1, 0]], [[1, 0, 1], [0, 1, 0], [1, 1, 1]]], np.float32)
  g = V.voxelize_mesh_oracle(quad, [2], (4, 4, 4), O.scale([4, 4, 4]).numpy(), image_resolution_multiplier=16)
  g = F.fill_inside_voxels_oracle(g)
  exp = np.zeros((4, 4, 4), np.float32)
  for z in range(4):
    exp[z, :, z] = 1
  np.testing.assert_array_equal(g[0], exp)


def test_conservative():
  c = cube_mesh(0.99)
  eye = np.eye(4, dtype=np.float32)
  g = V.voxelize_mesh_oracle(c, [12], (3, 3, 3), eye, image_resolution_multiplier=1)
  e = np.zeros((3, 3, 3))
  e[1, 1, [0, 2]] = e[1, [0, 2], 1] = e[[0, 2], 1, 1] = 1
  np.testing.assert_array_equal(g[0], e)
  g = V.voxelize_mesh_oracle(c, [12], (3, 3, 3), eye, image_resolution_multiplier=1,
                             conservative_rasterization=True)
  e = np.ones((3, 3, 3))
  e[1, 1, 1] = 0
  np.testing.assert_array_equal(g[0], e)


def test_sub_grid():
  c = cube_mesh(0.99)
  eye = np.eye(4, dtype=np.float32)
  g = V.voxelize_mesh_oracle(c, [12], (3, 3, 3), eye, sub_grid_sampling=True, image_resolution_multiplier=9,
                             conservative_rasterization=True)
  g = F.fill_inside_voxels_oracle(g)
  e = np.zeros((1, 7, 7, 7))
  e[0, 2:5, 2:5, 2:5] = 1
  np.testing.assert_array_equal(g, e)
  e = np.zeros((1, 3, 3, 3))
  e[0, 1, 1, 1] = 1
  np.testing.assert_array_equal(V.get_sub_grid_centers(g), e)
  cubes = np.concatenate([c, c - 0.5])
  tr = np.stack([O.translate([-0.5, 0, 0]).numpy(), O.translate([0.5, 1, 1]).numpy()])
  g = V.voxelize_mesh_oracle(cubes, [12, 12], (3, 3, 3), tr, sub_grid_sampling=True,
                             image_resolution_multiplier=9, conservative_rasterization=True)
  gc = V.get_sub_grid_centers(F.fill_inside_voxels_oracle(g))
  e1 = np.zeros((3, 3, 3)); e1[1, 1, [0, 1]] = 1
  e2 = np.zeros((3, 3, 3)); e2[1, [1, 2], 1] = e2[2, [1, 2], 1] = 1
  np.testing.assert_array_equal(gc[0], e1)
  np.testing.assert_array_equal(gc[1], e2)


def test_even_multiplier_rejected():
  with pytest.raises(ValueError):
    V.voxelize_mesh_oracle(cube_mesh(0.99), [12], (3, 3, 3), np.eye(4, dtype=np.float32),
                           sub_grid_sampling=True, image_resolution_multiplier=8)
