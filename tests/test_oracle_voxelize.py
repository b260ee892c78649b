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


# ---------------------------------------------------------------------------------------------------------------
# Properties any rasteriser following the GL rules has (no reference vectors exist beyond the three above; the CUDA
# rasteriser is compared bit for bit with this oracle on the GPU, so these transfer to it).
def _soup(n, res, seed):
  rng = np.random.default_rng(seed)
  c = rng.uniform(1.0, res - 1.0, size=(n, 1, 3))
  return (c + rng.normal(0, res / 6.0, size=(n, 3, 3))).astype(np.float32)


@pytest.mark.parametrize("seed", [0, 1])
def test_conservative_covers_plain(seed):
  """GL_CONSERVATIVE_RASTERIZATION_NV generates a fragment for every pixel the primitive touches: the fragments of
  plain rasterisation (pixel centre inside) are a subset, and both interpolate the same attributes at the same pixel
  centres, so the voxel sets nest too."""
  tri = _soup(40, 12, seed)
  eye = np.eye(4, dtype=np.float32)
  plain = V.voxelize_mesh_oracle(tri, [40], (12, 12, 12), eye, image_resolution_multiplier=3)
  cons = V.voxelize_mesh_oracle(tri, [40], (12, 12, 12), eye, image_resolution_multiplier=3,
                                conservative_rasterization=True)
  assert plain.sum() > 50 and cons.sum() > plain.sum()
  assert ((plain > 0) & (cons == 0)).sum() == 0


def test_integer_translation_shifts_the_grid_exactly():
  """Vertices on a dyadic lattice (k/8): every fp32 step is exact, so moving the mesh by whole voxels through the
  view2voxel matrix moves the occupancy by whole voxels."""
  rng = np.random.default_rng(3)
  tri = (rng.integers(16, 48, size=(25, 3, 3)) / 8.0).astype(np.float32)        # inside [2, 6)^3
  base = V.voxelize_mesh_oracle(tri, [25], (12, 12, 12), np.eye(4, dtype=np.float32), image_resolution_multiplier=4)
  moved = V.voxelize_mesh_oracle(tri, [25], (12, 12, 12), O.translate([3, 1, 2]).numpy(), image_resolution_multiplier=4)
  assert base.sum() > 30
  exp = np.zeros_like(base)
  exp[:, 2:, 1:, 3:] = base[:, :-2, :-1, :-3]
  np.testing.assert_array_equal(moved, exp)


def test_closed_cube_fills_to_the_expected_solid():
  """A watertight axis-aligned box with faces strictly inside voxels: surface voxels are exactly the voxels the faces
  pass through, and fill_inside_voxels makes the solid box of those voxels (voxelization.py:115-164 + fill)."""
  lo, hi = 2.3, 9.6
  c = cube_mesh(0.0) / 3.0 * (hi - lo) + lo
  g = V.voxelize_mesh_oracle(c, [12], (12, 12, 12), np.eye(4, dtype=np.float32), image_resolution_multiplier=5)
  shell = np.zeros((12, 12, 12), np.float32)
  shell[2:10, 2:10, 2:10] = 1
  shell[3:9, 3:9, 3:9] = 0
  np.testing.assert_array_equal(g[0], shell)
  solid = np.zeros((12, 12, 12), np.float32)
  solid[2:10, 2:10, 2:10] = 1
  np.testing.assert_array_equal(F.fill_inside_voxels_oracle(g)[0], solid)


def test_per_mesh_grids_are_independent_and_order_free():
  """Every mesh writes its own grid (voxelize.frag:45-57: layer = mesh id); triangle order inside a mesh is irrelevant
  (no depth test, no blending: a voxel is set or not)."""
  a, b = _soup(15, 10, 5), _soup(20, 10, 6)
  eye = np.eye(4, dtype=np.float32)
  both = V.voxelize_mesh_oracle(np.concatenate([a, b]), [15, 20], (10, 10, 10), eye, image_resolution_multiplier=3)
  np.testing.assert_array_equal(both[0], V.voxelize_mesh_oracle(a, [15], (10, 10, 10), eye, image_resolution_multiplier=3)[0])
  np.testing.assert_array_equal(both[1], V.voxelize_mesh_oracle(b[::-1], [20], (10, 10, 10), eye, image_resolution_multiplier=3)[0])
