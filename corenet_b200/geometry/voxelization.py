"""Drop-in for the reference's `corenet.geometry.voxelization` (boundary row a10).

`voxelize_mesh` keeps the signature and semantics of
src/corenet/geometry/voxelization.py:32-164, but rasterises with the CUDA
kernel in csrc/voxelize.cu instead of an EGL/OpenGL draw; nothing round-trips
through host memory.  `get_sub_grid_centers` is :167-182.
"""
from typing import Tuple

import torch as t

from corenet_b200 import ops


def _dynamic_tile(lengths: t.Tensor) -> t.Tensor:
  """[n0, n1, ..] -> n0 zeros, n1 ones, ... (misc_util.dynamic_tile, misc_util.py:32-48)."""
  return t.repeat_interleave(t.arange(lengths.shape[0], dtype=t.int32, device=lengths.device),
                             lengths.to(t.int64)).to(t.int32)


def voxelize_mesh(triangles, mesh_num_tri, resolution: Tuple[int, int, int], view2voxel,
                  sub_grid_sampling: bool = False, image_resolution_multiplier: float = 4,
                  conservative_rasterization: bool = False, projection_depth_multiplier: int = 1,
                  cuda_device=None):
  """Returns float32[num_meshes, D, H, W] (or [.., 2D+1, 2H+1, 2W+1] with sub-grid sampling) on CUDA."""
  dev = t.device("cuda", cuda_device if cuda_device is not None else t.cuda.current_device())
  triangles = t.as_tensor(triangles, dtype=t.float32)
  if triangles.dtype != t.float32:
    raise ValueError(f"Expecting type 'torch.float32', found '{triangles.dtype}'")
  assert triangles.shape[1:] == (3, 3)
  mesh_num_tri = t.as_tensor(mesh_num_tri, dtype=t.int32)
  assert mesh_num_tri.dim() == 1
  view2voxel = t.as_tensor(view2voxel, dtype=t.float32)
  if view2voxel.dim() == 2:
    view2voxel = view2voxel[None].expand(len(mesh_num_tri), 4, 4)
  assert view2voxel.shape == (len(mesh_num_tri), 4, 4)
  if sub_grid_sampling and image_resolution_multiplier % 2 == 0:
    raise ValueError("image_resolution_multiplier must be off if sub_grid_sampling is True")
  if sub_grid_sampling and projection_depth_multiplier == 0:
    raise ValueError("projection_depth_multiplier must be 1 if sub_grid_sampling is True")
  assert int(mesh_num_tri.sum()) == triangles.shape[0]
  tri_mesh = _dynamic_tile(mesh_num_tri).to(dev, non_blocking=True)
  return ops.voxelize_mesh(triangles.to(dev, non_blocking=True), tri_mesh, len(mesh_num_tri), tuple(resolution),
                           view2voxel.to(dev, non_blocking=True).contiguous(), sub_grid_sampling,
                           image_resolution_multiplier, conservative_rasterization,
                           projection_depth_multiplier)


def get_sub_grid_centers(grid: t.Tensor) -> t.Tensor:
  """float32[B, 2D+1, 2H+1, 2W+1] -> occupancy at the sub-grid centres float32[B, D, H, W]."""
  return grid[:, 1::2, 1::2, 1::2]
