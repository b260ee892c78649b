"""Device-resident ground-truth voxelisation of a batch.

`voxelize` mirrors `corenet.data.batched_example.voxelize`
(src/corenet/data/batched_example.py:121-197 of the reference): shifted world->voxel transform per scene,
one voxeliser call for all meshes of the batch, flood fill of the enclosed pockets, optional sub-grid
centres, label * occupancy, max over the meshes of a scene -> int32[B, D, H, W].  Everything stays on the
GPU: CUDA rasteriser -> `fill_inside_voxels_gpu(inplace=True)` -> merge kernel; the reference's GL context and
its two host round trips (`gl/rasterizer.py:156,221-224`, `batched_example.py:181`) are gone.
"""
from typing import Callable, List, Sequence, Tuple

import torch as t

from corenet_b200 import ops
from corenet_b200.cc import fill_voxels
from corenet_b200.geometry import transformations
from corenet_b200.geometry import voxelization


def voxel_content_mesh_index(batch_idx: int, mesh_idx: int) -> int:
  return mesh_idx + 1


def voxel_content_1(batch_idx: int, mesh_idx: int) -> int:
  return 1


class VoxelContentSemanticLabel:
  def __init__(self, semantic_labels: Sequence):
    self.semantic_labels = semantic_labels

  def __call__(self, batch_idx: int, mesh_idx: int) -> int:
    return int(self.semantic_labels[batch_idx][mesh_idx])


def voxelize(vertices: t.Tensor, mesh_num_tri: List[t.Tensor], grid_sampling_offset: t.Tensor,
             resolution: Tuple[int, int, int],
             voxel_content_fn: Callable[[int, int], int] = voxel_content_mesh_index,
             sub_grid_sampling: bool = False, conservative_rasterization: bool = False,
             image_resolution_multiplier=4, projection_depth_multiplier: int = 1, fill_inside: bool = True):
  """vertices: float32[total_triangles, 3, 3] (view space); mesh_num_tri: per scene int32[num_meshes];
  grid_sampling_offset: float32[B, 3].  Returns (v2x_transform float32[B,4,4], grid int32[B,D,H,W] on CUDA)."""
  with t.no_grad():
    d, h, w = resolution
    m = max(d, h, w)
    batch_size = grid_sampling_offset.shape[0]
    batch_v2x = transformations.scale([m, m, m]).expand(batch_size, 4, 4)
    grid_shift = transformations.translate(grid_sampling_offset.cpu() - 0.5)
    shifted_w2x = t.matmul(grid_shift, batch_v2x)
    num_meshes = [len(v) for v in mesh_num_tri]
    mesh_v2x = t.cat([shifted_w2x[i:i + 1].expand(n, 4, 4) for i, n in enumerate(num_meshes)], 0)
    meshes_grid = voxelization.voxelize_mesh(
        vertices, t.cat([t.as_tensor(v, dtype=t.int32) for v in mesh_num_tri], 0), resolution, mesh_v2x,
        sub_grid_sampling=sub_grid_sampling, image_resolution_multiplier=image_resolution_multiplier,
        conservative_rasterization=conservative_rasterization,
        projection_depth_multiplier=projection_depth_multiplier)
    if fill_inside:
      fill_voxels.fill_inside_voxels_gpu(meshes_grid, inplace=True)
    if sub_grid_sampling:
      meshes_grid = voxelization.get_sub_grid_centers(meshes_grid).contiguous()
    dev = meshes_grid.device
    labels = t.tensor([voxel_content_fn(b, i) for b, n in enumerate(num_meshes) for i in range(n)],
                      dtype=t.float32, device=dev)
    scene = t.tensor([b for b, n in enumerate(num_meshes) for _ in range(n)], dtype=t.int32, device=dev)
    grid = ops.merge_mesh_grids(meshes_grid, scene, labels, batch_size)
    return batch_v2x.contiguous(), grid
