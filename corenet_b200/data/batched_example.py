"""Device-resident ground-truth voxelisation of a batch.

`BatchedExample` / `batch` mirror batched_example.py:32-97 (object -> view space transform of every mesh),
`voxelize_example(ex, resolution, ...)` has the reference's call shape; `voxelize` is the same step on unpacked
arguments and mirrors `corenet.data.batched_example.voxelize`
(src/corenet/data/batched_example.py:121-197 of the reference): shifted world->voxel transform per scene,
one voxeliser call for all meshes of the batch, flood fill of the enclosed pockets, optional sub-grid
centres, label * occupancy, max over the meshes of a scene -> int32[B, D, H, W].  Everything stays on the
GPU: CUDA rasteriser -> `fill_inside_voxels_gpu(inplace=True)` -> merge kernel; the reference's GL context and
its two host round trips (`gl/rasterizer.py:156,221-224`, `batched_example.py:181`) are gone.
"""
import dataclasses
from typing import Callable, List, Optional, Sequence, Tuple

import torch as t

from corenet_b200 import ops
from corenet_b200.cc import fill_voxels
from corenet_b200.geometry import transformations
from corenet_b200.geometry import voxelization


@dataclasses.dataclass(frozen=True)
class BatchedExample:
  """Field for field the reference's BatchedExample (batched_example.py:32-64)."""
  vertices: t.Tensor                      # float32[total_triangles, 3, 3], view space
  view_transform: t.Tensor                # float32[B, 4, 4]
  camera_transform: t.Tensor              # float32[B, 4, 4]
  mesh_num_tri: List[t.Tensor]            # per scene int32[num_meshes]
  mesh_labels: List[t.Tensor]             # per scene int32[num_meshes]
  input_image: t.Tensor                   # uint8[B, 3, H, W]
  scene_id: List[str]
  grid_sampling_offset: t.Tensor          # float32[B, 3] in [0, 1]^3
  v2x_transform: Optional[t.Tensor] = None
  grid: Optional[t.Tensor] = None

  def to(self, device, non_blocking: bool = False) -> "BatchedExample":
    """Tensors (and lists of tensors) moved to `device` (the reference's TensorContainerMixin.to / .cuda)."""
    mv = lambda v: (v.to(device, non_blocking=non_blocking) if isinstance(v, t.Tensor)
                    else [mv(x) for x in v] if isinstance(v, list) else v)
    return dataclasses.replace(self, **{f.name: mv(getattr(self, f.name)) for f in dataclasses.fields(self)})

  def cuda(self) -> "BatchedExample":
    return self.to("cuda")


def batch(examples) -> BatchedExample:
  """Batches dataset elements (batched_example.py:67-97 of the reference): every mesh goes from object space to
  the scene's view space, o2v = view_transform @ o2w.  The reference loops over meshes; here every triangle picks its
  mesh's matrix and ONE batched product transforms the whole batch (on the device the elements live on)."""
  with t.no_grad():
    o2v, tri_per_mesh = [], []
    for ex in examples:
      w2v = t.as_tensor(ex.view_transform, dtype=t.float32)
      for num_tri, o2w in zip(ex.mesh_num_tri, ex.o2w_transforms):
        o2v.append(t.matmul(w2v, t.as_tensor(o2w, dtype=t.float32)))
        tri_per_mesh.append(int(num_tri))
    verts = t.cat([t.as_tensor(ex.mesh_vertices, dtype=t.float32)[:int(sum(int(n) for n in ex.mesh_num_tri))]
                   for ex in examples], 0)
    o2v = t.stack(o2v, 0).to(verts.device)
    per_tri = o2v[t.repeat_interleave(t.arange(len(tri_per_mesh), device=verts.device),
                                      t.tensor(tri_per_mesh, device=verts.device))]
    all_vertices = transformations.transform_mesh(verts[:, None], per_tri)[:, 0]
    return BatchedExample(
        vertices=all_vertices,
        view_transform=t.stack([e.view_transform for e in examples], 0),
        camera_transform=t.stack([e.camera_transform for e in examples], 0),
        mesh_num_tri=[e.mesh_num_tri for e in examples],
        mesh_labels=[e.mesh_labels for e in examples],
        input_image=t.stack([e.input_image for e in examples], 0),
        scene_id=[e.scene_id for e in examples],
        grid_sampling_offset=all_vertices.new_ones([len(examples), 3]) * 0.5)


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


def voxelize_example(ex: BatchedExample, resolution: Tuple[int, int, int], **kwargs) -> BatchedExample:
  """The reference's call shape, `voxelize(ex, resolution, ...) -> ex with v2x_transform and grid replaced`
  (batched_example.py:121-197), on top of `voxelize` above."""
  v2x, grid = voxelize(ex.vertices, ex.mesh_num_tri, ex.grid_sampling_offset, resolution, **kwargs)
  return dataclasses.replace(ex, v2x_transform=v2x, grid=grid)


def voxelize_batch(b: BatchedExample, voxelization_config) -> BatchedExample:
  """pipeline.voxelize_batch (pipeline.py:126-150 of the reference): the voxel content and the rasteriser settings of a
  `configuration.VoxelizationConfig` (SEMANTIC: the mesh's class, FG_BG: 1) applied to a batch."""
  from corenet_b200 import configuration
  content = {configuration.TaskType.SEMANTIC: VoxelContentSemanticLabel(b.mesh_labels),
             configuration.TaskType.FG_BG: voxel_content_1}[voxelization_config.task_type]
  return voxelize_example(
      b, dataclasses.astuple(voxelization_config.resolution), voxel_content_fn=content,
      sub_grid_sampling=voxelization_config.sub_grid_sampling,
      image_resolution_multiplier=voxelization_config.voxelization_image_resolution_multiplier,
      conservative_rasterization=voxelization_config.conservative_rasterization,
      projection_depth_multiplier=voxelization_config.voxelization_projection_depth_multiplier)
