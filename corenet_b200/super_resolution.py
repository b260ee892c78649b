"""Inference plug point and super-resolution wrapper (SURVEY §8b "Inference plug point", f4).

Same interface as the reference's `pipeline.InferenceFn` (src/corenet/pipeline.py:261-276) and
`super_resolution.SuperResolutionInference` / `super_resolution_from_state`
(src/corenet/super_resolution.py:28-129): run the fixed-size model `mult^3` times with shifted sample
offsets and interleave the pmfs.  The softmax runs in the fused CUDA kernel.
"""
from typing import Tuple

import torch as t

from corenet_b200.geometry import transformations


class InferenceFn:
  def __call__(self, input_image, camera_transform, view_to_voxel_transform, grid_offsets, output_resolution):
    raise NotImplementedError()


class MultiOffsetInferenceFn:
  """grid_offsets float32[n_off, B, 3] -> pmf float32[n_off, B, C, D, H, W] (super_resolution.py:28-43)."""

  def __call__(self, input_image, camera_transform, view_to_voxel_transform, grid_offsets):
    raise NotImplementedError()


class CoreNetMultiOffsetInferenceFn(MultiOffsetInferenceFn):
  """The B200 model as a multi-offset inference function: the encoder runs once per image batch, the decoder +
  ray-traced skips + softmax once per offset, each as a replayed CUDA graph (engine.corenet_multi_offset_pmf).
  The reference re-runs the whole model per offset (super_resolution.py:118-126)."""

  def __init__(self, model):
    self.model = model

  def __call__(self, input_image, camera_transform, view_to_voxel_transform, grid_offsets):
    from corenet_b200 import engine
    v2s = camera_transform @ view_to_voxel_transform.inverse()
    return engine.corenet_multi_offset_pmf(self.model, input_image, v2s, grid_offsets)


class SuperResolutionInference(InferenceFn):
  def __init__(self, inference_fn, resolution: Tuple[int, int, int]):
    self.resolution = tuple(resolution)
    self.inference_fn = inference_fn

  def get_resolution_multiplier(self, output_resolution) -> int:
    mult = [o / r for o, r in zip(output_resolution, self.resolution)]
    if any(m != int(m) or m < 1 for m in mult) or min(mult) != max(mult):
      raise ValueError("The output resolution should be divisible by the native resolution")
    return int(mult[0])

  def get_native_offsets(self, output_resolution, grid_offsets: t.Tensor) -> t.Tensor:
    """float32[mult^3, B, 3]: sample offsets in the native grid (super_resolution.py:66-90)."""
    mult = self.get_resolution_multiplier(tuple(output_resolution))
    zz, yy, xx = t.meshgrid([t.arange(mult)] * 3, indexing="ij")
    offsets = (t.stack([xx, yy, zz], -1) / mult).reshape(-1, 3).to(grid_offsets.device)
    return offsets[:, None] + grid_offsets[None, :] / mult

  def __call__(self, input_image, camera_transform, view_to_voxel_transform, grid_offsets, output_resolution):
    native_offsets = self.get_native_offsets(output_resolution, grid_offsets)
    mult = self.get_resolution_multiplier(tuple(output_resolution))
    b = input_image.shape[0]
    scale = transformations.scale([1 / mult] * 3).to(view_to_voxel_transform.device)
    pmfs = self.inference_fn(input_image, camera_transform, view_to_voxel_transform @ scale, native_offsets)
    _, _, c, d, h, w = pmfs.shape
    pmfs = pmfs.reshape(mult, mult, mult, b, c, d, h, w).permute(3, 4, 5, 0, 6, 1, 7, 2)
    return pmfs.reshape(b, c, mult * d, mult * h, mult * w)


def super_resolution_from_model(model) -> SuperResolutionInference:
  """Plugs a corenet_b200 CoreNet into the eval pipeline (super_resolution.py:115-129)."""
  return SuperResolutionInference(CoreNetMultiOffsetInferenceFn(model), model.config.decoder.resolution)


def super_resolution_from_state(state) -> SuperResolutionInference:
  """Same name and argument as the reference's (super_resolution.py:115): `state.model` is the CoreNet."""
  return super_resolution_from_model(state.model)
