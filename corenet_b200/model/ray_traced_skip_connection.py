"""Ray-traced skip connection module (boundary row a5).

Same constructor, parameters (`compress_channels.{weight,bias}`) and forward
signature as SampleGrid2d in src/corenet/model/ray_traced_skip_connection.py:25-144
of the reference.  forward = 1x1 compress conv + fused project/truncate/gather
kernel (csrc/skip.cu), both through the C-ABI.
"""
import torch as t
from torch import nn

from corenet_b200 import ops


class SampleGrid2d(nn.Module):
  def __init__(self, in_channels: int, out_channels: int, output_resolution):
    super().__init__()
    self.compress_channels = nn.Conv2d(in_channels, out_channels, kernel_size=1)
    self.output_resolution = tuple(int(v) for v in output_resolution)

  def forward(self, grid2d: t.Tensor, voxel_projection_matrix: t.Tensor,
              voxel_sample_location: t.Tensor, outside_value: float = 0, flip_x=False, flip_y=False):
    if outside_value != 0 or flip_x or flip_y:
      raise ValueError("corenet_b200: only outside_value=0, flip_x=flip_y=False (the reference's "
                       "only call site, reconstruction_decoder.py:116) are implemented")
    assert grid2d.dim() == 4 and grid2d.dtype == t.float32
    b = grid2d.shape[0]
    assert voxel_sample_location.shape == (b, 3)
    assert voxel_projection_matrix.shape == (b, 4, 4)
    return ops.sample_grid2d(grid2d, self.compress_channels.weight, self.compress_channels.bias,
                             self.output_resolution, voxel_projection_matrix, voxel_sample_location)
