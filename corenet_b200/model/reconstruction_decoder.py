"""3-D reconstruction decoder: parameter container with the reference's layout.

Boundary row a3/a4.  Same module tree / state_dict keys as
src/corenet/model/reconstruction_decoder.py:29-95 of the reference
(stage_0 Linear, stage_1..6 Sequentials r1/b1/c1/r2/b2/t1, rt_skip_2..5).
The arithmetic runs in the CUDA engine (corenet_b200/engine.py).
"""
import collections

from torch import nn

from corenet_b200 import configuration
from corenet_b200.model import batch_renorm
from corenet_b200.model import ray_traced_skip_connection

# (stage, conv kernel, convT kernel, convT padding, conv out channels, convT out channels,
#  encoder channels feeding the skip that follows the stage)
_PYRAMID = ((2, 3, 3, 1, 256, 128, 2048), (3, 5, 7, 3, 128, 64, 1024), (4, 5, 7, 3, 64, 32, 512),
            (5, 5, 7, 3, 32, 16, 256), (6, 5, 7, 3, 16, None, None))


class ReconstructionDecoder(nn.Module):
  def __init__(self, config: configuration.DecoderConfig):
    super().__init__()
    self.config = config
    depth, height, width = config.resolution
    div = 16 * config.last_upscale_factor
    assert depth % div == 0 and height % div == 0 and width % div == 0
    ir = (depth // div, height // div, width // div)
    if ir != (4, 4, 4) or config.last_upscale_factor != 2:
      # reconstruction_decoder.py:53-54: ConvTranspose3d(k=4, stride=ir) on a 1^3 input only
      # lines up with the skip grids when ir == 4.
      raise ValueError("the CoReNet decoder only exists at 128^3 with last_upscale_factor=2")
    bn = lambda c: batch_renorm.BatchRenorm(c, eps=0.001)
    lat = config.latent_channels
    self.stage_0 = nn.Linear(2048, lat)
    self.stage_1 = nn.Sequential(collections.OrderedDict(
        r1=nn.ReLU(), b1=bn(lat + 3), t1=nn.ConvTranspose3d(lat + 3, 256, 4, stride=ir)))
    cin, grid = 256, 4
    for stage, k, kt, pt, mid, t_out, enc_c in _PYRAMID:
      last = t_out is None
      t_out = config.num_output_channels if last else t_out
      layers = collections.OrderedDict(
          r1=nn.ReLU(), b1=bn(cin), c1=nn.Conv3d(cin, mid, k, padding=k // 2),
          r2=nn.ReLU(), b2=bn(mid),
          t1=nn.ConvTranspose3d(mid, t_out, kt, stride=2, padding=pt, output_padding=1))
      setattr(self, f"stage_{stage}", nn.Sequential(layers))
      grid *= 2
      if not last:
        skip_c = round(t_out * config.skip_fraction)
        if skip_c > 0:
          setattr(self, f"rt_skip_{stage}", ray_traced_skip_connection.SampleGrid2d(
              enc_c + 3, skip_c, (grid, grid, grid)))
        cin = t_out + skip_c
