"""CoreNet: image -> 3-D occupancy / semantic grid (boundary row a7).

Same constructor, attributes (`config`, `encoder`, `decoder`), forward
signature, parameters and state_dict keys as
src/corenet/model/core_net.py:25-61 of the reference, so `state.decode_state`,
`t.optim.Adam(model.parameters())` and `DistributedDataParallel(model)` work
unchanged.  forward() runs the sm_100a engine; there is no CPU path.
"""
import torch as t
from torch import nn

from corenet_b200 import configuration
from corenet_b200 import engine
from corenet_b200.model import reconstruction_decoder
from corenet_b200.model import resnet50


class CoreNet(nn.Module):
  def __init__(self, config: configuration.CoreNetConfig):
    super().__init__()
    self.config = config
    self.encoder = resnet50.ResNet50FeatureExtractor()
    self.decoder = reconstruction_decoder.ReconstructionDecoder(config.decoder)

  def forward(self, image: t.Tensor, voxel_projection_matrix: t.Tensor,
              voxel_sample_locations: t.Tensor) -> t.Tensor:
    """uint8[B,3,256,256], float32[B,4,4], float32[B,3] -> logits float32[B,C,D,H,W]."""
    return engine.corenet_forward(self, image, voxel_projection_matrix, voxel_sample_locations)

  def encode_features(self, image: t.Tensor) -> resnet50.ResNet50Features:
    """Encoder only (no grad): the six feature maps of resnet50.py:26-32 as NCHW tensors."""
    eng = engine.get_engine(self)
    plan = eng.get_plan(image.shape[0], image.device, False)
    with t.no_grad():
      plan.forward(image.contiguous(), None, None, self.training, want_features=True)
      return resnet50.ResNet50Features(*plan.features_nchw())

  def _apply(self, fn, *a, **k):
    r = super()._apply(fn, *a, **k)
    eng = self.__dict__.get("_crn_engine")
    if eng is not None:
      eng.invalidate()
    return r
