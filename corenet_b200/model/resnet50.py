"""ResNet-50 feature extractor: parameter container with the reference's layout.

Boundary row a2.  Same module tree / state_dict keys / initialisation order as
src/corenet/model/resnet50.py:26-186 of the reference (Caffe-style: stride on
the first 1x1, conv biases, BatchRenorm eps=1e-3, Kaiming-normal conv init), so
`model.encoder.load_state_dict(<reference encoder checkpoint>)` works.  The
arithmetic is not here: `CoreNet.forward` runs the CUDA engine
(corenet_b200/engine.py) over these parameters.
"""
import collections
from typing import NamedTuple, Sequence, Tuple

import torch as t
from torch import nn

from corenet_b200.model import batch_renorm


class ResNet50Features(NamedTuple):
  stage1_64x128x128: t.Tensor
  stage2_256x64x64: t.Tensor
  stage3_512x32x32: t.Tensor
  stage4_1024x16x16: t.Tensor
  stage5_2048x8x8: t.Tensor
  global_average_2048: t.Tensor


def _unit(cin: int, cout: int, k: int, stride: int = 1) -> nn.Sequential:
  """conv + BRN pair, initialised like ResNetBlock.init_weights (resnet50.py:39-46)."""
  conv = nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2)
  seq = nn.Sequential(collections.OrderedDict(conv=conv, bn=batch_renorm.BatchRenorm(cout, eps=0.001)))
  nn.init.kaiming_normal_(conv.weight, mode="fan_in", nonlinearity="relu")
  return seq


class Bottleneck(nn.Module):
  """One residual block.  `stride` is None for an identity block, else the
  stride of its projection shortcut (resnet50.py:49-115)."""

  def __init__(self, cin: int, filters: Tuple[int, int, int], stride=None, tap_pre_relu=False):
    super().__init__()
    f1, f2, f3 = filters
    self.out_channels = f3
    self.stride = stride
    self.return_output_before_relu = tap_pre_relu
    self.op_a = _unit(cin, f1, 1, stride or 1)
    self.op_b = _unit(f1, f2, 3)
    self.op_c = _unit(f2, f3, 1)
    if stride is not None:
      self.shortcut = _unit(cin, f3, 1, stride)


# (stage name, block letters, bottleneck filters, stride of the first block)
STAGES: Sequence = (
    ("stage2", "abc", (64, 64, 256), 1),
    ("stage3", "abcd", (128, 128, 512), 2),
    ("stage4", "abcdef", (256, 256, 1024), 2),
    ("stage5", "abc", (512, 512, 2048), 2),
)


class ResNet50FeatureExtractor(nn.Module):
  def __init__(self):
    super().__init__()
    stem = nn.Conv2d(3, 64, kernel_size=7, stride=2)
    self.stage1 = nn.Sequential(collections.OrderedDict(pad=nn.ZeroPad2d(3), conv=stem))
    nn.init.kaiming_normal_(stem.weight, mode="fan_in", nonlinearity="relu")
    self.stage1_part2 = nn.Sequential(collections.OrderedDict(
        bn=batch_renorm.BatchRenorm(64, eps=0.001), relu=nn.ReLU(), pad=nn.ZeroPad2d(1),
        pool=nn.MaxPool2d(kernel_size=3, stride=2)))
    cin = 64
    for name, letters, filters, stride in STAGES:
      blocks = collections.OrderedDict()
      for i, letter in enumerate(letters):
        blocks[letter] = Bottleneck(cin, filters, stride=stride if i == 0 else None,
                                    tap_pre_relu=(i == len(letters) - 1))
        cin = filters[2]
      setattr(self, name, nn.Sequential(blocks))

  def forward(self, input_image: t.Tensor) -> ResNet50Features:
    """float32[B,3,H,W] (already Caffe-preprocessed) -> the six feature maps (NCHW)."""
    from corenet_b200 import engine
    return engine.encoder_forward(self, input_image)


def preprocess_image_caffe(image: t.Tensor) -> t.Tensor:
  """uint8[B,3,H,W] RGB -> float32 BGR + mean (resnet50.py:189-204 of the reference).

  Host-visible helper with the reference's semantics (it ADDS the mean); the
  engine fuses the same arithmetic into its NHWC staging kernel."""
  assert image.dtype == t.uint8 and image.dim() == 4 and image.shape[1] == 3
  image = image.to(t.float32).flip(1)
  return image + image.new_tensor([103.939, 116.779, 123.68])[None, :, None, None]
