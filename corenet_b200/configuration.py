"""Model configuration dataclasses with the reference's field names.

Mirrors `DecoderConfig` / `CoreNetConfig` / `TaskType` of the reference
(src/corenet/configuration.py:73-79,297-316) including `to_dict` /
`from_dict`, which is what `state.encode_state` / `decode_state`
(src/corenet/state.py:74-97) use, so checkpoints stay interchangeable.
"""
import dataclasses
import enum
from typing import Tuple


class TaskType(enum.Enum):
  FG_BG = "FG_BG"
  SEMANTIC = "SEMANTIC"


class _DictMixin:
  def to_dict(self):
    return dataclasses.asdict(self)

  @classmethod
  def from_dict(cls, d):
    kwargs = {}
    for f in dataclasses.fields(cls):
      v = d[f.name]
      if dataclasses.is_dataclass(f.type) and isinstance(v, dict):
        v = f.type.from_dict(v)
      elif isinstance(v, list):
        v = tuple(v)
      kwargs[f.name] = v
    return cls(**kwargs)


@dataclasses.dataclass(frozen=True)
class DecoderConfig(_DictMixin):
  resolution: Tuple[int, int, int]      # (depth, height, width) of the output grid
  num_output_channels: int
  last_upscale_factor: int = 2
  latent_channels: int = 64
  skip_fraction: float = 0.75


@dataclasses.dataclass(frozen=True)
class CoreNetConfig(_DictMixin):
  decoder: DecoderConfig


def default_config(num_output_channels: int = 2) -> CoreNetConfig:
  """The only geometry the reference decoder can be instantiated with."""
  return CoreNetConfig(decoder=DecoderConfig(
      resolution=(128, 128, 128), num_output_channels=num_output_channels,
      last_upscale_factor=2, latent_channels=64, skip_fraction=0.75))
