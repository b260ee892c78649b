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


# ----------------------------------------------------------------------------------------------------------------
# Config-file surface of the hot path: configs/models/*.json5 and configs/paper_tf_models/*.json5 of the reference
# (written by generate_configs.py) carry, besides dataset/IO settings that are out of scope here, the three blocks
# the path needs: `voxelization_config` (GT pipeline), `data_loader.batch_size` and the model hyper-parameters.
# ----------------------------------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class Resolution(_DictMixin):
  """configuration.py:86-92 of the reference: field order is (depth, height, width)."""
  depth: int
  height: int
  width: int


@dataclasses.dataclass(frozen=True)
class VoxelizationConfig(_DictMixin):
  """configuration.py:95-151 of the reference (same field names and defaults)."""
  task_type: TaskType
  resolution: Resolution
  sub_grid_sampling: bool = False
  conservative_rasterization: bool = True
  voxelization_image_resolution_multiplier: int = 5
  voxelization_projection_depth_multiplier: int = 1

  @classmethod
  def from_dict(cls, d):
    kw = dict(d)
    kw["task_type"] = TaskType(kw["task_type"]) if not isinstance(kw["task_type"], TaskType) else kw["task_type"]
    if isinstance(kw["resolution"], dict):
      kw["resolution"] = Resolution(**kw["resolution"])
    return cls(**{f.name: kw[f.name] for f in dataclasses.fields(cls) if f.name in kw})


def parse_json5(text: str):
  """The JSON5 subset the reference's generated configs use: // and /* */ comments, unquoted or single-quoted keys,
  single-quoted strings, trailing commas.  (The reference depends on the `json5` package, pipeline.py:28.)"""
  import json
  import re
  out, i, n = [], 0, len(text)
  while i < n:                                   # strip comments, normalise quotes (outside strings only)
    c = text[i]
    if c in "\"'":
      j = i + 1
      while j < n and text[j] != c:
        j += 2 if text[j] == "\\" else 1
      body = text[i + 1:j]
      if c == "'":
        body = body.replace('\\\'', "'").replace('"', '\\"')
      out.append('"' + body + '"')
      i = j + 1
    elif text.startswith("//", i):
      j = text.find("\n", i)
      i = n if j < 0 else j
    elif text.startswith("/*", i):
      j = text.find("*/", i + 2)
      i = n if j < 0 else j + 2
    else:
      out.append(c)
      i += 1
  s = "".join(out)
  # quote bare keys, drop trailing commas -- applied outside of string literals only
  parts = re.split(r'("(?:[^"\\]|\\.)*")', s)
  for k in range(0, len(parts), 2):
    parts[k] = re.sub(r'([{,]\s*)([A-Za-z_$][A-Za-z0-9_$]*)(\s*:)', r'\1"\2"\3', parts[k])
  s = "".join(parts)
  parts = re.split(r'("(?:[^"\\]|\\.)*")', s)
  for k in range(0, len(parts), 2):
    parts[k] = re.sub(r',(\s*[}\]])', r'\1', parts[k])
  return json.loads("".join(parts))


def expand_templates(cfg):
  """Resolves `string_templates` ({key} placeholders, possibly nested) in every string of the config
  (configuration.py:305-345 of the reference)."""
  table = {}
  for e in cfg.get("string_templates", []):
    v = e["value"]
    for _ in range(8):
      nv = v.format_map(_Keep(table))
      if nv == v:
        break
      v = nv
    table[e["key"]] = v

  def walk(x):
    if isinstance(x, str):
      return x.format_map(_Keep(table)) if "{" in x else x
    if isinstance(x, list):
      return [walk(v) for v in x]
    if isinstance(x, dict):
      return {k: (v if k == "string_templates" else walk(v)) for k, v in x.items()}
    return x
  return walk(cfg)


class _Keep(dict):
  def __missing__(self, key):
    return "{" + key + "}"


def load_config(path: str):
  """Parses a reference config file and expands its string templates.  Returns a plain dict."""
  with open(path) as f:
    return expand_templates(parse_json5(f.read()))


def hot_path_settings(cfg) -> dict:
  """The settings of the hot path from a parsed train (configs/models/*.json5) or eval / tf-eval
  (configs/paper_tf_models/*.json5) config: voxelization config, per-GPU batch size, loss, model config."""
  root = cfg.get("train") or cfg.get("eval_config") or cfg
  data = root["data"]
  vox = VoxelizationConfig.from_dict(data["voxelization_config"])
  out = {"voxelization_config": vox, "batch_size": data["data_loader"]["batch_size"],
         "loss": "iou_fgbg" if vox.task_type == TaskType.FG_BG else "xent_times_iou_agnostic",
         "resolution": dataclasses.astuple(vox.resolution),
         # state.create_initial_state hands the decoder the REVERSED tuple (state.py:58; identical for the cubic grids
         # the decoder supports)
         "model_resolution": dataclasses.astuple(vox.resolution)[::-1]}
  if "train" in cfg:
    tr = cfg["train"]
    out.update(initial_learning_rate=tr.get("initial_learning_rate", 0.0004), adam_epsilon=tr.get("adam_epsilon", 1e-4),
               latent_channels=tr.get("latent_channels", 64), skip_fraction=tr.get("skip_fraction", 0.75),
               last_upscale_factor=tr.get("last_upscale_factor", 2))
  return out
