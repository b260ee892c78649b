"""Overlay of the B200 path onto the reference's own package names (SURVEY §8b "Model module").

`install()` makes `import corenet.model.core_net`, `corenet.cc.fill_voxels`, `corenet.geometry.voxelization`,
`corenet.model.losses` ... resolve to the modules of this package, so the reference's untouched `state.py`
(`create_initial_state` / `encode_state` / `decode_state`, src/corenet/state.py:45-97), `pipeline.py`
(`TrainPipeline`, `voxelize_batch`, src/corenet/pipeline.py:126-240), `super_resolution.py` and therefore
`train.py` / `eval.py` build, checkpoint, wrap in DistributedDataParallel and run the B200 CoreNet without a change:

    import corenet_b200.compat as compat
    compat.install()              # before the first `import corenet.pipeline`
    from corenet import state, pipeline, super_resolution   # the reference's own files

Only the hot-path modules are overlaid; everything else (`configuration`, `pipeline`, `state`, `data.*`,
`distributed`, ...) stays the reference's.  When no reference checkout is importable, a namespace package `corenet`
with just the overlaid modules is registered, so code written against `corenet.model.core_net.CoreNet` still imports.
"""
import importlib
import sys
import types

# reference module name -> module of this package that replaces it
OVERLAY = {
    "corenet.model.core_net": "corenet_b200.model.core_net",
    "corenet.model.resnet50": "corenet_b200.model.resnet50",
    "corenet.model.reconstruction_decoder": "corenet_b200.model.reconstruction_decoder",
    "corenet.model.ray_traced_skip_connection": "corenet_b200.model.ray_traced_skip_connection",
    "corenet.model.batch_renorm": "corenet_b200.model.batch_renorm",
    "corenet.model.losses": "corenet_b200.model.losses",
    "corenet.cc.fill_voxels": "corenet_b200.cc.fill_voxels",
    "corenet.geometry.voxelization": "corenet_b200.geometry.voxelization",
}
# only when the reference itself is absent (the reference's own versions are kept otherwise)
FALLBACK = {
    "corenet.configuration": "corenet_b200.configuration",
    "corenet.geometry.transformations": "corenet_b200.geometry.transformations",
    "corenet.super_resolution": "corenet_b200.super_resolution",
    "corenet.data.batched_example": "corenet_b200.data.batched_example",
}


def _package(name: str) -> types.ModuleType:
  m = sys.modules.get(name)
  if m is None:
    try:
      m = importlib.import_module(name)
    except ImportError:
      m = types.ModuleType(name)
      m.__path__ = []
      sys.modules[name] = m
      if "." in name:
        parent, _, leaf = name.rpartition(".")
        setattr(_package(parent), leaf, m)
  return m


def install(with_fallbacks: bool = None) -> None:
  """Registers the overlay in sys.modules (idempotent)."""
  try:
    importlib.import_module("corenet")
    have_ref = True
  except ImportError:
    have_ref = False
  table = dict(OVERLAY)
  if with_fallbacks if with_fallbacks is not None else not have_ref:
    table.update(FALLBACK)
  for ref_name, mine in table.items():
    mod = importlib.import_module(mine)
    parent, _, leaf = ref_name.rpartition(".")
    pkg = _package(parent)
    sys.modules[ref_name] = mod
    setattr(pkg, leaf, mod)
