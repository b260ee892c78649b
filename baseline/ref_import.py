"""Imports the UNMODIFIED reference package `corenet` (from the staged copy baseline/_ref/src -- written by
baseline/stage_ref.py in the build container, travels to the GPU box -- else from /root/reference/src) under torch 2.11.
TEST / BASELINE INFRASTRUCTURE ONLY.

The reference pins packages that are absent from this image (json5, jq, google-cloud-storage, moderngl, tensorboard,
dataclasses_jsonschema ...): none of them is on the model path, so they are replaced by inert stub modules created on
demand by a meta-path finder restricted to that list.
"""
import dataclasses
import importlib.abc
import importlib.machinery
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STUB_ROOTS = ("json5", "jq", "google", "moderngl", "glcontext", "OpenGL", "skimage", "tensorboard", "tensorflow",
              "dataclasses_jsonschema", "tqdm", "PIL", "cv2", "pandas_stub_never")


class _Stub(types.ModuleType):
  __path__ = []

  def __getattr__(self, name):
    if name.startswith("__"):
      raise AttributeError(name)
    v = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})
    setattr(self, name, v)
    return v


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
  def find_spec(self, fullname, path=None, target=None):
    if fullname.split(".")[0] in STUB_ROOTS:
      return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
    return None

  def create_module(self, spec):
    return _Stub(spec.name)

  def exec_module(self, module):
    pass


def reference_src():
  # the staged copy first: bench.py and the GPU tests must not depend on /root/reference (absent on the GPU box);
  # the checkout itself only serves the CPU tests of a build container in which nothing has been staged yet
  for p in (os.path.join(HERE, "_ref", "src"), "/root/reference/src"):
    if os.path.isdir(os.path.join(p, "corenet")):
      return p
  return None


def import_reference():
  """Returns the reference's top-level `corenet` package, or None when no copy of it is present."""
  src = reference_src()
  if src is None:
    return None
  if "corenet" in sys.modules and getattr(sys.modules["corenet"], "__file__", "").startswith(src):
    return sys.modules["corenet"]
  for k in [k for k in sys.modules if k == "corenet" or k.startswith("corenet.")]:
    del sys.modules[k]
  js = types.ModuleType("dataclasses_jsonschema")

  class JsonSchemaMixin:
    def to_dict(self):
      return dataclasses.asdict(self)

    @classmethod
    def from_dict(cls, d):
      kw = {}
      for f in dataclasses.fields(cls):
        v = d[f.name]
        ft = f.type if isinstance(f.type, type) else None
        if ft is not None and dataclasses.is_dataclass(ft) and isinstance(v, dict):
          v = ft.from_dict(v)
        elif isinstance(v, list):
          v = tuple(v)
        kw[f.name] = v
      return cls(**kw)
  js.JsonSchemaMixin = JsonSchemaMixin
  sys.modules["dataclasses_jsonschema"] = js
  try:
    import torch.utils.tensorboard  # noqa: F401
  except Exception:
    sys.modules["torch.utils.tensorboard"] = _Stub("torch.utils.tensorboard")
    import torch.utils
    torch.utils.tensorboard = sys.modules["torch.utils.tensorboard"]
  if not any(isinstance(f, _Finder) for f in sys.meta_path):
    sys.meta_path.append(_Finder())
  sys.path.insert(0, src)
  import corenet
  return corenet


def reference_native_module():
  """The reference's own compiled extension (fill_inside_voxels_{cpu,gpu}) staged by baseline/stage_ref.py."""
  import importlib.util
  import torch  # noqa: F401
  p = os.path.join(HERE, "_ref", "corenet_cpp", "corenet_cpp.so")
  if not os.path.exists(p):
    return None
  spec = importlib.util.spec_from_file_location("corenet_cpp", p)
  m = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(m)
  return m
