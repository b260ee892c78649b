"""Builds oracle/_ref/: the REFERENCE's own fill_voxels_cpu.cc, compiled where it
lies under /root/reference, for validating the oracle and as a CPU baseline.

    python oracle/build_ref.py            # container only (needs /root/reference)

Recipe: g++ on the reference source + a 10-line pybind shim that
binds only the CPU entry point, a `small_vector` shim header (boost is absent)
and a CHECK_EQ compat header.  Output goes to oracle/_ref/ (git-ignored, but it
travels to the GPU box).  No reference source is copied into the repo.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/corenet/cc/fill_voxels_cpu.cc"
OUT = os.path.join(HERE, "_ref")
NAME = "corenet_ref_cpu"


def so_path():
  return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
  if not os.path.exists(REF_SRC):
    return None
  if os.path.exists(so_path()) and os.path.getmtime(so_path()) > os.path.getmtime(REF_SRC):
    return so_path()
  os.makedirs(OUT, exist_ok=True)
  from torch.utils import cpp_extension
  shims = os.path.join(HERE, "ref_shims")
  cpp_extension.load(
      name=NAME, sources=[REF_SRC, os.path.join(shims, "module_cpu.cc")], build_directory=OUT,
      extra_cflags=["-std=c++17", "-O2", "-DAT_PARALLEL_OPENMP", "-fopenmp", "-I" + shims, "-include",
                    os.path.join(shims, "compat.h")],
      with_cuda=False, verbose=verbose)
  return so_path()


def load():
  """Imports the prebuilt reference module (None if it was never built)."""
  if not os.path.exists(so_path()):
    return None
  import importlib.util
  import torch  # noqa: F401  (libtorch must be loaded first)
  spec = importlib.util.spec_from_file_location(NAME, so_path())
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


if __name__ == "__main__":
  print(build(verbose="--verbose" in sys.argv))
