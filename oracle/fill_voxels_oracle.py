"""Python face of the C fill oracle (test infrastructure only).

fill_inside_voxels_oracle(grid) reproduces
fill_inside_voxels_cpu (/root/reference/src/corenet/cc/fill_voxels_cpu.cc:158-183):
voxels of regions not connected to the outside become 1, everything else keeps
its value.  `fill_inside_scipy` is an independent restatement of the same rule
 used to cross-check the C code.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fill_voxels_oracle.c")
LIB = os.path.join(HERE, "_build", "libfill_oracle.so")


def build():
  os.makedirs(os.path.dirname(LIB), exist_ok=True)
  if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, SRC])
  return LIB


_lib = None


def _get():
  global _lib
  if _lib is None:
    _lib = ctypes.CDLL(build())
    _lib.fill_oracle.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4
    _lib.fill_oracle.restype = ctypes.c_int
  return _lib


def fill_mask_oracle(grid: np.ndarray) -> np.ndarray:
  """uint8[N,D,H,W] mask of the voxels the reference sets to 1."""
  assert grid.ndim == 4
  occ = np.ascontiguousarray(grid > 0).astype(np.uint8)
  out = np.empty_like(occ)
  n, d, h, w = occ.shape
  if occ.size:
    rc = _get().fill_oracle(occ.ctypes.data, out.ctypes.data, n, d, h, w)
    assert rc == 0
  return out


def fill_inside_voxels_oracle(grid: np.ndarray) -> np.ndarray:
  """Same dtype/shape as grid; mask voxels are 1, the rest untouched (CPU reference semantics)."""
  out = np.array(grid, copy=True)
  out[fill_mask_oracle(grid).astype(bool)] = 1
  return out


def fill_inside_scipy(grid: np.ndarray) -> np.ndarray:
  """Independent restatement: 6-connected empty components not touching z=0 / y=0 / x=0 are filled."""
  from scipy import ndimage
  out = np.array(grid, copy=True)
  for b in range(grid.shape[0]):
    lab, _ = ndimage.label(grid[b] <= 0)
    outside = (set(np.unique(lab[0])) | set(np.unique(lab[:, 0])) | set(np.unique(lab[:, :, 0]))) - {0}
    filled = (grid[b] > 0) | ~np.isin(lab, list(outside))
    out[b][filled & ~(grid[b] > 0)] = 1
    out[b][(grid[b] > 0)] = 1 if False else out[b][(grid[b] > 0)]
  return out
