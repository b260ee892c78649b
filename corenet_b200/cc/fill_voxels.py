"""Drop-in for the reference's `corenet.cc.fill_voxels` (boundary row a8/a9).

Same public names and conventions as src/corenet/cc/fill_voxels.py:61-107 +
cc/module.cc:18-29 of the reference: `get_module()` returns an object exposing
`fill_inside_voxels_gpu(grid, inplace=False)` / `fill_inside_voxels_cpu(grid)`;
rank/device errors are `ValueError`; `inplace=True` returns the same tensor.
The implementation is the bit-packed flood fill in csrc/fill_voxels.cu (no JIT
compile: the C-ABI library is prebuilt by `python -m corenet_b200.build`).
"""
import torch as t

from corenet_b200 import _lib
from corenet_b200 import ops


class _Module:
  """Stands in for the pybind11 extension module `corenet_cpp`."""

  @staticmethod
  def fill_inside_voxels_gpu(grid: t.Tensor, inplace: bool = False) -> t.Tensor:
    if not grid.is_cuda:
      raise ValueError("Only CUDA tensors are supported by this OP")
    return ops.fill_inside_voxels(grid, inplace)

  @staticmethod
  def fill_inside_voxels_cpu(grid: t.Tensor) -> t.Tensor:
    """CPU tensor in, CPU tensor out -- computed on the GPU (this package has no
    CPU arithmetic path; it raises if no CUDA device is present)."""
    if grid.is_cuda:
      raise ValueError("Only CPU tensors are supported currently")
    if grid.dim() != 4:
      raise ValueError("Expecting rank 4 tensor")
    if not t.cuda.is_available():
      raise RuntimeError("corenet_b200 has no CPU fallback: fill_inside_voxels_cpu stages through cuda:0")
    return ops.fill_inside_voxels(grid.cuda(), True).cpu()


def get_module(verbose=False):
  _ = verbose
  _lib.lib()      # raises if the prebuilt library is missing
  return _Module


def fill_inside_voxels_cpu(grid: t.Tensor) -> t.Tensor:
  return get_module().fill_inside_voxels_cpu(grid)


def fill_inside_voxels_gpu(grid: t.Tensor, inplace=False) -> t.Tensor:
  return get_module().fill_inside_voxels_gpu(grid, inplace)
