"""Stages the UNMODIFIED reference for runs on the GPU box (which has no /root/reference).

    python baseline/stage_ref.py          # build container only

  baseline/_ref/src/corenet/   a verbatim copy of /root/reference/src/corenet (python modules, shaders, cc sources)
  baseline/_ref/corenet_cpp/corenet_cpp.so
                               the reference's OWN native op -- cc/fill_voxels_cpu.cc, cc/fill_voxels_gpu.cu (kernels
                               K1/K2, fill_voxels_gpu.cu:96-132) and cc/module.cc -- compiled for sm_100a from the
                               sources where they lie, in the layout `CORENET_PRECOMPILED_CPP_MODULE_PATH` expects
                               (cc/fill_voxels.py:75-81).  Recipe: SURVEY.md Appendix A.2.

baseline/_ref/ is git-ignored (never part of the history) but not gpurun-ignored, so it travels to the GPU box,
where bench.py's `--impl reference-gpu` arm, the `fill` workload's baseline leg and the boundary tests in
tests/test_gpu_boundary.py import it.  Nothing under corenet_b200/ imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/corenet"
OUT = os.path.join(HERE, "_ref")
SRC_OUT = os.path.join(OUT, "src", "corenet")
EXT_DIR = os.path.join(OUT, "corenet_cpp")
SHIMS = os.path.join(os.path.dirname(HERE), "oracle", "ref_shims")


def ext_path():
  return os.path.join(EXT_DIR, "corenet_cpp.so")


def stage(verbose=False):
  if not os.path.isdir(REF):
    return None
  if os.path.isdir(SRC_OUT):
    shutil.rmtree(SRC_OUT)
  shutil.copytree(REF, SRC_OUT, ignore=shutil.ignore_patterns("__pycache__"))
  cc = os.path.join(REF, "cc")
  srcs = [os.path.join(cc, f) for f in ("fill_voxels_cpu.cc", "fill_voxels_gpu.cu", "module.cc")]
  if os.path.exists(ext_path()) and all(os.path.getmtime(ext_path()) > os.path.getmtime(s) for s in srcs):
    return OUT
  os.makedirs(EXT_DIR, exist_ok=True)
  os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
  from torch.utils import cpp_extension
  cpp_extension.load(
      name="corenet_cpp", sources=srcs, build_directory=EXT_DIR, with_cuda=True, verbose=verbose,
      is_python_module=False,
      extra_cflags=["-std=c++17", "-O2", "-DAT_PARALLEL_OPENMP", "-fopenmp", "-I" + SHIMS, "-include",
                    os.path.join(SHIMS, "compat.h")],
      extra_cuda_cflags=["-O2", "-I" + SHIMS, "-include", os.path.join(SHIMS, "compat.h")])
  return OUT


if __name__ == "__main__":
  print(stage(verbose="--verbose" in sys.argv))
