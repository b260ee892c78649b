"""Builds libcorenet_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: python -m corenet_b200.build [--force] [--verbose]
The library has no torch / Python dependency (static cudart); the host side
binds it with ctypes (corenet_b200/_lib.py).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(OUT_DIR, "libcorenet_b200.so")
STAMP = os.path.join(OUT_DIR, "build.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
] + ([] if os.environ.get("CRN_NO_DIAG") == "1" else ["-DCRN_DIAG"])   # diagnostics: include/corenet_b200_diag.h


def sources():
  return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
  h = hashlib.sha256()
  for f in sorted(os.listdir(CSRC)) + ["../../include/corenet_b200.h", "../../include/corenet_b200_diag.h"]:
    p = os.path.join(CSRC, f)
    if os.path.isfile(p):
      h.update(f.encode())
      h.update(open(p, "rb").read())
  h.update(" ".join(NVCC_FLAGS).encode())
  return h.hexdigest()


def nvcc_path():
  cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
  p = os.path.join(cuda_home, "bin", "nvcc")
  return p if os.path.exists(p) else "nvcc"


def build(force: bool = False, verbose: bool = False) -> str:
  os.makedirs(OUT_DIR, exist_ok=True)
  digest = _digest()
  if (not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP)
      and open(STAMP).read().strip() == digest):
    return LIB_PATH
  objs = []
  procs = []
  for src in sources():
    obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".o")
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs.append(obj)
  failed = False
  for src, p in procs:
    out, _ = p.communicate()
    if p.returncode != 0 or verbose:
      sys.stderr.write(f"--- {os.path.basename(src)}\n{out}\n")
    failed |= p.returncode != 0
  if failed:
    raise RuntimeError("nvcc failed")
  cmd = [nvcc_path(), "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
  subprocess.check_call(cmd)
  with open(STAMP, "w") as f:
    f.write(digest)
  return LIB_PATH


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
