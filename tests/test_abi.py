"""The C-ABI library loads on a CPU-only box and exports every symbol include/corenet_b200.h declares."""
import ctypes
import os
import re

from tests.conftest import ROOT


def header_symbols(names=("corenet_b200.h", "corenet_b200_diag.h")):
  """Entry points declared by the boundary header and by the diagnostics header (built with -DCRN_DIAG)."""
  out = set()
  for name in names:
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out |= set(re.findall(r"\b(crn_[a-z0-9_]+)\s*\(", src))
  return sorted(out)


def test_library_exports_every_declared_symbol():
  from corenet_b200 import build
  path = build.build()
  lib = ctypes.CDLL(path)
  syms = header_symbols()
  assert len(syms) >= 30
  for s in syms:
    assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_ctypes_signatures_cover_the_header():
  from corenet_b200 import _lib
  assert sorted(_lib.EXPORTS) == header_symbols()
  lib = _lib.lib()
  assert lib.crn_build_arch() == b"sm_100a"
  assert lib.crn_version() >= 100
  # pure host-side query, no GPU needed: 2 bit-planes of ceil(W/32) words per row
  assert lib.crn_fill_workspace_bytes(2, 128, 128, 128) == 2 * 2 * 128 * 128 * 4 * 4


def test_sass_is_sm100a():
  import subprocess
  from corenet_b200 import build
  out = subprocess.run(["cuobjdump", "-lelf", build.LIB_PATH], capture_output=True, text=True).stdout
  assert "sm_100a" in out


def test_diagnostics_are_not_in_the_boundary_header():
  boundary = header_symbols(("corenet_b200.h",))
  assert not [s for s in boundary if "probe" in s or "debug" in s]
  assert {"crn_tc_probe", "crn_tc_probe_mn", "crn_gemm_tc_debug_read"} <= set(header_symbols(("corenet_b200_diag.h",)))
