"""Host-side kernel routing (no GPU): the `*_supported` predicates and packed-size queries of the C-ABI are pure
host functions, so the layer -> kernel table of DESIGN.md section 4 is checked on the CPU box."""
import ctypes as C

import pytest

from corenet_b200 import _lib
from corenet_b200._lib import ConvDesc


def desc(n, cin, cout, idims, odims, k, stride, pad, transposed, x_cs=None, y_cs=None):
  d = ConvDesc()
  d.N, d.Cin, d.Cout = n, cin, cout
  d.iD, d.iH, d.iW = idims
  d.oD, d.oH, d.oW = odims
  d.kD, d.kH, d.kW = k
  d.stride, d.pad, d.transposed = stride, pad, int(transposed)
  r4 = lambda c: (c + 3) // 4 * 4
  d.x_cs, d.x_co, d.y_cs, d.y_co = x_cs or r4(cin), 0, y_cs or r4(cout), 0
  d.CinP, d.CoutP, d.y_planar, d.bias_n_stride = r4(cin), r4(cout), 0, 0
  return d


def cube(g):
  return (g, g, g)


# decoder layers at the reference configuration (B = 4): (cin, cout, grid)
C1 = {"stage_6.c1": (28, 16, 64), "stage_5.c1": (56, 32, 32), "stage_4.c1": (112, 64, 16), "stage_3.c1": (224, 128, 8)}
T1 = {"stage_6.t1": (16, 2, 64), "stage_5.t1": (32, 16, 32), "stage_4.t1": (64, 32, 16), "stage_3.t1": (128, 64, 8)}


def test_k5_weight_gradient_routing():
  lib = _lib.lib()
  for name, (cin, cout, g) in C1.items():
    d = desc(4, cin, cout, cube(g), cube(g), (5, 5, 5), 1, 2, False)
    line = lib.crn_conv_wgrad_line_supported(C.byref(d))
    xline = lib.crn_conv_wgrad_xline_supported(C.byref(d))
    # narrow layers: tap-stacked line kernel; wide coarse layers: x-line kernel; never both
    assert (line, xline) == {"stage_6.c1": (1, 0), "stage_5.c1": (1, 0), "stage_4.c1": (0, 1), "stage_3.c1": (0, 1)}[name], name


def test_transposed_k7_weight_gradient_routing():
  lib = _lib.lib()
  want = {"stage_6.t1": 2, "stage_5.t1": 1, "stage_4.t1": 0, "stage_3.t1": 0}
  for name, (cin, cout, g) in T1.items():
    y_cs = 4 if name == "stage_6.t1" else None
    d = desc(4, cin, cout, cube(g), cube(2 * g), (7, 7, 7), 2, 3, True, y_cs=y_cs)
    assert lib.crn_convt7_wgrad_line_supported(C.byref(d)) == want[name], name
  # stage_4.t1 decomposes into 2 x 2 (32-channel Cin, 16-channel Cout) blocks that the class-channel kernel takes
  d = desc(4, 32, 16, cube(16), cube(32), (7, 7, 7), 2, 3, True, x_cs=64, y_cs=56)
  assert lib.crn_convt7_wgrad_line_supported(C.byref(d)) == 1
  # the semantic head (15 classes) stays on the FFMA kernel
  d = desc(4, 16, 15, cube(64), cube(128), (7, 7, 7), 2, 3, True, y_cs=16)
  assert lib.crn_convt7_wgrad_line_supported(C.byref(d)) == 0


def test_packed_weight_sizes():
  lib = _lib.lib()
  # implicit GEMM: tiles of 64 / 128 output channels x stages of 16 input channels, hi + lo
  assert lib.crn_gemm_tc_packed_floats(64, 256, 1) == 2 * 4 * (2 * 4 * 128 * 4)
  assert lib.crn_gemm_tc_packed_floats(512, 512, 9) == 4 * 9 * 32 * (2 * 4 * 128 * 4)
  assert lib.crn_gemm_tc_packed_floats(112, 64, 125) == 1 * 125 * 7 * (2 * 4 * 64 * 4)
  # kz-stacked k5: per 8-channel pass 5 ky rows of 5 taps x 10 KB
  assert lib.crn_tc5s_packed_floats(28) == 4 * 5 * (5 * 10240 // 4)
  assert lib.crn_tc5s_packed_floats(16) == 2 * 5 * (5 * 10240 // 4)


@pytest.mark.parametrize("bad", ["stride", "kernel", "width"])
def test_unsupported_shapes_are_rejected(bad):
  lib = _lib.lib()
  d = desc(1, 28, 16, cube(64), cube(64), (5, 5, 5), 1, 2, False)
  if bad == "stride":
    d.stride = 2
  elif bad == "kernel":
    d.kD = 3
  else:
    d.iW = d.oW = 48
  assert lib.crn_conv_wgrad_line_supported(C.byref(d)) == 0
  assert lib.crn_conv_wgrad_xline_supported(C.byref(d)) == 0
