"""Per-kernel parity on a real B200: every C-ABI op against the CPU oracle.

Integer / index / data-movement ops: bit-exact.  fp32 arithmetic ops: relative
tolerances written next to each assert (the kernels use fp32 FFMA; only the
summation order differs from the oracle).
"""
import zlib

import numpy as np
import pytest
import torch as t
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import corenet_oracle as O
from oracle import fill_voxels_oracle as FO
from oracle import voxelize_oracle as VO
from tests.conftest import cube_mesh, fill_test_grids
from tests.test_oracle_losses import GT, LOGITS


def dev():
  return t.device("cuda", 0)


def rel_err(a: t.Tensor, b: t.Tensor) -> float:
  """max|a-b| / max|b| (SURVEY 8d parity metric)."""
  a, b = a.detach().double().cpu(), b.detach().double().cpu()
  return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


# ---------------------------------------------------------------------------- convolutions
CONV_CASES = [
    # (name, x shape NC.., weight shape, stride, pad, transposed, output_padding)
    ("1x1", (2, 64, 16, 16), (96, 64, 1, 1), 1, 0, False, 0),
    ("1x1_s2", (2, 64, 16, 16), (128, 64, 1, 1), 2, 0, False, 0),
    ("3x3", (2, 32, 14, 14), (48, 32, 3, 3), 1, 1, False, 0),
    ("stem7x7_s2_cin3", (2, 3, 40, 40), (64, 3, 7, 7), 2, 3, False, 0),
    ("stem7x7_s2_ow64_stemkernel", (2, 3, 128, 128), (64, 3, 7, 7), 2, 3, False, 0),
    ("stem7x7_s2_ow128_stemkernel", (1, 3, 256, 256), (64, 3, 7, 7), 2, 3, False, 0),
    ("small_cout", (1, 28, 12, 12, 12), (16, 28, 5, 5, 5), 1, 2, False, 0),
    ("3d_k3", (2, 16, 6, 6, 6), (24, 16, 3, 3, 3), 1, 1, False, 0),
    ("3d_k5_wide", (1, 56, 8, 8, 8), (132, 56, 5, 5, 5), 1, 2, False, 0),
    ("convT_k7_s2", (1, 16, 6, 6, 6), (16, 12, 7, 7, 7), 2, 3, True, 1),
    ("convT_k3_s2", (2, 32, 4, 4, 4), (32, 20, 3, 3, 3), 2, 1, True, 1),
    ("convT_k4_s4_latent", (3, 67, 1, 1, 1), (67, 40, 4, 4, 4), 4, 0, True, 0),
    ("convT_cout2", (1, 16, 8, 8, 8), (16, 2, 7, 7, 7), 2, 3, True, 1),
    ("convT_k7_s2_rowkernel", (2, 32, 9, 8, 10), (32, 16, 7, 7, 7), 2, 3, True, 1),
    ("conv_k5_rowkernel_56_32", (2, 56, 7, 9, 11), (32, 56, 5, 5, 5), 1, 2, False, 0),
    ("conv_k5_rowkernel_112_64", (1, 112, 6, 6, 16), (64, 112, 5, 5, 5), 1, 2, False, 0),
    # shapes inside the row-direct envelope (x extent a multiple of 8, <= 64 channels)
    ("rd_conv_k5_28_16", (1, 28, 6, 9, 16), (16, 28, 5, 5, 5), 1, 2, False, 0),
    ("rd_conv_k5_56_32", (2, 56, 5, 7, 32), (32, 56, 5, 5, 5), 1, 2, False, 0),
    ("rd_conv_k5_16_3", (1, 16, 5, 6, 8), (3, 16, 5, 5, 5), 1, 2, False, 0),
    ("rd_convT_k7_32_16", (2, 32, 5, 6, 8), (32, 16, 7, 7, 7), 2, 3, True, 1),
    ("rd_convT_k7_16_2", (1, 16, 6, 9, 32), (16, 2, 7, 7, 7), 2, 3, True, 1),
    ("rd_convT_k7_16_15", (1, 16, 4, 5, 8), (16, 15, 7, 7, 7), 2, 3, True, 1),
    ("rd_convT_k7_64_32", (1, 64, 4, 4, 16), (64, 32, 7, 7, 7), 2, 3, True, 1),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_fwd_bwd(case):
  from corenet_b200 import ops
  name, xs, ws, stride, pad, transposed, opad = case
  g = t.Generator().manual_seed(zlib.crc32(name.encode()) % 1000)
  x = t.randn(xs, generator=g)
  w = t.randn(ws, generator=g) * 0.1
  cout = ws[1] if transposed else ws[0]
  b = t.randn(cout, generator=g)
  nd = len(xs) - 2
  x_o, w_o, b_o = [v.clone().requires_grad_(True) for v in (x, w, b)]
  if transposed:
    y_o = F.conv_transpose3d(x_o, w_o, b_o, stride=stride, padding=pad, output_padding=opad)
  else:
    y_o = (F.conv2d if nd == 2 else F.conv3d)(x_o, w_o, b_o, stride=stride, padding=pad)
  gy = t.randn(y_o.shape, generator=g)
  y_o.backward(gy)
  x_c, w_c, b_c = [v.clone().to(dev()).requires_grad_(True) for v in (x, w, b)]
  if transposed:
    y_c = ops.conv_transpose(x_c, w_c, b_c, stride=stride, padding=pad, output_padding=opad)
  else:
    y_c = ops.conv(x_c, w_c, b_c, stride=stride, padding=pad)
  assert tuple(y_c.shape) == tuple(y_o.shape)
  y_c.backward(gy.to(dev()))
  # fp32 FFMA vs fp32 oneDNN: only the summation order differs -> 2e-5 of the tensor's max
  assert rel_err(y_c, y_o) < 2e-5, "fwd"
  assert rel_err(x_c.grad, x_o.grad) < 2e-5, "dgrad"
  assert rel_err(w_c.grad, w_o.grad) < 5e-5, "wgrad"
  assert rel_err(b_c.grad, b_o.grad) < 2e-5, "bias grad"


@pytest.mark.parametrize("n,cin,cout,dhw", [(1, 8, 16, (8, 16, 8)), (2, 28, 16, (16, 32, 16)), (1, 56, 32, (8, 16, 16)),
                                              (1, 16, 12, (8, 16, 24)),
                                              (3, 8, 40, (16, 32, 32)),      # 96 four-plane items (N <= 64 kernel)
                                              (1, 112, 64, (16, 16, 16))])   # stage_4.c1: one plane per item, fwd only
@pytest.mark.parametrize("kind", [0, 1], ids=["fwd", "dgrad"])
def test_conv5_tcgen05(n, cin, cout, dhw, kind):
  """Conv3d k=5 on the tcgen05 tensor cores (3xTF32 + fp32 TMEM accumulation) against the fp64 oracle.
  Tolerance 2e-4 of the tensor max: the tensor core's fp32 accumulator truncates (measured ~3e-5 over 375
  accumulations), operands are exact to ~2^-21."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  if kind == 1 and cin > 64:
    pytest.skip("the plane kernel holds at most 64 output channels (dgrad: N = Cin)")
  d, h, w = dhw
  g = t.Generator().manual_seed(cin * 7 + cout + kind)
  wt = t.randn(cout, cin, 5, 5, 5, generator=g) * 0.05
  bias = t.randn(cout, generator=g)
  r4 = lambda c: (c + 3) // 4 * 4
  if kind == 0:
    x = t.randn(n, cin, d, h, w, generator=g)
    ref = F.conv3d(x.double(), wt.double(), bias.double(), padding=2)
    K, N = cin, cout
  else:
    x = t.randn(n, cout, d, h, w, generator=g)
    ref = F.conv_transpose3d(x.double(), wt.double(), None, padding=2)
    K, N = cout, cin
  xin = t.zeros(n * d * h * w, r4(K), device=dev())
  xin[:, :K] = x.permute(0, 2, 3, 4, 1).reshape(-1, K).to(dev())
  out = t.full((n * d * h * w, r4(N)), float("nan"), device=dev())
  wtc = t.zeros(_lib.lib().crn_tc5_packed_floats(K, N), device=dev())
  st = _lib.stream_ptr()
  _lib.call("crn_tc5_pack", wt.to(dev()).contiguous().data_ptr(), cout, cin, kind, wtc.data_ptr(), st)
  desc = ops.make_desc(n, cin, cout, (d, h, w), (d, h, w), (5, 5, 5), 1, 2, False, r4(cin), r4(cout))
  status = t.zeros(1, dtype=t.int32, device=dev())
  b = bias.to(dev())
  _lib.call("crn_conv5_tc", C.byref(desc), kind, xin.data_ptr(), wtc.data_ptr(), b.data_ptr(), out.data_ptr(),
            status.data_ptr(), st)
  t.cuda.synchronize()
  assert int(status) == 0
  got = out[:, :N].reshape(n, d, h, w, N).permute(0, 4, 1, 2, 3)
  assert rel_err(got, ref) < 2e-4


GEMM_TC_CASES = [
    # (name, x shape, weight shape, stride, pad)
    ("enc_1x1", (2, 64, 16, 16), (256, 64, 1, 1), 1, 0),
    ("enc_1x1_s2", (2, 256, 16, 16), (128, 256, 1, 1), 2, 0),
    ("enc_3x3", (2, 128, 14, 14), (128, 128, 3, 3), 1, 1),
    ("enc_3x3_splitk", (1, 512, 8, 8), (512, 512, 3, 3), 1, 1),
    ("enc_1x1_k2048", (1, 2048, 8, 8), (512, 2048, 1, 1), 1, 0),
    ("dec_k5_112_64", (1, 112, 6, 7, 9), (64, 112, 5, 5, 5), 1, 2),
    ("dec_k5_224_128", (1, 224, 4, 4, 4), (128, 224, 5, 5, 5), 1, 2),
    ("odd_channels_36_40", (1, 36, 5, 6, 7), (40, 36, 3, 3, 3), 1, 1),
]


@pytest.mark.parametrize("case", GEMM_TC_CASES, ids=[c[0] for c in GEMM_TC_CASES])
@pytest.mark.parametrize("kind", [0, 1], ids=["fwd", "dgrad"])
def test_conv_gemm_tcgen05(case, kind):
  """Implicit-GEMM conv on tcgen05 (3xTF32, accumulator flushed every 72 MMAs) against the fp64 oracle.
  Tolerance 2e-5 of the tensor max, the same as the fp32 FFMA kernels."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  name, xs, ws, stride, pad = case
  g = t.Generator().manual_seed(zlib.crc32(name.encode()) % 1000 + kind)
  nd = len(xs) - 2
  conv = F.conv2d if nd == 2 else F.conv3d
  convt = F.conv_transpose2d if nd == 2 else F.conv_transpose3d
  cout, cin = ws[0], ws[1]
  wt = t.randn(ws, generator=g) * 0.05
  bias = t.randn(cout, generator=g)
  x = t.randn(xs, generator=g)
  y_ref = conv(x.double(), wt.double(), bias.double(), stride=stride, padding=pad)
  if kind == 0:
    src, ref, K, N = x, y_ref, cin, cout
  else:
    src = t.randn(y_ref.shape, generator=g)
    if stride == 1:
      ref = convt(src.double(), wt.double(), None, stride=1, padding=pad)
    else:                                        # strided dgrad: class-ordered transposed-gather mode
      xg = x.double().requires_grad_(True)
      conv(xg, wt.double(), None, stride=stride, padding=pad).backward(src.double())
      ref = xg.grad
    K, N = cout, cin
  perm = [0] + list(range(2, nd + 2)) + [1]
  inv = [0, nd + 1] + list(range(1, nd + 1))
  xin = src.permute(perm).reshape(-1, K).contiguous().to(dev())
  out_sp = [ref.shape[0]] + list(ref.shape[2:])
  rows_out = int(np.prod(out_sp))
  out = t.full((rows_out, N), float("nan"), device=dev())
  wtc = ops.gemm_tc_pack([wt.to(dev()).contiguous()], [kind])[0]
  d3 = lambda sp: tuple([1] * (3 - len(sp)) + list(sp))
  desc = ops.make_desc(xs[0], cin, cout, d3(xs[2:]), d3(y_ref.shape[2:]), d3(ws[2:]), stride, pad, False, cin, cout)
  status = t.zeros(1, dtype=t.int32, device=dev())
  b = t.zeros(cout + 1, device=dev())[1:]          # parameters are 4-byte aligned views of a flat buffer
  b.copy_(bias)
  _lib.call("crn_conv_gemm_tc", C.byref(desc), kind, xin.data_ptr(), wtc.data_ptr(), b.data_ptr(), out.data_ptr(), 0,
            status.data_ptr(), _lib.stream_ptr())
  t.cuda.synchronize()
  assert int(status) == 0
  got = out.reshape(out_sp + [N]).permute(inv)
  assert rel_err(got, ref) < 2e-5
  # accumulate: out += conv
  _lib.call("crn_conv_gemm_tc", C.byref(desc), kind, xin.data_ptr(), wtc.data_ptr(), b.data_ptr(), out.data_ptr(), 1,
            status.data_ptr(), _lib.stream_ptr())
  t.cuda.synchronize()
  ref2 = 2 * ref - (bias.double().view([1, -1] + [1] * nd) if kind == 0 else 0)
  assert rel_err(out.reshape(out_sp + [N]).permute(inv), ref2) < 2e-5


@pytest.mark.parametrize("n,cin,cout,dhw,k,pad", [(1, 64, 32, (4, 5, 6), 7, 3), (2, 128, 64, (3, 4, 4), 7, 3),
                                                    (1, 256, 128, (4, 4, 4), 3, 1)])
def test_conv_transpose_dgrad_as_strided_conv_tcgen05(n, cin, cout, dhw, k, pad):
  """dgrad of ConvTranspose3d(stride 2) = stride-2 forward convolution of dy with the SAME weight tensor read as
  [Cout'=Cin][Cin'=Cout][taps] (engine._conv_dispatch): crn_conv_gemm_tc kind 0 against the fp64 oracle."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  g = t.Generator().manual_seed(cin + cout + k)
  wt = t.randn(cin, cout, k, k, k, generator=g) * 0.05
  x = t.randn((n, cin) + dhw, generator=g).double().requires_grad_(True)
  y = F.conv_transpose3d(x, wt.double(), None, stride=2, padding=pad, output_padding=1)
  gy = t.randn(y.shape, generator=g)
  y.backward(gy.double())
  fine = tuple(y.shape[2:])
  gr = gy.permute(0, 2, 3, 4, 1).reshape(-1, cout).contiguous().to(dev())
  dx = t.full((n * dhw[0] * dhw[1] * dhw[2], cin), float("nan"), device=dev())
  wtc = ops.gemm_tc_pack([wt.to(dev()).contiguous()], [0])[0]          # weight read as conv [Cout'=cin][Cin'=cout]
  desc = ops.make_desc(n, cout, cin, fine, dhw, (k, k, k), 2, pad, False, cout, cin)
  status = t.zeros(1, dtype=t.int32, device=dev())
  _lib.call("crn_conv_gemm_tc", C.byref(desc), 0, gr.data_ptr(), wtc.data_ptr(), None, dx.data_ptr(), 0,
            status.data_ptr(), _lib.stream_ptr())
  t.cuda.synchronize()
  assert int(status) == 0
  got = dx.reshape((n,) + dhw + (cin,)).permute(0, 4, 1, 2, 3)
  assert rel_err(got, x.grad) < 2e-5


@pytest.mark.parametrize("n,cin,cout,dhw,k,pad", [(1, 128, 64, (4, 4, 4), 7, 3), (2, 64, 32, (3, 5, 4), 7, 3),
                                                    (1, 256, 128, (4, 4, 4), 3, 1), (2, 40, 36, (2, 3, 5), 3, 1)])
def test_conv_transpose_fwd_gemm_tcgen05(n, cin, cout, dhw, k, pad):
  """ConvTranspose3d(stride 2) forward through the class-ordered transposed-gather mode of the implicit-GEMM tcgen05
  kernel (weights packed with flipped taps, [Cin][Cout][taps] passed as Cout' = Cin, Cin' = Cout) vs the fp64 oracle."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  g = t.Generator().manual_seed(cin + cout + k)
  wt = t.randn(cin, cout, k, k, k, generator=g) * 0.05
  bias = t.randn(cout, generator=g)
  x = t.randn((n, cin) + dhw, generator=g)
  ref = F.conv_transpose3d(x.double(), wt.double(), bias.double(), stride=2, padding=pad, output_padding=1)
  fine = tuple(ref.shape[2:])
  xr = x.permute(0, 2, 3, 4, 1).reshape(-1, cin).contiguous().to(dev())
  out = t.full((n * fine[0] * fine[1] * fine[2], cout), float("nan"), device=dev())
  wtc = ops.gemm_tc_pack([wt.to(dev()).contiguous()], [1])[0]          # shape[0] = Cin read as Cout', flipped taps
  desc = ops.make_desc(n, cin, cout, dhw, fine, (k, k, k), 2, pad, True, cin, cout)
  status = t.zeros(1, dtype=t.int32, device=dev())
  b = bias.to(dev())
  _lib.call("crn_conv_gemm_tc", C.byref(desc), 0, xr.data_ptr(), wtc.data_ptr(), b.data_ptr(), out.data_ptr(), 0,
            status.data_ptr(), _lib.stream_ptr())
  t.cuda.synchronize()
  assert int(status) == 0
  got = out.reshape((n,) + fine + (cout,)).permute(0, 4, 1, 2, 3)
  assert rel_err(got, ref) < 2e-5


WGRAD_TC_CASES = GEMM_TC_CASES + [
    ("convT_k7_s2_128_64", (1, 128, 4, 4, 4), (128, 64, 7, 7, 7), 2, 3),
    ("convT_k3_s2_32_20", (2, 32, 4, 5, 6), (32, 20, 3, 3, 3), 2, 1),
    ("rows_not_multiple_of_16", (1, 64, 5, 7), (96, 64, 3, 3), 1, 1),
]


@pytest.mark.parametrize("case", WGRAD_TC_CASES, ids=[c[0] for c in WGRAD_TC_CASES])
def test_conv_wgrad_tcgen05(case):
  """Weight gradient on tcgen05 (MN-major tf32 operands, 3xTF32) against the fp64 oracle; 5e-5 of the tensor max
  (the FFMA wgrad kernels' tolerance)."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  name, xs, ws, stride, pad = case
  transposed = name.startswith("convT")
  g = t.Generator().manual_seed(zlib.crc32(name.encode()) % 1000)
  nd = len(xs) - 2
  x = t.randn(xs, generator=g)
  wt = (t.randn(ws, generator=g) * 0.05).double().requires_grad_(True)
  if transposed:
    cin, cout = ws[0], ws[1]
    y = F.conv_transpose3d(x.double(), wt, None, stride=stride, padding=pad, output_padding=1)
  else:
    cout, cin = ws[0], ws[1]
    y = (F.conv2d if nd == 2 else F.conv3d)(x.double(), wt, None, stride=stride, padding=pad)
  gy = t.randn(y.shape, generator=g)
  y.backward(gy.double())
  perm = [0] + list(range(2, nd + 2)) + [1]
  xr = x.permute(perm).reshape(-1, cin).contiguous().to(dev())
  gr = gy.permute(perm).reshape(-1, cout).contiguous().to(dev())
  taps = int(np.prod(ws[2:]))
  dw = t.zeros(taps, cin, cout, device=dev())
  d3 = lambda sp: tuple([1] * (3 - len(sp)) + list(sp))
  desc = ops.make_desc(xs[0], cin, cout, d3(xs[2:]), d3(y.shape[2:]), d3(ws[2:]), stride, pad, transposed, cin, cout)
  status = t.zeros(1, dtype=t.int32, device=dev())
  _lib.call("crn_conv_wgrad_tc", C.byref(desc), xr.data_ptr(), gr.data_ptr(), dw.data_ptr(), status.data_ptr(),
            _lib.stream_ptr())
  t.cuda.synchronize()
  assert int(status) == 0
  gw = wt.grad.reshape(ws[0], ws[1], taps)
  ref = gw.permute(2, 0, 1) if transposed else gw.permute(2, 1, 0)     # [tap][ci][co]
  assert rel_err(dw, ref) < 5e-5


@pytest.mark.parametrize("n,cin,cout,dhw", [(1, 28, 16, (5, 6, 64)), (2, 28, 16, (3, 70, 64)), (1, 12, 8, (4, 5, 32)),
                                              (2, 56, 32, (4, 6, 32)), (1, 36, 20, (3, 40, 32))])
def test_conv5_wgrad_line_tcgen05(n, cin, cout, dhw):
  """Narrow Conv3d k=5 weight gradient with filter taps stacked into the MMA tile (MN-major tf32, all four hi/lo
  products) against the fp64 oracle; 5e-5 of the tensor max like the FFMA wgrad kernels."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  d, h, w = dhw
  g = t.Generator().manual_seed(cin * 13 + cout + d)
  x = t.randn(n, cin, d, h, w, generator=g)
  wt = (t.randn(cout, cin, 5, 5, 5, generator=g) * 0.05).double().requires_grad_(True)
  y = F.conv3d(x.double(), wt, None, padding=2)
  gy = t.randn(y.shape, generator=g)
  y.backward(gy.double())
  xr = x.permute(0, 2, 3, 4, 1).reshape(-1, cin).contiguous().to(dev())
  gr = gy.permute(0, 2, 3, 4, 1).reshape(-1, cout).contiguous().to(dev())
  dw = t.zeros(125, cin, cout, device=dev())
  desc = ops.make_desc(n, cin, cout, dhw, dhw, (5, 5, 5), 1, 2, False, cin, cout)
  assert _lib.lib().crn_conv_wgrad_line_supported(C.byref(desc)) == 1
  status = t.zeros(1, dtype=t.int32, device=dev())
  _lib.call("crn_conv_wgrad_line", C.byref(desc), xr.data_ptr(), gr.data_ptr(), dw.data_ptr(), status.data_ptr(),
            _lib.stream_ptr())
  t.cuda.synchronize()
  assert int(status) == 0
  ref = wt.grad.reshape(cout, cin, 125).permute(2, 1, 0)
  assert rel_err(dw, ref) < 5e-5


@pytest.mark.parametrize("n,cin,cout,dhw", [(1, 112, 64, (4, 5, 16)), (2, 224, 128, (3, 4, 8)), (1, 64, 32, (5, 3, 16)),
                                              (1, 132, 68, (3, 3, 8))])
def test_conv5_wgrad_xline_tcgen05(n, cin, cout, dhw):
  """Wide coarse Conv3d k=5 weight gradient: one staged image row per step, 5 kx taps as row shifts (MN-major tf32,
  3xTF32) against the fp64 oracle; 5e-5 of the tensor max."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  g = t.Generator().manual_seed(cin + cout)
  x = t.randn((n, cin) + dhw, generator=g)
  wt = (t.randn(cout, cin, 5, 5, 5, generator=g) * 0.05).double().requires_grad_(True)
  y = F.conv3d(x.double(), wt, None, padding=2)
  gy = t.randn(y.shape, generator=g)
  y.backward(gy.double())
  xr = x.permute(0, 2, 3, 4, 1).reshape(-1, cin).contiguous().to(dev())
  gr = gy.permute(0, 2, 3, 4, 1).reshape(-1, cout).contiguous().to(dev())
  dw = t.zeros(125, cin, cout, device=dev())
  desc = ops.make_desc(n, cin, cout, dhw, dhw, (5, 5, 5), 1, 2, False, cin, cout)
  assert _lib.lib().crn_conv_wgrad_xline_supported(C.byref(desc)) == 1
  status = t.zeros(1, dtype=t.int32, device=dev())
  _lib.call("crn_conv_wgrad_xline", C.byref(desc), xr.data_ptr(), gr.data_ptr(), dw.data_ptr(), status.data_ptr(),
            _lib.stream_ptr())
  t.cuda.synchronize()
  assert int(status) == 0
  ref = wt.grad.reshape(cout, cin, 125).permute(2, 1, 0)
  assert rel_err(dw, ref) < 5e-5


@pytest.mark.parametrize("n,cin,cout,dhw,ycs", [(1, 32, 16, (4, 5, 32), 16), (2, 32, 16, (3, 37, 32), 28),
                                                  (1, 20, 16, (5, 4, 16), 16), (1, 16, 2, (4, 5, 64), 4),
                                                  (2, 12, 3, (3, 6, 32), 4)])
def test_convt7_wgrad_line_tcgen05(n, cin, cout, dhw, ycs):
  """ConvTranspose3d k=7 s=2 weight gradient through the class-channel tap-stacked tcgen05 kernels (Cout = 16, and the
  narrow logits layer with Cout <= 4) against the fp64 oracle; dy may sit inside a wider concat buffer (stride ycs)."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  g = t.Generator().manual_seed(cin + dhw[1])
  x = t.randn((n, cin) + dhw, generator=g)
  wt = (t.randn(cin, cout, 7, 7, 7, generator=g) * 0.05).double().requires_grad_(True)
  y = F.conv_transpose3d(x.double(), wt, None, stride=2, padding=3, output_padding=1)
  gy = t.randn(y.shape, generator=g)
  y.backward(gy.double())
  fine = tuple(y.shape[2:])
  xr = x.permute(0, 2, 3, 4, 1).reshape(-1, cin).contiguous().to(dev())
  if cout == 16:
    gr = t.randn(gy.numel() // cout, ycs, generator=g)                    # other channels of the concat buffer: noise
  else:
    gr = t.zeros(gy.numel() // cout, ycs)                                 # padded logits-gradient rows
  gr[:, :cout] = gy.permute(0, 2, 3, 4, 1).reshape(-1, cout)
  gr = gr.to(dev())
  coutp = (cout + 3) // 4 * 4
  dw = t.zeros(343, cin, coutp, device=dev())
  desc = ops.make_desc(n, cin, cout, dhw, fine, (7, 7, 7), 2, 3, True, cin, ycs)
  assert _lib.lib().crn_convt7_wgrad_line_supported(C.byref(desc)) == (1 if cout == 16 else 2)
  status = t.zeros(1, dtype=t.int32, device=dev())
  _lib.call("crn_convt7_wgrad_line", C.byref(desc), xr.data_ptr(), gr.data_ptr(), dw.data_ptr(), status.data_ptr(),
            _lib.stream_ptr())
  t.cuda.synchronize()
  assert int(status) == 0
  ref = wt.grad.reshape(cin, cout, 343).permute(2, 0, 1)
  assert rel_err(dw[:, :, :cout], ref) < 5e-5


@pytest.mark.parametrize("n,cin,cout,dhw", [(1, 16, 15, (4, 5, 64)), (2, 12, 7, (3, 6, 32))])
def test_convt7_wgrad_line_channel_slices(n, cin, cout, dhw):
  """The C > 4 logits layer's weight gradient as 4-output-channel slices of the narrow tap-stacked kernel: dy rows of
  pitch 16 (pad channels zero), dW columns land in a packed row of CoutP = r4(Cout)."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  g = t.Generator().manual_seed(cin + cout)
  x = t.randn((n, cin) + dhw, generator=g)
  wt = (t.randn(cin, cout, 7, 7, 7, generator=g) * 0.05).double().requires_grad_(True)
  y = F.conv_transpose3d(x.double(), wt, None, stride=2, padding=3, output_padding=1)
  gy = t.randn(y.shape, generator=g)
  y.backward(gy.double())
  fine = tuple(y.shape[2:])
  xr = x.permute(0, 2, 3, 4, 1).reshape(-1, cin).contiguous().to(dev())
  cp = (cout + 3) // 4 * 4
  gr = t.zeros(gy.numel() // cout, cp)
  gr[:, :cout] = gy.permute(0, 2, 3, 4, 1).reshape(-1, cout)
  gr = gr.to(dev())
  dw = t.zeros(343, cin, cp, device=dev())
  status = t.zeros(1, dtype=t.int32, device=dev())
  for co0 in range(0, cp, 4):
    desc = ops.make_desc(n, cin, 4, dhw, fine, (7, 7, 7), 2, 3, True, cin, cp)
    desc.y_co, desc.CoutP = co0, cp
    assert _lib.lib().crn_convt7_wgrad_line_supported(C.byref(desc)) == 2
    _lib.call("crn_convt7_wgrad_line", C.byref(desc), xr.data_ptr(), gr.data_ptr(), dw.data_ptr() + 4 * co0,
              status.data_ptr(), _lib.stream_ptr())
  t.cuda.synchronize()
  assert int(status) == 0
  ref = wt.grad.reshape(cin, cout, 343).permute(2, 0, 1)
  assert rel_err(dw[:, :, :cout], ref) < 5e-5
  assert float(dw[:, :, cout:].abs().max()) == 0.0 if cp > cout else True


@pytest.mark.parametrize("n,cin,cout,dhw", [(1, 8, 16, (8, 16, 8)), (2, 28, 16, (12, 32, 16)), (1, 12, 8, (8, 16, 24)),
                                              (1, 28, 16, (20, 16, 8))])
def test_conv5_kz_stacked_tcgen05(n, cin, cout, dhw):
  """Conv3d k=5 forward with the kz taps stacked into N (csrc/conv_tc5s.cu) against the fp64 oracle; same tolerance
  as the plain tcgen05 kernel (2e-4 of the tensor max)."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  d, h, w = dhw
  g = t.Generator().manual_seed(cin * 5 + cout + d)
  wt = t.randn(cout, cin, 5, 5, 5, generator=g) * 0.05
  bias = t.randn(cout, generator=g)
  x = t.randn(n, cin, d, h, w, generator=g)
  ref = F.conv3d(x.double(), wt.double(), bias.double(), padding=2)
  xin = x.permute(0, 2, 3, 4, 1).reshape(-1, cin).contiguous().to(dev())
  out = t.full((n * d * h * w, cout), float("nan"), device=dev())
  wtc = t.zeros(_lib.lib().crn_tc5s_packed_floats(cin), device=dev())
  st = _lib.stream_ptr()
  _lib.call("crn_tc5s_pack", wt.to(dev()).contiguous().data_ptr(), cout, cin, wtc.data_ptr(), st)
  desc = ops.make_desc(n, cin, cout, dhw, dhw, (5, 5, 5), 1, 2, False, cin, cout)
  status = t.zeros(1, dtype=t.int32, device=dev())
  b = bias.to(dev())
  _lib.call("crn_conv5_tcs", C.byref(desc), xin.data_ptr(), wtc.data_ptr(), b.data_ptr(), out.data_ptr(), status.data_ptr(), st)
  t.cuda.synchronize()
  assert int(status) == 0
  got = out.reshape(n, d, h, w, cout).permute(0, 4, 1, 2, 3)
  assert rel_err(got, ref) < 2e-4


@pytest.mark.parametrize("n,cin,cout,dhw,kind", [(1, 56, 32, (8, 16, 16), 0), (2, 28, 16, (12, 32, 16), 1),
                                                   (1, 12, 8, (8, 16, 8), 1), (1, 20, 24, (8, 16, 8), 0),
                                                   (1, 32, 56, (8, 16, 8), 1)])
def test_conv5_kz_stacked_general_tcgen05(n, cin, cout, dhw, kind):
  """kz-stacked Conv3d k=5 on tcgen05, general form: up to 32 output channels (three stacked MMAs per product) and
  dgrad (flipped / transposed weights), against the fp64 oracle."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  d, h, w = dhw
  g = t.Generator().manual_seed(cin * 3 + cout + kind)
  wt = t.randn(cout, cin, 5, 5, 5, generator=g) * 0.05
  bias = t.randn(cout, generator=g)
  if kind == 0:
    src = t.randn(n, cin, d, h, w, generator=g)
    ref = F.conv3d(src.double(), wt.double(), bias.double(), padding=2)
    K, N = cin, cout
  else:
    src = t.randn(n, cout, d, h, w, generator=g)
    ref = F.conv_transpose3d(src.double(), wt.double(), None, padding=2)
    K, N = cout, cin
  xin = src.permute(0, 2, 3, 4, 1).reshape(-1, K).contiguous().to(dev())
  out = t.full((n * d * h * w, N), float("nan"), device=dev())
  wtc = t.zeros(_lib.lib().crn_tc5s_packed_floats(K), device=dev())
  st = _lib.stream_ptr()
  _lib.call("crn_tc5s_pack2", wt.to(dev()).contiguous().data_ptr(), cout, cin, kind, wtc.data_ptr(), st)
  desc = ops.make_desc(n, cin, cout, dhw, dhw, (5, 5, 5), 1, 2, False, cin, cout)
  status = t.zeros(1, dtype=t.int32, device=dev())
  b = bias.to(dev())
  _lib.call("crn_conv5_tcs2", C.byref(desc), kind, xin.data_ptr(), wtc.data_ptr(), b.data_ptr(), out.data_ptr(),
            status.data_ptr(), st)
  t.cuda.synchronize()
  assert int(status) == 0
  got = out.reshape(n, d, h, w, N).permute(0, 4, 1, 2, 3)
  assert rel_err(got, ref) < 2e-4


@pytest.mark.parametrize("n,cin,cout,dhw,planar", [(1, 8, 2, (8, 16, 8), True), (2, 16, 2, (8, 32, 16), True),
                                                     (1, 12, 3, (8, 16, 8), True), (1, 8, 2, (8, 16, 8), False),
                                                     (1, 32, 16, (8, 16, 16), False), (1, 16, 8, (8, 16, 8), False),
                                                     (1, 20, 4, (16, 16, 8), False),
                                                     (2, 8, 16, (16, 32, 32), False)])   # 128 items: two planes per item
def test_conv_transpose7_tcgen05(n, cin, cout, dhw, planar):
  """ConvTranspose3d k=7 s=2 p=3 op=1 forward on the tcgen05 kernel (8 parity classes as one 4^3-tap conv with a
  scatter epilogue) against torch fp64; channels-last output inside a wider (concat) row, or planar logits.
  Tolerance 2e-5 of the tensor max (3xTF32, accumulation chains of <= 96 MMAs, measured ~2e-6)."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  d, h, w = dhw
  g = t.Generator().manual_seed(cin * 7 + cout)
  wt = t.randn(cin, cout, 7, 7, 7, generator=g) * 0.05
  bias = t.randn(cout, generator=g)
  x = t.randn(n, cin, d, h, w, generator=g)
  ref = F.conv_transpose3d(x.double(), wt.double(), bias.double(), stride=2, padding=3, output_padding=1)
  r4 = lambda c: (c + 3) // 4 * 4
  xin = t.zeros(n * d * h * w, r4(cin), device=dev())
  xin[:, :cin] = x.permute(0, 2, 3, 4, 1).reshape(-1, cin).to(dev())
  S = 8 * d * h * w
  ycs = r4(cout) + 4
  out = t.full((n * cout * S,) if planar else (n * S, ycs), float("nan"), device=dev())
  wtc = t.zeros(_lib.lib().crn_tct_packed_floats(cin, cout, 0), device=dev())
  st = _lib.stream_ptr()
  _lib.call("crn_tct_pack", wt.to(dev()).contiguous().data_ptr(), cin, cout, 0, wtc.data_ptr(), st)
  desc = ops.make_desc(n, cin, cout, (d, h, w), (2 * d, 2 * h, 2 * w), (7, 7, 7), 2, 3, True, r4(cin), ycs)
  desc.y_planar = int(planar)
  status = t.zeros(1, dtype=t.int32, device=dev())
  b = bias.to(dev())
  _lib.call("crn_convt7_tc", C.byref(desc), xin.data_ptr(), wtc.data_ptr(), b.data_ptr(), out.data_ptr(),
            status.data_ptr(), st)
  t.cuda.synchronize()
  assert int(status) == 0
  if planar:
    got = out.reshape(n, cout, 2 * d, 2 * h, 2 * w)
  else:
    got = out[:, :cout].reshape(n, 2 * d, 2 * h, 2 * w, cout).permute(0, 4, 1, 2, 3)
    assert bool(t.isnan(out[:, cout:]).all()), "columns outside the layer's slice must stay untouched"
  assert rel_err(got, ref) < 2e-5


@pytest.mark.parametrize("n,cin,cout,dhw", [(1, 8, 4, (8, 16, 8)), (1, 32, 16, (8, 16, 16)), (2, 20, 8, (8, 16, 8)),
                                              (1, 64, 32, (8, 16, 8))])
def test_conv_transpose7_dgrad_tcgen05(n, cin, cout, dhw):
  """dgrad of ConvTranspose3d k=7 s=2 on the tcgen05 kernel (class-channel gather from dy) against torch fp64:
  the adjoint of a transposed conv is the strided conv with the same weight tensor."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  d, h, w = dhw
  g = t.Generator().manual_seed(cin * 7 + cout + 1)
  wt = t.randn(cin, cout, 7, 7, 7, generator=g) * 0.05
  dy = t.randn(n, cout, 2 * d, 2 * h, 2 * w, generator=g)
  ref = F.conv3d(dy.double(), wt.double(), None, stride=2, padding=3)
  r4 = lambda c: (c + 3) // 4 * 4
  ycs = r4(cout) + 4
  dyin = t.zeros(n * 8 * d * h * w, ycs, device=dev())
  dyin[:, :cout] = dy.permute(0, 2, 3, 4, 1).reshape(-1, cout).to(dev())
  out = t.full((n * d * h * w, r4(cin)), float("nan"), device=dev())
  wtc = t.zeros(_lib.lib().crn_tct_packed_floats(cin, cout, 1), device=dev())
  st = _lib.stream_ptr()
  _lib.call("crn_tct_pack", wt.to(dev()).contiguous().data_ptr(), cin, cout, 1, wtc.data_ptr(), st)
  desc = ops.make_desc(n, cin, cout, (d, h, w), (2 * d, 2 * h, 2 * w), (7, 7, 7), 2, 3, True, r4(cin), ycs)
  status = t.zeros(1, dtype=t.int32, device=dev())
  _lib.call("crn_convt7_tc_dgrad", C.byref(desc), dyin.data_ptr(), wtc.data_ptr(), out.data_ptr(), status.data_ptr(), st)
  t.cuda.synchronize()
  assert int(status) == 0
  got = out[:, :cin].reshape(n, d, h, w, cin).permute(0, 4, 1, 2, 3)
  assert rel_err(got, ref) < 2e-5


@pytest.mark.parametrize("n,cin,cout,dhw,planar", [(1, 16, 2, (8, 16, 8), True), (2, 16, 2, (4, 32, 16), True),
                                                     (1, 12, 2, (8, 16, 8), False), (1, 20, 1, (4, 16, 8), True)])
def test_conv_transpose7_fwd_stacked_tcgen05(n, cin, cout, dhw, planar):
  """crn_convt7_tcs_fwd (the FG_BG logits layer: Cout <= 2, jz taps stacked into N) against torch fp64."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  d, h, w = dhw
  g = t.Generator().manual_seed(cin * 5 + cout)
  wt = t.randn(cin, cout, 7, 7, 7, generator=g) * 0.05
  bias = t.randn(cout, generator=g)
  x = t.randn(n, cin, d, h, w, generator=g)
  ref = F.conv_transpose3d(x.double(), wt.double(), bias.double(), stride=2, padding=3, output_padding=1)
  r4 = lambda c: (c + 3) // 4 * 4
  xin = t.zeros(n * d * h * w, r4(cin), device=dev())
  xin[:, :cin] = x.permute(0, 2, 3, 4, 1).reshape(-1, cin).to(dev())
  S = 8 * d * h * w
  ycs = r4(cout) + 4
  out = t.full((n * cout * S,) if planar else (n * S, ycs), float("nan"), device=dev())
  wtc = t.zeros(_lib.lib().crn_tctsf_packed_floats(cin), device=dev())
  st = _lib.stream_ptr()
  wd, b = wt.to(dev()).contiguous(), bias.to(dev())
  _lib.call("crn_tctsf_pack", wd.data_ptr(), cin, cout, wtc.data_ptr(), st)
  desc = ops.make_desc(n, cin, cout, (d, h, w), (2 * d, 2 * h, 2 * w), (7, 7, 7), 2, 3, True, r4(cin), ycs)
  desc.y_planar = int(planar)
  status = t.zeros(1, dtype=t.int32, device=dev())
  _lib.call("crn_convt7_tcs_fwd", C.byref(desc), xin.data_ptr(), wtc.data_ptr(), b.data_ptr(), out.data_ptr(),
            status.data_ptr(), st)
  t.cuda.synchronize()
  assert int(status) == 0
  if planar:
    got = out.reshape(n, cout, 2 * d, 2 * h, 2 * w)
  else:
    got = out[:, :cout].reshape(n, 2 * d, 2 * h, 2 * w, cout).permute(0, 4, 1, 2, 3)
    assert bool(t.isnan(out[:, cout:]).all()), "columns outside the layer's slice must stay untouched"
  assert rel_err(got, ref) < 2e-5


@pytest.mark.parametrize("n,cin,cout,dhw", [(1, 8, 4, (8, 16, 8)), (1, 32, 16, (8, 16, 16)), (2, 20, 8, (4, 16, 8)),
                                              (1, 16, 16, (12, 32, 16)), (1, 12, 4, (4, 16, 8))])
def test_conv_transpose7_dgrad_stacked_tcgen05(n, cin, cout, dhw):
  """crn_convt7_tcs_dgrad (jz taps stacked into N, Cin <= 32) against torch fp64 and against the un-stacked kernel."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  d, h, w = dhw
  g = t.Generator().manual_seed(cin * 11 + cout + 3)
  wt = t.randn(cin, cout, 7, 7, 7, generator=g) * 0.05
  dy = t.randn(n, cout, 2 * d, 2 * h, 2 * w, generator=g)
  ref = F.conv3d(dy.double(), wt.double(), None, stride=2, padding=3)
  r4 = lambda c: (c + 3) // 4 * 4
  ycs = r4(cout) + 4
  dyin = t.zeros(n * 8 * d * h * w, ycs, device=dev())
  dyin[:, :cout] = dy.permute(0, 2, 3, 4, 1).reshape(-1, cout).to(dev())
  out = t.full((n * d * h * w, r4(cin) + 4), float("nan"), device=dev())
  wtc = t.zeros(_lib.lib().crn_tcts_packed_floats(cout), device=dev())
  st = _lib.stream_ptr()
  wd = wt.to(dev()).contiguous()
  _lib.call("crn_tcts_pack", wd.data_ptr(), cin, cout, wtc.data_ptr(), st)
  desc = ops.make_desc(n, cin, cout, (d, h, w), (2 * d, 2 * h, 2 * w), (7, 7, 7), 2, 3, True, r4(cin) + 4, ycs)
  status = t.zeros(1, dtype=t.int32, device=dev())
  _lib.call("crn_convt7_tcs_dgrad", C.byref(desc), dyin.data_ptr(), wtc.data_ptr(), out.data_ptr(), status.data_ptr(), st)
  t.cuda.synchronize()
  assert int(status) == 0
  assert bool(t.isnan(out[:, r4(cin):]).all()), "columns outside the layer's slice must stay untouched"
  got = out[:, :cin].reshape(n, d, h, w, cin).permute(0, 4, 1, 2, 3)
  assert rel_err(got, ref) < 2e-5


@pytest.mark.parametrize("n,cin,dhw", [(1, 16, (8, 16, 8)), (2, 12, (4, 16, 16)), (1, 32, (4, 32, 8))])
def test_conv_transpose7_dgrad_stacked_planar_logits(n, cin, dhw):
  """crn_convt7_tcs_dgrad on the PLANAR gradient of the two FG_BG logits (float2 class gathers) against torch fp64."""
  import ctypes as C
  from corenet_b200 import _lib, ops
  d, h, w = dhw
  g = t.Generator().manual_seed(cin * 7 + d)
  wt = t.randn(cin, 2, 7, 7, 7, generator=g) * 0.05
  dy = t.randn(n, 2, 2 * d, 2 * h, 2 * w, generator=g)
  ref = F.conv3d(dy.double(), wt.double(), None, stride=2, padding=3)
  dyd = dy.to(dev()).contiguous()
  out = t.full((n * d * h * w, cin + 4), float("nan"), device=dev())
  wtc = t.zeros(_lib.lib().crn_tcts_packed_floats(2), device=dev())
  st = _lib.stream_ptr()
  wd = wt.to(dev()).contiguous()
  _lib.call("crn_tcts_pack", wd.data_ptr(), cin, 2, wtc.data_ptr(), st)
  desc = ops.make_desc(n, cin, 2, (d, h, w), (2 * d, 2 * h, 2 * w), (7, 7, 7), 2, 3, True, cin + 4, 4, planar=True)
  status = t.zeros(1, dtype=t.int32, device=dev())
  _lib.call("crn_convt7_tcs_dgrad", C.byref(desc), dyd.data_ptr(), wtc.data_ptr(), out.data_ptr(), status.data_ptr(), st)
  t.cuda.synchronize()
  assert int(status) == 0
  assert bool(t.isnan(out[:, cin:]).all()), "columns outside the layer's slice must stay untouched"
  got = out[:, :cin].reshape(n, d, h, w, cin).permute(0, 4, 1, 2, 3)
  assert rel_err(got, ref) < 2e-5


def test_linear():
  from corenet_b200 import ops
  g = t.Generator().manual_seed(5)
  x, w, b = t.randn(4, 2048, generator=g), t.randn(64, 2048, generator=g) * 0.02, t.randn(64, generator=g)
  y = ops.conv(x.to(dev()), w.to(dev()), b.to(dev()))
  assert rel_err(y, F.linear(x, w, b)) < 2e-5


# ---------------------------------------------------------------------------- BatchRenorm
@pytest.mark.parametrize("shape", [(4, 64, 9, 9), (2, 28, 6, 6, 6), (4, 67, 1, 1, 1), (3, 2048, 2, 2)],
                         ids=["2d", "3d_c28", "latent_c67", "wide"])
@pytest.mark.parametrize("training,nbt", [(True, 0), (True, 50000), (False, 123)])
def test_batch_renorm(shape, training, nbt):
  from corenet_b200.model.batch_renorm import BatchRenorm
  c = shape[1]
  g = t.Generator().manual_seed(c + nbt)
  x = t.randn(shape, generator=g) * 2 + 0.7
  st = {"weight": t.rand(c, generator=g) + 0.5, "bias": t.randn(c, generator=g),
        "running_mean": t.randn(c, generator=g) * 0.3, "running_var": t.rand(c, generator=g) + 0.4,
        "num_batches_tracked": t.tensor(nbt, dtype=t.int64)}
  gy = t.randn(shape, generator=g)
  so = {k: v.clone().requires_grad_(k in ("weight", "bias")) for k, v in st.items()}
  xo = x.clone().requires_grad_(True)
  nb = {}
  yo = O.batch_renorm(xo, so, "", training, nb)
  yo.backward(gy)
  m = BatchRenorm(c, eps=0.001).to(dev())
  m.load_state_dict(st)
  m.train(training)
  xc = x.clone().to(dev()).requires_grad_(True)
  yc = m(xc)
  yc.backward(gy.to(dev()))
  tol = 1e-4 if shape[0] * int(np.prod(shape[2:])) > 8 else 2e-3   # tiny batches: 1/sigma amplifies rounding
  assert rel_err(yc, yo) < tol
  assert rel_err(xc.grad, xo.grad) < 5 * tol
  assert rel_err(m.weight.grad, so["weight"].grad) < 5 * tol
  assert rel_err(m.bias.grad, so["bias"].grad) < 1e-4
  if training:
    assert int(m.num_batches_tracked) == nbt + 1
    assert rel_err(m.running_mean, nb["running_mean"]) < 1e-5
    assert rel_err(m.running_var, nb["running_var"]) < 1e-5
  else:
    assert int(m.num_batches_tracked) == nbt


# ---------------------------------------------------------------------------- small encoder ops
@pytest.mark.parametrize("rows,c,relu_in,relu_out,res,training,nbt", [
    (4 * 64 * 64, 64, 0, 1, 0, 1, 0), (4 * 16 * 16, 1024, 0, 1, 1, 1, 50000), (256, 2048, 0, 0, 0, 1, 12000),
    (2048, 224, 1, 0, 0, 1, 0), (1000, 8, 1, 1, 1, 1, 7000), (16384, 256, 0, 1, 1, 1, 100), (4 * 32 * 32, 512, 0, 1, 1, 0, 0), (77, 16, 0, 0, 0, 1, 3)])
def test_batch_renorm_fused_matches_three_kernel_path(rows, c, relu_in, relu_out, res, training, nbt):
  """csrc/brn_fused.cu (one launch per direction, cluster-owned channels, no atomics) against the stats / finalize /
  apply and reduce / dx kernels of csrc/brn.cu on the same inputs, plus run-to-run bit reproducibility."""
  from corenet_b200 import _lib
  lib = _lib.lib()
  assert lib.crn_brn_fused_supported(rows, c)
  d = dev()
  g = t.Generator().manual_seed(rows + c)
  cs = c + 8                                     # rows wider than the channel slice
  x = (t.randn(rows, cs, generator=g) * 2 + 0.5).to(d)
  resid = t.randn(rows, cs, generator=g).to(d) if res else None
  w, b = (t.rand(c, generator=g) + 0.5).to(d), t.randn(c, generator=g).to(d)
  rm0, rv0 = t.randn(c, generator=g).to(d) * 0.1, (t.rand(c, generator=g) + 0.5).to(d)
  dy = t.randn(rows, cs, generator=g).to(d)
  gex = t.randn(rows, cs, generator=g).to(d) if res else None
  st = _lib.stream_ptr()
  eps, mom = 1e-3, 0.01

  def ptr(v):
    return v.data_ptr() if v is not None else None

  def run(fused):
    rm, rv = rm0.clone(), rv0.clone()
    cnt = t.tensor(nbt, dtype=t.int64, device=d)
    coef = t.zeros(6 * c, device=d)
    y = t.full((rows, cs), float("nan"), device=d)
    ypre = t.full((rows, cs), float("nan"), device=d) if res else None
    acc = t.zeros(3 * c, dtype=t.float64, device=d)
    if fused:
      snap = t.zeros(1, dtype=t.int64, device=d)
      if training:
        plist = t.tensor([cnt.data_ptr()], dtype=t.int64, device=d)
        _lib.call("crn_brn_nbt_snapshot", plist.data_ptr(), 1, snap.data_ptr(), st)
      _lib.call("crn_brn_fwd_fused", x.data_ptr(), rows, c, cs, 0, relu_in, w.data_ptr(), b.data_ptr(), rm.data_ptr(),
                rv.data_ptr(), snap.data_ptr() if training else None, eps, mom, training, ptr(resid), relu_out,
                y.data_ptr(), cs, 0, ptr(ypre), coef.data_ptr(), st)
    else:
      if training:
        _lib.call("crn_brn_stats", x.data_ptr(), rows, c, cs, 0, relu_in, acc.data_ptr(), st)
      _lib.call("crn_brn_finalize", acc.data_ptr(), rows, c, w.data_ptr(), b.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                cnt.data_ptr(), eps, mom, training, coef.data_ptr(), st)
      _lib.call("crn_brn_apply", x.data_ptr(), rows, c, cs, 0, coef.data_ptr(), ptr(resid), relu_in, relu_out,
                y.data_ptr(), cs, 0, ptr(ypre), st)
    # backward
    gout = t.full((rows, cs), float("nan"), device=d) if (relu_out or res) else None
    dx = t.full((rows, cs), float("nan"), device=d)
    dw, db = t.zeros(c, device=d), t.zeros(c, device=d)
    dxsum = t.zeros(c, dtype=t.float64, device=d)
    bacc = t.zeros(2 * c, dtype=t.float64, device=d)
    if fused:
      _lib.call("crn_brn_bwd_fused", dy.data_ptr(), cs, 0, y.data_ptr() if relu_out else None, ptr(gex), x.data_ptr(),
                cs, 0, rows, c, coef.data_ptr(), relu_in, relu_out, training, ptr(gout), dx.data_ptr(), cs, 0, 0,
                dw.data_ptr(), db.data_ptr(), dxsum.data_ptr(), st)
    else:
      _lib.call("crn_brn_bwd_reduce", dy.data_ptr(), cs, 0, y.data_ptr() if relu_out else None, ptr(gex), x.data_ptr(),
                cs, 0, rows, c, coef.data_ptr(), relu_in, relu_out, ptr(gout), bacc.data_ptr(), st)
      gsrc = gout if gout is not None else dy
      _lib.call("crn_brn_bwd_dx", gsrc.data_ptr(), cs, 0, x.data_ptr(), cs, 0, rows, c, coef.data_ptr(), bacc.data_ptr(),
                None, relu_in, training, dx.data_ptr(), cs, 0, 0, dw.data_ptr(), db.data_ptr(), dxsum.data_ptr(), st)
    t.cuda.synchronize()
    return dict(y=y[:, :c], ypre=None if ypre is None else ypre[:, :c], coef=coef, rm=rm, rv=rv, cnt=int(cnt),
                gout=None if gout is None else gout[:, :c], dx=dx[:, :c], dw=dw, db=db, dxsum=dxsum,
                pad=(y[:, c:], dx[:, c:]))

  a, b1, b2 = run(False), run(True), run(True)
  assert a["cnt"] == b1["cnt"] == nbt + (1 if training else 0)
  for k in ("y", "ypre", "coef", "rm", "rv", "gout", "dx", "dw", "db", "dxsum"):
    if a[k] is None:
      continue
    tol = 2e-5 if k in ("dx", "dw", "db") else 5e-6
    if k == "dxsum":      # column sums of dx cancel to ~0 in training mode: compare against the scale of the summands
      scale = a["dx"].double().abs().sum(0)
      assert bool(((b1[k] - a[k]).abs() <= 1e-5 * scale + 1e-12).all()), k
    else:
      assert rel_err(b1[k], a[k]) <= tol, (k, rel_err(b1[k], a[k]))
    assert t.equal(b1[k], b2[k]), f"{k}: the fused path must be bit-reproducible"
  assert bool(t.isnan(b1["pad"][0]).all()) and bool(t.isnan(b1["pad"][1]).all()), "pad columns must stay untouched"


def test_preprocess_and_maxpool_and_mean():
  from corenet_b200 import _lib
  g = t.Generator().manual_seed(0)
  img = t.randint(0, 256, (2, 3, 256, 256), dtype=t.uint8, generator=g)
  out = t.empty(2, 256, 256, 4, device=dev())
  _lib.call("crn_preprocess_image", img.to(dev()).data_ptr(), 2, 256, 256, out.data_ptr(), _lib.stream_ptr())
  ref = O.preprocess_image_caffe(img).permute(0, 2, 3, 1)
  assert t.equal(out[..., :3].cpu(), ref) and float(out[..., 3].abs().max()) == 0.0   # exact
  # ZeroPad2d(1) + MaxPool2d(3, 2) on post-ReLU data, forward values exact, gradient exact
  x = t.randn(2, 8, 12, 12, generator=g).relu()
  xo = x.clone().requires_grad_(True)
  yo = F.max_pool2d(F.pad(xo, [1, 1, 1, 1]), 3, 2)
  gy = t.randn(yo.shape, generator=g)
  yo.backward(gy)
  xr = x.permute(0, 2, 3, 1).contiguous().to(dev())
  y = t.empty(2, 6, 6, 8, device=dev())
  idx = t.empty(2 * 6 * 6 * 8, dtype=t.int8, device=dev())
  _lib.call("crn_maxpool_fwd", xr.data_ptr(), 2, 12, 12, 8, y.data_ptr(), idx.data_ptr(), _lib.stream_ptr())
  assert t.equal(y.cpu().permute(0, 3, 1, 2), yo.detach())
  dx = t.empty_like(xr)
  gyr = gy.permute(0, 2, 3, 1).contiguous().to(dev())
  _lib.call("crn_maxpool_bwd", gyr.data_ptr(), idx.data_ptr(), 2, 12, 12, 8, dx.data_ptr(), _lib.stream_ptr())
  # ties can only happen at 0 (post-ReLU) where the ReLU mask kills the gradient anyway
  mask = (x > 0).permute(0, 2, 3, 1)
  assert t.allclose((dx.cpu() * mask), (xo.grad.permute(0, 2, 3, 1) * mask), rtol=0, atol=1e-6)
  # spatial mean
  f = t.randn(3, 64, 2048, generator=g)
  m = t.empty(3, 2048, device=dev())
  _lib.call("crn_spatial_mean_fwd", f.to(dev()).data_ptr(), 3, 64, 2048, m.data_ptr(), _lib.stream_ptr())
  assert rel_err(m, f.mean(1)) < 1e-6


# ---------------------------------------------------------------------------- ray-traced skip
def _skip_inputs(b, seed, dense=False):
  g = t.Generator().manual_seed(seed)
  v2s = O.default_v2s(b).clone()
  offs = t.rand(b, 3, generator=g)
  offs[0] = 0.5
  if b > 1:   # one camera translated so a chunk of the grid falls outside the image / behind it
    v2s[1] = O.dataset_camera() @ O.translate([0.35, -0.2, -1.1]) @ O.scale([128.0] * 3).inverse()
  if dense and b > 2:
    v2s[2] = v2s[2] + 0.01 * t.randn(4, 4, generator=g)
  return v2s, offs


@pytest.mark.parametrize("g3,hw", [(8, 8), (16, 16), (32, 32), (64, 64), (16, 8)])
def test_skip_indices_bit_exact(g3, hw):
  """Index parity with the reference's projection arithmetic: bit-exact."""
  from corenet_b200 import ops
  b = 3
  v2s, offs = _skip_inputs(b, g3, dense=True)
  mat = v2s.matmul(O.scale([128.0 / g3] * 3))
  ix, iy, front = O.sample_grid2d_indices(b, (g3, g3, g3), (hw, hw), mat, offs)
  exp = t.where(front, iy * (hw + 2) + ix, t.full_like(ix, -1)).to(t.int32)
  got = ops.skip_indices(b, (hw, hw), (g3, g3, g3), mat.to(dev()), offs.to(dev())).cpu()
  assert (front.float().mean() < 1.0) and (front.float().mean() > 0.3)   # the edge cases are exercised
  assert t.equal(got, exp)


@pytest.mark.parametrize("g3,c_in,c_out", [(8, 40, 96), (16, 36, 48), (32, 20, 24), (64, 8, 12)])
def test_sample_grid2d_module(g3, c_in, c_out):
  from corenet_b200.model.ray_traced_skip_connection import SampleGrid2d
  b, hw = 2, g3
  g = t.Generator().manual_seed(g3)
  v2s, offs = _skip_inputs(b, g3)
  mat = v2s.matmul(O.scale([128.0 / g3] * 3))
  x = t.randn(b, c_in, hw, hw, generator=g)
  m = SampleGrid2d(c_in, c_out, (g3, g3, g3))
  w, bias = m.compress_channels.weight.detach().clone(), m.compress_channels.bias.detach().clone()
  xo, wo, bo = x.clone().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
  yo = O.sample_grid2d(xo, wo, bo, (g3, g3, g3), mat, offs)
  gy = t.randn(yo.shape, generator=g)
  yo.backward(gy)
  m = m.to(dev())
  xc = x.clone().to(dev()).requires_grad_(True)
  yc = m(xc, mat.to(dev()), offs.to(dev()))
  yc.backward(gy.to(dev()))
  assert rel_err(yc, yo) < 2e-5
  # the gather itself is pure data movement: exact zero pattern (outside / behind camera)
  assert t.equal(yc.cpu() == 0, yo == 0)
  assert rel_err(xc.grad, xo.grad) < 5e-5
  assert rel_err(m.compress_channels.weight.grad, wo.grad) < 5e-5
  assert rel_err(m.compress_channels.bias.grad, bo.grad) < 5e-5


def test_skip_gather_bit_exact():
  """With the compress conv factored out the kernel's output must equal the reference gather bit for bit."""
  from corenet_b200.ops import _SkipGatherFn
  b, g3, hw, c = 2, 16, 16, 48
  v2s, offs = _skip_inputs(b, 7)
  mat = v2s.matmul(O.scale([128.0 / g3] * 3))
  cmap = t.randn(b, c, hw, hw, generator=t.Generator().manual_seed(3))
  ix, iy, front = O.sample_grid2d_indices(b, (g3,) * 3, (hw, hw), mat, offs)
  padded = F.pad(cmap, [1, 1, 1, 1])
  bb = t.arange(b)[:, None, None, None].expand_as(ix)
  exp = padded[bb, :, iy, ix].permute(0, 4, 1, 2, 3) * front[:, None]
  got = _SkipGatherFn.apply(cmap.to(dev()), (g3,) * 3, mat.to(dev()), offs.to(dev())).cpu()
  assert t.equal(got, exp)


@pytest.mark.parametrize("g3,hw,c", [(64, 64, 12), (32, 32, 24), (16, 16, 48), (8, 8, 96), (16, 8, 20)])
def test_skip_backward_sorted_is_exact_and_reproducible(g3, hw, c):
  """The atomics-free backward (voxels sorted by sampled pixel, per-pixel gather in list order): equals a float64
  index_add of the same scatter to fp32 rounding, equals the atomic kernel to rounding, covers every voxel exactly
  once, and two runs are bit-identical."""
  from corenet_b200 import _lib, ops
  b = 3
  v2s, offs = _skip_inputs(b, g3 + c, dense=True)
  mat = v2s.matmul(O.scale([128.0 / g3] * 3)).contiguous()
  ix, iy, front = O.sample_grid2d_indices(b, (g3,) * 3, (hw, hw), mat, offs)
  valid = front & (ix >= 1) & (ix <= hw) & (iy >= 1) & (iy <= hw)
  pix = (t.arange(b)[:, None, None, None] * hw + (iy - 1)) * hw + (ix - 1)
  cs, co = c + 8, 4                         # gradient rows wider than the slice, like a concat buffer
  g = t.Generator().manual_seed(g3)
  dout = t.randn(b * g3 ** 3, cs, generator=g)
  exp = t.zeros(b * hw * hw, c, dtype=t.float64)
  exp.index_add_(0, pix[valid].reshape(-1), dout.reshape(b, g3, g3, g3, cs)[valid][:, co:co + c].double())
  d = dev()
  args = (dout.to(d), cs, co, b, hw, hw, c, mat.to(d), offs.to(d), (g3,) * 3)
  got = ops.skip_scatter_sorted(*args)
  got2 = ops.skip_scatter_sorted(*args)
  assert t.equal(got, got2)
  assert rel_err(got, exp) < 1e-5
  # the lists: a permutation of all voxels, segment p holds exactly the voxels that sample pixel p, ascending
  nb = _lib.lib().crn_skip_lists_workspace_bytes(b, hw, hw, g3, g3, g3)
  ws = t.empty(nb, dtype=t.uint8, device=d)
  sv = t.empty(b * g3 ** 3, dtype=t.int32, device=d)
  starts = t.empty(b * hw * hw + 1, dtype=t.int32, device=d)
  _lib.call("crn_skip_build_lists", b, hw, hw, args[7].data_ptr(), args[8].data_ptr(), g3, g3, g3, ws.data_ptr(), nb,
            sv.data_ptr(), starts.data_ptr(), _lib.stream_ptr())
  sv, starts = sv.cpu().long(), starts.cpu().long()
  assert t.equal(sv.sort().values, t.arange(b * g3 ** 3))
  key = t.where(valid, pix, t.full_like(pix, b * hw * hw)).reshape(-1)
  assert t.equal(key[sv], key[sv].sort().values)                       # sorted by pixel
  assert t.equal(starts, t.searchsorted(key[sv], t.arange(b * hw * hw + 1)))
  seg = key[sv]
  same = seg[1:] == seg[:-1]
  assert bool((sv[1:][same] > sv[:-1][same]).all())                    # stable: ascending voxel index per pixel
  # the atomic kernel computes the same sums (different order)
  dmap = t.zeros(b * hw * hw, c, dtype=t.float32, device=d)
  _lib.call("crn_skip_sample_bwd", args[0].data_ptr(), cs, co, b, hw, hw, c, c, args[7].data_ptr(),
            args[8].data_ptr(), g3, g3, g3, dmap.data_ptr(), _lib.stream_ptr())
  assert rel_err(dmap, exp) < 1e-5


# ---------------------------------------------------------------------------- losses
def test_losses_known_answers():
  """src/corenet/test/losses_test.py:72-88 through the CUDA kernels."""
  from corenet_b200.model import losses
  lg, gt = LOGITS.contiguous().to(dev()), GT.to(dev())
  np.testing.assert_allclose(losses.iou_fgbg(gt, lg).item(), 0.3579613, rtol=1e-5, atol=1e-6)
  want = (1 + 0.8060565) * (1 + 1.4547757)
  np.testing.assert_allclose(losses.xent_times_iou_agnostic(gt, lg).item(), want, rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(losses.iou_agnostic(gt, lg).item(), 0.8060565, rtol=1e-5, atol=1e-6)     # :73
  np.testing.assert_allclose(losses.xent(gt, lg).item(), 1.4547757, rtol=1e-5, atol=1e-6)             # :85
  np.testing.assert_allclose(losses.xent_times_iou_fgbg(gt, lg).item(), (1 + 0.3579613) * (1 + 1.4547757), rtol=1e-5)


@pytest.mark.parametrize("c,loss", [(2, "iou_fgbg"), (5, "iou_fgbg"), (5, "xent_times_iou_agnostic"),
                                    (15, "xent_times_iou_agnostic"), (5, "iou_agnostic"), (15, "xent"), (2, "xent")])
def test_losses_random(c, loss):
  from corenet_b200.model import losses
  g = t.Generator().manual_seed(c)
  lg = t.randn(3, c, 12, 10, 14, generator=g) * 3
  gt = t.randint(0, c, (3, 12, 10, 14), generator=g)
  gt[2] = 0     # a scene without foreground
  lo = lg.clone().requires_grad_(True)
  vo = getattr(O, loss)(gt, lo)
  (vo * 1.7).backward()
  lc = lg.clone().to(dev()).requires_grad_(True)
  vc = getattr(losses, loss)(gt.to(dev()), lc)
  (vc * 1.7).backward()
  np.testing.assert_allclose(vc.item(), vo.item(), rtol=2e-6, atol=1e-6)
  assert rel_err(lc.grad, lo.grad) < 1e-4
  vc32 = getattr(losses, loss)(gt.to(dev()).to(t.int32), lc.detach())
  assert vc32.item() == vc.item()


def test_softmax_and_confusion():
  from corenet_b200 import ops
  g = t.Generator().manual_seed(1)
  lg = t.randn(2, 6, 9, 9, 9, generator=g)
  gt = t.randint(0, 6, (2, 9, 9, 9), generator=g)
  assert rel_err(ops.softmax_channels(lg.to(dev())), lg.softmax(1)) < 1e-6
  cm = ops.argmax_confusion(lg.to(dev()), gt.to(dev())).cpu()
  assert t.equal(cm, O.confusion_matrix(lg.argmax(1), gt, 6))      # integer: exact


# ---------------------------------------------------------------------------- fill_inside_voxels
@pytest.mark.parametrize("dtype", [t.float32, t.uint8, t.int32, t.float64, t.int64, t.int8])
def test_fill_known_answers(dtype):
  """EmptyRegionFillTests (src/corenet/test/voxelization_test.py:197-244)."""
  from corenet_b200.cc import fill_voxels
  grids, expected = fill_test_grids()
  gi = t.from_numpy(grids).to(dtype)
  out = fill_voxels.fill_inside_voxels_gpu(gi.to(dev()), inplace=False)
  assert out.dtype == dtype
  np.testing.assert_array_equal(out.cpu().numpy(), expected.astype(out.cpu().numpy().dtype))
  np.testing.assert_array_equal(fill_voxels.fill_inside_voxels_cpu(gi).numpy(),
                                expected.astype(out.cpu().numpy().dtype))
  x = gi.to(dev())
  assert fill_voxels.fill_inside_voxels_gpu(x, inplace=True) is x
  np.testing.assert_array_equal(x.cpu().numpy(), expected.astype(out.cpu().numpy().dtype))


@pytest.mark.parametrize("shape,p", [((3, 16, 16, 16), 0.3), ((2, 9, 20, 45), 0.45), ((2, 7, 7, 7), 0.5),
                                     ((1, 40, 33, 70), 0.38), ((1, 1, 1, 1), 0.5), ((2, 1, 8, 3), 0.4),
                                     ((2, 64, 64, 64), 0.34), ((1, 33, 65, 129), 0.36), ((1, 5, 5, 256), 0.3),
                                     ((2, 17, 30, 128), 0.4), ((1, 130, 20, 40), 0.42), ((1, 6, 9, 300), 0.35),
                                     ((1, 3, 4, 1000), 0.3), ((20, 24, 24, 24), 0.4)])
def test_fill_random_bit_exact(shape, p):
  """Random grids (percolating densities: long winding paths) through the shared-memory cluster kernel, the
  register line-sweep kernel (W <= 256) and the wide kernel (any W; the reference has no size limit): bit-exact."""
  from corenet_b200 import _lib
  from corenet_b200.cc import fill_voxels
  rng = np.random.default_rng(sum(shape))
  g = (rng.random(shape) < p).astype(np.float32)
  exp = FO.fill_inside_voxels_oracle(g)
  out = fill_voxels.fill_inside_voxels_gpu(t.from_numpy(g).to(dev())).cpu().numpy()
  np.testing.assert_array_equal(out, exp)
  _lib.lib().crn_set_flags(4096)             # A/B: global-memory line sweeps
  try:
    out2 = fill_voxels.fill_inside_voxels_gpu(t.from_numpy(g).to(dev())).cpu().numpy()
  finally:
    _lib.lib().crn_set_flags(0)
  np.testing.assert_array_equal(out2, exp)


def test_fill_maze_and_shells_128():
  """Worst-case-ish connectivity at the full 128^3 size: a serpentine corridor (many sweeps) plus
  closed shells; checked against the oracle bit for bit."""
  from corenet_b200.cc import fill_voxels
  g = np.zeros((2, 128, 128, 128), np.float32)
  # scene 0: walls every 4 voxels along y with alternating gaps -> serpentine empty corridor
  for i, y in enumerate(range(3, 128, 4)):
    g[0, :, y, :] = 1
    if i % 2 == 0:
      g[0, :, y, 120:] = 0
    else:
      g[0, :, y, :8] = 0
  g[0, :, :, 127] = 0
  # scene 1: nested closed shells + one shell with a hole towards the far face
  for lo, hi in ((10, 60), (20, 50), (70, 120)):
    g[1, lo:hi, lo:hi, lo:hi] = 1
    g[1, lo + 1:hi - 1, lo + 1:hi - 1, lo + 1:hi - 1] = 0
  g[1, 95, 95, 119:] = 0
  out = fill_voxels.fill_inside_voxels_gpu(t.from_numpy(g).to(dev())).cpu().numpy()
  np.testing.assert_array_equal(out, FO.fill_inside_voxels_oracle(g))
  # size-independent properties: idempotent, monotone, occupied voxels stay occupied
  again = fill_voxels.fill_inside_voxels_gpu(t.from_numpy(out).to(dev())).cpu().numpy()
  np.testing.assert_array_equal(again, out)
  assert (out >= g).all()


def test_fill_errors():
  from corenet_b200.cc import fill_voxels
  with pytest.raises(ValueError):
    fill_voxels.fill_inside_voxels_gpu(t.zeros(4, 4, 4, device=dev()))
  with pytest.raises(ValueError):
    fill_voxels.fill_inside_voxels_gpu(t.zeros(1, 4, 4, 4))
  assert fill_voxels.fill_inside_voxels_gpu(t.zeros(0, 4, 4, 4, device=dev())).shape == (0, 4, 4, 4)


# ---------------------------------------------------------------------------- voxelize_mesh
def test_voxelize_reference_vectors():
  """VoxelizationTests (src/corenet/test/voxelization_test.py:53-147) end to end on the GPU."""
  from corenet_b200.cc import fill_voxels
  from corenet_b200.geometry import transformations, voxelization
  quad = t.tensor([[[0, 0, 0], [1, 0, 1], [0, 1, 0]], [[1, 0, 1], [0, 1, 0], [1, 1, 1]]], dtype=t.float32)
  vg = voxelization.voxelize_mesh(quad, [2], (4, 4, 4), transformations.scale([4, 4, 4]),
                                  image_resolution_multiplier=16)
  fill_voxels.fill_inside_voxels_gpu(vg, inplace=True)
  exp = np.zeros((4, 4, 4), np.float32)
  for z in range(4):
    exp[z, :, z] = 1
  np.testing.assert_array_equal(vg.cpu().numpy(), exp[None])
  cube = t.from_numpy(cube_mesh(0.99))
  eye = transformations.scale([1, 1, 1])
  grid = voxelization.voxelize_mesh(cube, [12], (3, 3, 3), eye, image_resolution_multiplier=1)
  e = np.zeros((3, 3, 3)); e[1, 1, [0, 2]] = e[1, [0, 2], 1] = e[[0, 2], 1, 1] = 1
  np.testing.assert_array_equal(grid.cpu().numpy(), e[None])
  grid = voxelization.voxelize_mesh(cube, [12], (3, 3, 3), eye, image_resolution_multiplier=1,
                                    conservative_rasterization=True)
  e = np.ones((3, 3, 3)); e[1, 1, 1] = 0
  np.testing.assert_array_equal(grid.cpu().numpy(), e[None])
  grid = voxelization.voxelize_mesh(cube, [12], (3, 3, 3), eye, sub_grid_sampling=True,
                                    image_resolution_multiplier=9, conservative_rasterization=True)
  grid = fill_voxels.fill_inside_voxels_gpu(grid, inplace=False)
  e = np.zeros((1, 7, 7, 7)); e[0, 2:5, 2:5, 2:5] = 1
  np.testing.assert_array_equal(grid.cpu().numpy(), e)
  e = np.zeros((1, 3, 3, 3)); e[0, 1, 1, 1] = 1
  np.testing.assert_array_equal(voxelization.get_sub_grid_centers(grid).cpu().numpy(), e)
  cubes = t.cat([cube, cube - 0.5])
  transf = t.stack([transformations.translate([-0.5, 0, 0]), transformations.translate([0.5, 1, 1])])
  grid = voxelization.voxelize_mesh(cubes, [12, 12], (3, 3, 3), transf, sub_grid_sampling=True,
                                    image_resolution_multiplier=9, conservative_rasterization=True)
  grid = voxelization.get_sub_grid_centers(fill_voxels.fill_inside_voxels_gpu(grid)).cpu().numpy()
  e1 = np.zeros((3, 3, 3)); e1[1, 1, [0, 1]] = 1
  e2 = np.zeros((3, 3, 3)); e2[1, [1, 2], 1] = e2[2, [1, 2], 1] = 1
  np.testing.assert_array_equal(grid[0], e1)
  np.testing.assert_array_equal(grid[1], e2)
  with pytest.raises(ValueError):
    voxelization.voxelize_mesh(cube, [12], (3, 3, 3), eye, sub_grid_sampling=True, image_resolution_multiplier=8)


@pytest.mark.parametrize("cons,sub,mult,pdm", [(False, False, 4, 1), (True, False, 3, 1), (False, False, 4, 2),
                                               (True, True, 5, 1), (False, True, 3, 1)])
def test_voxelize_random_meshes_bit_exact(cons, sub, mult, pdm):
  """Random triangle soups (incl. slivers and out-of-bounds parts) vs the numpy oracle: bit-exact."""
  from corenet_b200.geometry import voxelization
  rng = np.random.default_rng(mult * 10 + pdm + cons)
  res = (6, 7, 8)
  ntri = [25, 17]
  tris = rng.uniform(-0.2, 1.2, size=(sum(ntri), 3, 3)).astype(np.float32)
  tris[5] = tris[5, 0][None] + rng.normal(0, 0.01, (3, 3)).astype(np.float32)      # sliver
  v2x = np.stack([np.diag([8, 7, 6, 1]).astype(np.float32), np.diag([8, 7, 6, 1]).astype(np.float32)])
  v2x[1, :3, 3] = [0.3, -0.4, 0.25]
  exp = VO.voxelize_mesh_oracle(tris, ntri, res, v2x, sub_grid_sampling=sub, image_resolution_multiplier=mult,
                                conservative_rasterization=cons, projection_depth_multiplier=pdm)
  got = voxelization.voxelize_mesh(t.from_numpy(tris), ntri, res, t.from_numpy(v2x), sub_grid_sampling=sub,
                                   image_resolution_multiplier=mult, conservative_rasterization=cons,
                                   projection_depth_multiplier=pdm).cpu().numpy()
  assert exp.sum() > 20
  np.testing.assert_array_equal(got, exp)


@pytest.mark.parametrize("cons", [False, True])
def test_voxelize_large_triangles_and_translated_lattice_bit_exact(cons):
  """The inputs of the oracle's property tests (tests/test_oracle_voxelize.py): triangles spanning the whole 12^3 grid
  and beyond, a dyadic-lattice mesh moved by whole voxels through view2voxel, and a closed box -- vs the oracle."""
  from corenet_b200.geometry import voxelization
  from tests.conftest import cube_mesh
  rng = np.random.default_rng(7)
  c = rng.uniform(1.0, 11.0, size=(40, 1, 3))
  soup = (c + rng.normal(0, 2.0, size=(40, 3, 3))).astype(np.float32)
  lattice = (rng.integers(16, 48, size=(25, 3, 3)) / 8.0).astype(np.float32)
  box = (cube_mesh(0.0) / 3.0 * (9.6 - 2.3) + 2.3).astype(np.float32)
  tris = np.concatenate([soup, lattice, box])
  ntri = [40, 25, 12]
  v2x = np.stack([np.eye(4, dtype=np.float32)] * 3)
  v2x[1, :3, 3] = [3, 1, 2]
  res = (12, 12, 12)
  exp = VO.voxelize_mesh_oracle(tris, ntri, res, v2x, image_resolution_multiplier=5, conservative_rasterization=cons)
  got = voxelization.voxelize_mesh(t.from_numpy(tris), ntri, res, t.from_numpy(v2x), image_resolution_multiplier=5,
                                   conservative_rasterization=cons).cpu().numpy()
  assert exp[0].sum() > 100 and exp[1].sum() > 30 and exp[2].sum() > 200
  np.testing.assert_array_equal(got, exp)


def test_voxelize_fill_roundtrip_128():
  """Full size (128^3, mult 8): a closed icosphere-ish mesh voxelised + filled must be solid:
  every voxel whose centre is well inside the sphere is 1, everything well outside is 0."""
  from corenet_b200.cc import fill_voxels
  from corenet_b200.geometry import transformations, voxelization
  # UV sphere, radius 0.3 at (0.5, 0.5, 0.5) in the unit cube
  nu, nv = 48, 24
  th = np.linspace(0, 2 * np.pi, nu + 1)
  ph = np.linspace(0, np.pi, nv + 1)
  P = lambda i, j: np.array([0.5 + 0.3 * np.sin(ph[j]) * np.cos(th[i]), 0.5 + 0.3 * np.sin(ph[j]) * np.sin(th[i]),
                             0.5 + 0.3 * np.cos(ph[j])], np.float32)
  tris = []
  for i in range(nu):
    for j in range(nv):
      a, b, c, d = P(i, j), P(i + 1, j), P(i + 1, j + 1), P(i, j + 1)
      tris += [[a, b, c], [a, c, d]]
  tris = t.from_numpy(np.array(tris, np.float32))
  grid = voxelization.voxelize_mesh(tris, [tris.shape[0]], (128, 128, 128), transformations.scale([128.0] * 3),
                                    image_resolution_multiplier=8)
  shell = grid.clone()
  fill_voxels.fill_inside_voxels_gpu(grid, inplace=True)
  g = grid[0].cpu().numpy()
  zz, yy, xx = np.meshgrid(*[np.arange(128) + 0.5] * 3, indexing="ij")
  r = np.sqrt((zz - 64) ** 2 + (yy - 64) ** 2 + (xx - 64) ** 2) / 128
  assert g[r < 0.28].min() == 1 and g[r > 0.32].max() == 0
  assert shell[0].cpu().numpy()[r < 0.25].max() == 0        # surface only before the fill
