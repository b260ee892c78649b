"""Execution engine: CoReNet forward/backward as a static plan of C-ABI kernel launches.

This is the host side of the hot path.  It replaces the
autograd graph of ATen ops the reference builds in
src/corenet/model/core_net.py:36-43 -> resnet50.py:176-186 ->
reconstruction_decoder.py:119-152.  Design (DESIGN.md "Engine"):

  * activations are channels-last rows x C buffers allocated once per
    (batch, mode) plan; concat buffers are written in place by their two
    producers (transposed conv + ray-traced skip kernel);
  * weights are re-packed tap-major into one arena by ONE kernel per step,
    weight gradients land in a mirror arena and are un-packed by ONE kernel;
  * every launch goes through the C-ABI in include/corenet_b200.h on the
    current torch stream; torch is used for memory, streams and tiny 4x4 /
    [B,C]-sized glue only;
  * gradients reach autograd through ONE torch.autograd.Function whose inputs
    are the model's parameters, so DDP hooks / optimisers see ordinary grads.
"""
import ctypes as C
import os
from typing import Dict, List, Optional

import torch as t

from corenet_b200 import _lib
from corenet_b200._lib import ConvDesc, PackItem, UnpackItem

BRN_EPS = 1e-3        # every BatchRenorm of the model is built with eps=0.001
BRN_MOMENTUM = 0.01

# (stage, cin is derived, conv k, convT k, convT pad, conv out, convT out, encoder channels of the skip)
DECODER_PYRAMID = ((2, 3, 3, 1, 256, 128, 2048), (3, 5, 7, 3, 128, 64, 1024), (4, 5, 7, 3, 64, 32, 512),
                   (5, 5, 7, 3, 32, 16, 256), (6, 5, 7, 3, 16, None, None))
ENC_STAGE_OF = {2048: "stage5", 1024: "stage4", 512: "stage3", 256: "stage2"}


def _r4(c: int) -> int:
  return (c + 3) // 4 * 4


# Gradient chunks in the order the backward pass completes them (prefixes of parameter names).  In
# named_parameters() order they are three contiguous ranges of the flat buffer: [2 | 1 | 0].
GRAD_CHUNKS = (("decoder.",), ("encoder.stage4.", "encoder.stage5."), ("encoder.",))


def grad_chunk_of(name: str) -> int:
  for i, prefixes in enumerate(GRAD_CHUNKS):
    if name.startswith(prefixes):
      return i
  raise KeyError(name)


def _call(name, *args, hbm=None):
  """C-ABI launch on the current stream.  Under PROFILE the HBM-bound kernels of the path are bracketed with CUDA
  events: entries ("hbm", family, algorithmic bytes, start, end) -- bench.py turns them into achieved GB/s.
  hbm = (label, bytes) overrides the per-function table."""
  if NCU_PICK is not None and NCU_PICK("call", name):
    _ncu_bracket(name, args)
    return
  fam = (hbm and (lambda a: hbm)) or HBM_BYTES.get(name) if PROFILE is not None else None
  if fam is None:
    _lib.call(name, *args)
    return
  e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
  e0.record()
  _lib.call(name, *args)
  e1.record()
  label, nbytes = fam(args)
  PROFILE.append(("hbm", label, nbytes, e0, e1))


# scripts/ncu_capture.py sets this to a predicate (kind, name) -> bool: the chosen launches are bracketed with
# cudaProfilerStart/Stop so that `ncu --profile-from-start off` captures exactly them (kind = "call" with the C-ABI
# name for non-conv launches, else the conv tag fwd_tc / dgrad_gt / wgrad_tc ... with the layer name).
NCU_PICK = None


def _ncu_bracket(fn, args):
  t.cuda.cudart().cudaProfilerStart()
  try:
    _lib.call(fn, *args)
  finally:
    t.cuda.cudart().cudaProfilerStop()


def _b_skip_fwd(a):      # (cmap, B, h, w, Cs, cmap_cs, mat, offs, gd, gh, gw, out, out_cs, out_co, st)
  b, h, w, cs, gd, gh, gw = a[1], a[2], a[3], a[4], a[8], a[9], a[10]
  return f"skip_fwd g={gd}", 4 * cs * b * (gd * gh * gw + h * w)


def _b_skip_bwd(a):      # (dy, dy_cs, dy_co, B, h, w, Cs, cmap_cs, mat, offs, gd, gh, gw, dmap, st)
  b, h, w, cs, gd, gh, gw = a[3], a[4], a[5], a[6], a[10], a[11], a[12]
  return f"skip_bwd g={gd}", 4 * cs * b * (gd * gh * gw + h * w)


# C-ABI name -> args -> (family label, algorithmic bytes): compulsory reads + writes of the op
HBM_BYTES = {
    "crn_skip_sample_fwd": _b_skip_fwd,
    "crn_skip_sample_bwd": _b_skip_bwd,
    "crn_brn_stats": lambda a: ("brn_stats", 4 * a[1] * a[2]),
    "crn_brn_apply": lambda a: ("brn_apply", (8 + (4 if a[6] else 0) + (4 if a[12] else 0)) * a[1] * a[2]),
    "crn_brn_bwd_reduce": lambda a: ("brn_bwd_reduce", (8 + (4 if a[3] else 0) + (4 if a[4] else 0)
                                                          + (4 if a[13] else 0)) * a[8] * a[9]),
    "crn_brn_bwd_dx": lambda a: ("brn_bwd_dx", 12 * a[6] * a[7]),
    "crn_loss_sums": lambda a: ("loss_sums", (4 * a[4] + (8 if a[2] else 4)) * a[3] * a[5]),
    "crn_loss_bwd": lambda a: ("loss_bwd", (8 * a[4] + (8 if a[2] else 4)) * a[3] * a[5]),
    "crn_softmax_planar": lambda a: ("softmax", 8 * a[1] * a[2] * a[3]),
    "crn_loss_sums_l": lambda a: ("loss_sums", (4 * (a[1] or a[5]) + (8 if a[3] else 4)) * a[4] * a[6]),
    "crn_loss_bwd_l": lambda a: ("loss_bwd", (4 * (a[1] or a[5]) + 4 * (a[11] or a[5]) + (8 if a[3] else 4)) * a[4] * a[6]),
    "crn_softmax_l": lambda a: ("softmax", 4 * ((a[1] or a[3]) + a[3]) * a[2] * a[4]),
    "crn_argmax_confusion_l": lambda a: ("argmax_confusion", (4 * (a[1] or a[5]) + (8 if a[3] else 4)) * a[4] * a[6]),
    "crn_rows_to_planar": lambda a: ("rows_to_planar", 4 * (a[4] + a[2]) * a[1] * a[3]),
    "crn_argmax_confusion_labeled": lambda a: ("argmax_confusion", (4 * a[4] + (8 if a[2] else 4)) * a[3] * a[5]),
    "crn_adam_step_guarded": lambda a: ("adam", 28 * a[4]),
    "crn_unpack_wgrads": lambda a: ("unpack_wgrads", 8 * a[3]),
    "crn_planar_to_rows": lambda a: ("planar_to_rows", 8 * a[1] * a[2] * a[3]),
}


# bench.py sets this to a list to time every convolution launch with CUDA events:
# entries are (kind, layer name, MACs, start event, end event).
PROFILE = None
_CONV_FN = {"fwd": "crn_conv_fwd", "dgrad": "crn_conv_dgrad", "wgrad": "crn_conv_wgrad"}


def conv_macs(d: ConvDesc) -> int:
  """Algorithmic multiply-accumulates of one conv launch (same for fwd, dgrad and wgrad)."""
  taps = d.kD * d.kH * d.kW
  pos = d.iD * d.iH * d.iW if d.transposed else d.oD * d.oH * d.oW
  return d.N * pos * taps * d.Cin * d.Cout


_ENG = None      # engine whose plan is currently enqueuing (set by Plan.forward / Plan.backward)


def _conv_dispatch(kind, layer, d, args):
  """-> (tag, C-ABI name, argument tuple): wide layers go to the tcgen05 implicit-GEMM kernels
  (csrc/conv_gemm_tc.cu, csrc/conv_wgrad_tc.cu), everything else to the FFMA kernels."""
  eng = _ENG
  if USE_TC and eng is not None:
    status = eng.tc_status.data_ptr()
    if kind == "fwd" and layer.name in eng.gt_w:
      x, _, bias, y, acc, st = args
      return "fwd_gt", "crn_conv_gemm_tc", (C.byref(d), 0, x, eng.gt_w[layer.name][0].data_ptr(), bias, y, acc, status, st)
    if kind == "fwd" and layer.name in eng.gt_tf:
      x, _, bias, y, acc, st = args
      return "fwd_gt", "crn_conv_gemm_tc", (C.byref(d), 0, x, eng.gt_tf[layer.name].data_ptr(), bias, y, acc, status, st)
    if kind == "dgrad" and eng.gt_w.get(layer.name, (None, None))[1] is not None:
      dy, _, dx, acc, st = args
      return "dgrad_gt", "crn_conv_gemm_tc", (C.byref(d), 1, dy, eng.gt_w[layer.name][1].data_ptr(), None, dx, acc, status, st)
    if kind == "dgrad" and layer.name in eng.gt_td:
      # dgrad of a transposed convolution IS a strided forward convolution of dy with the same weight tensor
      # ([Cin][Cout][taps] read as conv weight [Cout'=Cin][Cin'=Cout][taps]): dx[i] = sum_k dy[s*i - pad + k] W[k]
      dy, _, dx, acc, st = args
      d2 = ConvDesc()
      d2.N, d2.Cin, d2.Cout = d.N, d.Cout, d.Cin
      d2.iD, d2.iH, d2.iW, d2.oD, d2.oH, d2.oW = d.oD, d.oH, d.oW, d.iD, d.iH, d.iW
      d2.kD, d2.kH, d2.kW, d2.stride, d2.pad, d2.transposed = d.kD, d.kH, d.kW, d.stride, d.pad, 0
      d2.x_cs, d2.x_co, d2.y_cs, d2.y_co = d.y_cs, d.y_co, d.x_cs, d.x_co
      d2.CinP, d2.CoutP, d2.y_planar, d2.bias_n_stride = d.CoutP, d.CinP, 0, 0
      return "dgrad_gt", "crn_conv_gemm_tc", (C.byref(d2), 0, dy, eng.gt_td[layer.name].data_ptr(), None, dx, acc, status, st)
    if kind == "wgrad" and layer.k == (5, 5, 5) and not layer.transposed and layer.name in eng.gt_wgrad:
      ok = eng.wl_ok.get(layer.name)
      if ok is None:                        # wide coarse Conv3d k=5: one staged image row serves the 5 kx taps
        ok = eng.wl_ok[layer.name] = bool(_lib.lib().crn_conv_wgrad_xline_supported(C.byref(d)))
      if ok:
        x, dy, dw, st = args
        return "wgrad_line", "crn_conv_wgrad_xline", (C.byref(d), x, dy, dw, status, st)
    if kind == "wgrad" and layer.name in eng.gt_wgrad:
      x, dy, dw, st = args
      return "wgrad_tc", "crn_conv_wgrad_tc", (C.byref(d), x, dy, dw, status, st)
    if kind == "wgrad" and layer.k == (7, 7, 7) and layer.transposed:
      ok = eng.wl_ok.get(layer.name)
      if ok is None:                        # ConvTranspose3d k=7 s=2, Cout == 16: class-channel variant of the same kernel
        ok = eng.wl_ok[layer.name] = bool(_lib.lib().crn_convt7_wgrad_line_supported(C.byref(d)))
      if ok:
        x, dy, dw, st = args
        return "wgrad_line", "crn_convt7_wgrad_line", (C.byref(d), x, dy, dw, status, st)
    if kind == "wgrad" and layer.k == (5, 5, 5) and not layer.transposed:
      ok = eng.wl_ok.get(layer.name)
      if ok is None:                        # narrow Conv3d k=5 layers: tap-stacked tcgen05 kernel (csrc/conv_wgrad_line.cu)
        ok = eng.wl_ok[layer.name] = bool(_lib.lib().crn_conv_wgrad_line_supported(C.byref(d)))
      if ok:
        x, dy, dw, st = args
        return "wgrad_line", "crn_conv_wgrad_line", (C.byref(d), x, dy, dw, status, st)
  return kind, _CONV_FN[kind], (C.byref(d),) + tuple(args)


def _wgrad_t7_slices(eng, layer, d, args):
  """ConvTranspose3d k=7 s=2 weight gradient of a layer wider than the class-channel tcgen05 kernel takes
  (Cin <= 32, Cout == 16): one launch per (32-channel Cin block, 16-channel Cout block), each writing its own block
  of the packed dW.  Returns the list of calls, or None when the layer does not decompose."""
  key = (layer.name, d.N)
  calls = eng.wl_split.get(key)
  if calls is None:
    calls = []
    if d.Cin <= 16 and 4 < d.Cout <= 16 and d.y_cs % 4 == 0 and d.CoutP % 4 == 0:
      # the C > 4 logits layer (channels-last gradient rows of pitch y_cs): 4-output-channel slices through the narrow
      # tap-stacked kernel; pad channels of the last slice carry zero gradient
      for co0 in range(0, _r4(d.Cout), 4):
        ds = ConvDesc.from_buffer_copy(bytes(d))
        ds.Cout, ds.y_co = 4, d.y_co + co0
        if _lib.lib().crn_convt7_wgrad_line_supported(C.byref(ds)) != 2:
          calls = []
          break
        da = ConvDesc.from_buffer_copy(bytes(ds))       # accounting: the real channels of the slice
        da.Cout = min(4, d.Cout - co0)
        calls.append((ds, 4 * co0, da))
    elif d.Cin % 32 == 0 and d.Cout % 16 == 0 and d.Cin <= 64 and d.Cout <= 32 and (d.Cin > 32 or d.Cout > 16):
      for ci0 in range(0, d.Cin, 32):
        for co0 in range(0, d.Cout, 16):
          ds = ConvDesc.from_buffer_copy(bytes(d))
          ds.Cin, ds.x_co, ds.Cout, ds.y_co = 32, d.x_co + ci0, 16, d.y_co + co0
          if not _lib.lib().crn_convt7_wgrad_line_supported(C.byref(ds)):
            calls = []
            break
          calls.append((ds, 4 * (ci0 * d.CoutP + co0), ds))
        if not calls:
          break
    eng.wl_split[key] = calls
  if not calls:
    return None
  x, dy, dw, st = args
  status = eng.tc_status.data_ptr()
  return [("wgrad_line", "crn_convt7_wgrad_line", (C.byref(ds), x, dy, dw + off, status, st), da) for ds, off, da in calls]


def conv_call(kind, layer, d, *args):
  calls = None
  if kind == "wgrad" and USE_TC and _ENG is not None and layer.transposed and layer.k == (7, 7, 7):
    calls = _wgrad_t7_slices(_ENG, layer, d, args)
  if calls is None:
    tag, fn, a = _conv_dispatch(kind, layer, d, args)
    calls = [(tag, fn, a, d)]
  for tag, fn, a, dd in calls:
    if NCU_PICK is not None and NCU_PICK(tag, layer.name):
      _ncu_bracket(fn, a)
      continue
    if PROFILE is None:
      _lib.call(fn, *a)
      continue
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call(fn, *a)
    e1.record()
    PROFILE.append((tag, layer.name, conv_macs(dd), e0, e1))


# The module path (CoreNet.forward + autograd) replays CUDA graphs: a step is ~600 launches whose Python/ctypes
# enqueue alone takes as long as their execution.  CRN_NO_GRAPH=1 (or an active PROFILE) keeps it eager.
USE_GRAPHS = os.environ.get("CRN_NO_GRAPH", "0") != "1"
GRAPH_WARMUP = 2


def graphs_enabled() -> bool:
  return USE_GRAPHS and PROFILE is None and not t.cuda.is_current_stream_capturing()


# Conv3d k=5 layers at >= 32^3 run forward/dgrad on the tcgen05 tensor cores (3xTF32, csrc/conv_tc5.cu).
USE_TC = True
# weight-gradient launches go to a second stream (see Plan.backward)
WGRAD_SIDE_STREAM = True
# stage_6.c1 forward through the kz-stacked kernel (csrc/conv_tc5s.cu)
USE_TC5S = True
# ray-traced skip backward as a sorted per-pixel gather (bit-reproducible) instead of float atomics
DETERMINISTIC_SKIP_BWD = True
# logits layer with > 4 classes through the padded channels-last rows path (see Engine._ensure_device)
USE_ROWS_LOGITS = True
# ConvTranspose3d k=7 dgrad with Cin <= 32 through the jz-stacked kernel (crn_convt7_tcs_dgrad)
USE_TCTS = True
# BatchRenorm of small maps through the single-launch cluster kernels (csrc/brn_fused.cu)
USE_FUSED_BRN = True

# Arithmetic of the tcgen05 convolution kernels.  "3xtf32" (default, fp32-class: every product is the three-term
# hi/lo TF32 split, see DESIGN.md "Precision") or "tf32" (single pass: only hi x hi is issued -- what cuDNN does with
# torch.backends.cudnn.allow_tf32 = True; measured in profiles/r02_precision_study.md).  The non-default mode is an
# opt-in: the reference's default arithmetic for Conv3d on CUDA is TF32 too, but the parity bar of this repo (logits
# <= 1e-3 against the fp64 oracle) is only met by "3xtf32".
PRECISION = "3xtf32"
_PRECISION_FLAG = 1 << 13


def set_precision(mode: str) -> None:
  """Selects the MMA arithmetic of every tcgen05 kernel (bit 13 of crn_set_flags).  Captured CUDA graphs bake the
  choice in; Plan / Trainer / Evaluator key their graphs on it, so a change re-captures on the next call."""
  global PRECISION
  if mode not in ("3xtf32", "tf32"):
    raise ValueError(f"precision must be '3xtf32' or 'tf32', got {mode!r}")
  PRECISION = mode
  _lib.lib().crn_set_flags(_PRECISION_FLAG if mode == "tf32" else 0)


def convt7_tc_call(layer, d, inp, wtc, bias, out, status, st, acct=None, fn="crn_convt7_tc"):
  """ConvTranspose3d k=7 s=2 forward through crn_convt7_tc (or the jz-stacked crn_convt7_tcs_fwd).  acct: descriptor
  whose MACs are reported (the launch descriptor may carry zero-padded output channels)."""
  tag = "fwd_tcs" if fn == "crn_convt7_tcs_fwd" else "fwd_tc"
  if NCU_PICK is not None and NCU_PICK(tag, layer.name):
    return _ncu_bracket(fn, (C.byref(d), inp, wtc, bias, out, status, st))
  if PROFILE is None:
    _lib.call(fn, C.byref(d), inp, wtc, bias, out, status, st)
    return
  e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
  e0.record()
  _lib.call(fn, C.byref(d), inp, wtc, bias, out, status, st)
  e1.record()
  PROFILE.append((tag, layer.name, conv_macs(acct or d), e0, e1))


def convt7_tc_dgrad_call(layer, d, dy, wtc, dx, status, st, acct=None, fn="crn_convt7_tc_dgrad"):
  """ConvTranspose3d k=7 s=2 dgrad through crn_convt7_tc_dgrad (or the jz-stacked crn_convt7_tcs_dgrad)."""
  tag = "dgrad_tcs" if fn == "crn_convt7_tcs_dgrad" else "dgrad_tc"
  if NCU_PICK is not None and NCU_PICK(tag, layer.name):
    return _ncu_bracket(fn, (C.byref(d), dy, wtc, dx, status, st))
  if PROFILE is None:
    _lib.call(fn, C.byref(d), dy, wtc, dx, status, st)
    return
  e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
  e0.record()
  _lib.call(fn, C.byref(d), dy, wtc, dx, status, st)
  e1.record()
  PROFILE.append((tag, layer.name, conv_macs(acct or d), e0, e1))


def conv5_tcs_call(layer, d, inp, wtc, bias, out, status, st, kind=0):
  """Conv3d k=5 forward (kind 0) / dgrad (kind 1) with <= 32 output channels through crn_conv5_tcs2 (kz taps
  stacked into N)."""
  if NCU_PICK is not None and NCU_PICK("fwd_tcs" if kind == 0 else "dgrad_tcs", layer.name):
    return _ncu_bracket("crn_conv5_tcs2", (C.byref(d), kind, inp, wtc, bias, out, status, st))
  if PROFILE is None:
    _lib.call("crn_conv5_tcs2", C.byref(d), kind, inp, wtc, bias, out, status, st)
    return
  e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
  e0.record()
  _lib.call("crn_conv5_tcs2", C.byref(d), kind, inp, wtc, bias, out, status, st)
  e1.record()
  PROFILE.append(("fwd_tcs" if kind == 0 else "dgrad_tcs", layer.name, conv_macs(d), e0, e1))


def conv5_tc_call(kind, layer, d, inp, wtc, bias, out, status, st):
  """kind: 'fwd' | 'dgrad' through crn_conv5_tc."""
  k = 0 if kind == "fwd" else 1
  if NCU_PICK is not None and NCU_PICK(kind + "_tc", layer.name):
    return _ncu_bracket("crn_conv5_tc", (C.byref(d), k, inp, wtc, bias, out, status, st))
  if PROFILE is None:
    _lib.call("crn_conv5_tc", C.byref(d), k, inp, wtc, bias, out, status, st)
    return
  e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
  e0.record()
  _lib.call("crn_conv5_tc", C.byref(d), k, inp, wtc, bias, out, status, st)
  e1.record()
  PROFILE.append((kind + "_tc", layer.name, conv_macs(d), e0, e1))


class Buf:
  """rows x C activation, channels-last with channel stride cs; .g is its gradient."""

  def __init__(self, dev, n, spatial, C_, cs=None, grad=True):
    self.n, self.spatial, self.C = n, tuple(spatial), C_
    self.cs = cs or _r4(C_)
    self.rows = n
    for s in spatial:
      self.rows *= s
    self.v = t.zeros(self.rows, self.cs, dtype=t.float32, device=dev)
    self.g = t.zeros_like(self.v) if grad else None

  @property
  def p(self):
    return self.v.data_ptr()

  @property
  def gp(self):
    return self.g.data_ptr()


class Slot:
  """n doubles inside one of the plan's accumulator arenas."""

  def __init__(self, plan, which, off, n):
    self.plan, self.which, self.off, self.n = plan, which, off, n

  def _arena(self):
    return self.plan.arena_fwd if self.which == "fwd" else self.plan.arena_bwd

  def data_ptr(self):
    return self._arena().data_ptr() + 8 * self.off

  def view(self):
    return self._arena()[self.off:self.off + self.n]


class ConvLayer:
  """One (transposed) convolution / linear layer bound to a parameter."""

  def __init__(self, eng, name, cin, cout, k, stride=1, pad=0, transposed=False, src_cin=None):
    self.name, self.cin, self.cout = name, cin, cout
    self.k = tuple(k)
    self.taps = k[0] * k[1] * k[2]
    self.stride, self.pad, self.transposed = stride, pad, transposed
    self.src_cin = src_cin or cin          # channels of the parameter's input dim
    self.cinp, self.coutp = _r4(max(self.src_cin, cin)), _r4(cout)
    self.size = self.taps * self.cinp * self.coutp
    self.off = eng._reserve(self)

  def desc(self, x_cs, idims, y_cs, odims, n, planar=False, bias_n_stride=0, x_co=0, y_co=0) -> ConvDesc:
    d = ConvDesc()
    d.N, d.Cin, d.Cout = n, self.cin, self.cout
    d.iD, d.iH, d.iW = idims
    d.oD, d.oH, d.oW = odims
    d.kD, d.kH, d.kW = self.k
    d.stride, d.pad, d.transposed = self.stride, self.pad, int(self.transposed)
    d.x_cs, d.x_co, d.y_cs, d.y_co = x_cs, x_co, y_cs, y_co
    d.CinP, d.CoutP = self.cinp, self.coutp
    d.y_planar, d.bias_n_stride = int(planar), bias_n_stride
    return d


class Engine:
  """Per-model state: layer table, weight arenas, cached plans."""

  def __init__(self, model):
    self.model = model
    self.layers: List[ConvLayer] = []
    self.total = 0
    self.dev = None
    self.plans: Dict = {}
    self._ptr_sig = None
    self._ver_sig = None
    self._pcache = None
    # bumped by anything that changes the weights through raw pointers (the fused Adam kernel, also inside replayed
    # CUDA graphs, never touches tensor._version): part of the re-pack signature
    self.weights_epoch = 0
    self._build_layers()

  # ------------------------------------------------------------------ layers
  def _reserve(self, layer):
    off = self.total
    self.total += layer.size
    self.layers.append(layer)
    return off

  def _build_layers(self):
    from corenet_b200.model import resnet50
    L = {}
    mk = lambda *a, **k: ConvLayer(self, *a, **k)
    L["stem"] = mk("encoder.stage1.conv", 4, 64, (1, 7, 7), stride=2, pad=3, src_cin=3)
    cin = 64
    for sname, letters, (f1, f2, f3), stride in resnet50.STAGES:
      for i, letter in enumerate(letters):
        p = f"encoder.{sname}.{letter}."
        s = stride if i == 0 else 1
        L[p + "op_a"] = mk(p + "op_a.conv", cin, f1, (1, 1, 1), stride=s)
        L[p + "op_b"] = mk(p + "op_b.conv", f1, f2, (1, 3, 3), pad=1)
        L[p + "op_c"] = mk(p + "op_c.conv", f2, f3, (1, 1, 1))
        if i == 0:
          L[p + "shortcut"] = mk(p + "shortcut.conv", cin, f3, (1, 1, 1), stride=s)
        cin = f3
    dc = self.model.config.decoder
    lat = dc.latent_channels
    self.lat = lat
    L["stage_0"] = mk("decoder.stage_0", 2048, lat, (1, 1, 1))
    L["stage_1.t1"] = mk("decoder.stage_1.t1", _r4(lat + 3), 256, (4, 4, 4), stride=4, transposed=True,
                         src_cin=lat + 3)
    self.dec_plan = []
    cin, grid = 256, 4
    for stage, k, kt, pt, mid, t_out, enc_c in DECODER_PYRAMID:
      last = t_out is None
      t_out = dc.num_output_channels if last else t_out
      L[f"stage_{stage}.c1"] = mk(f"decoder.stage_{stage}.c1", cin, mid, (k, k, k), pad=k // 2)
      L[f"stage_{stage}.t1"] = mk(f"decoder.stage_{stage}.t1", mid, t_out, (kt, kt, kt), stride=2, pad=pt,
                                  transposed=True)
      skip_c = 0
      if not last:
        skip_c = round(t_out * dc.skip_fraction)
        if skip_c > 0:
          L[f"rt_skip_{stage}"] = mk(f"decoder.rt_skip_{stage}.compress_channels", enc_c, skip_c, (1, 1, 1),
                                     src_cin=enc_c + 3)
      self.dec_plan.append((stage, cin, mid, t_out, skip_c, enc_c, grid))
      grid *= 2
      cin = t_out + skip_c
    self.L = L

  # ------------------------------------------------------------------ arenas
  def _ensure_device(self, dev):
    if self.dev == dev:
      return
    self.dev = dev
    self.w_fwd = t.zeros(self.total, dtype=t.float32, device=dev)
    self.w_dgrad = t.zeros(self.total, dtype=t.float32, device=dev)
    self.dw = t.zeros(self.total, dtype=t.float32, device=dev)
    # tcgen05 path: per eligible Conv3d(k=5) layer a pre-split (hi/lo) packed copy for fwd and for dgrad
    self.tc_status = t.zeros(1, dtype=t.int32, device=dev)
    self.tc_w = {}
    lib = _lib.lib()
    for stage, cin, mid, t_out, skip_c, enc_c, g in self.dec_plan:
      l = self.L[f"stage_{stage}.c1"]
      if l.k == (5, 5, 5) and g >= 32 and g % 16 == 0 and cin <= 64 and mid <= 64:
        self.tc_w[l.name] = (t.zeros(lib.crn_tc5_packed_floats(cin, mid), dtype=t.float32, device=dev),
                             t.zeros(lib.crn_tc5_packed_floats(mid, cin), dtype=t.float32, device=dev))
    # forward only (Cout <= 64, any Cin) on 16^3 grids: stage_4.c1 (112 -> 64).  The plane kernel stages an input plane
    # once per kz and reads its 25 (ky, kx) taps from shared memory; the implicit-GEMM kernel re-gathers every tap from L2
    # (0.64 ms at 46 TFLOP/s).  One output plane per work item fills the machine at this size (crn_conv5_tc).
    self.tc_wf = {}
    for stage, cin, mid, t_out, skip_c, enc_c, g in self.dec_plan:
      l = self.L[f"stage_{stage}.c1"]
      if (USE_TC and l.name not in self.tc_w and l.k == (5, 5, 5) and g == 16 and mid <= 64 and mid % 4 == 0
          and cin % 4 == 0):
        self.tc_wf[l.name] = t.zeros(lib.crn_tc5_packed_floats(cin, mid), dtype=t.float32, device=dev)
    # forward of the <= 16-output-channel layers: kz taps stacked into N (csrc/conv_tc5s.cu)
    # (<= 32 output channels; the dgrad of a layer is the same kernel with N = Cin)
    self.tcs_w = {}
    self.tcs_wd = {}
    self.tcs_wd_slices = {}
    if USE_TC5S:
      for stage, cin, mid, t_out, skip_c, enc_c, g in self.dec_plan:
        l = self.L[f"stage_{stage}.c1"]
        if l.name in self.tc_w and mid % 4 == 0 and cin % 4 == 0 and g % 16 == 0:
          if mid <= 32:
            self.tcs_w[l.name] = t.zeros(lib.crn_tc5s_packed_floats(cin), dtype=t.float32, device=dev)
          if cin <= 32:
            self.tcs_wd[l.name] = t.zeros(lib.crn_tc5s_packed_floats(mid), dtype=t.float32, device=dev)
          elif cin <= 64 and cin % 8 == 0:
            # dgrad with 32 < Cin <= 64 input channels (stage_5.c1: 56): two launches of the stacked kernel, each on one
            # half of the input channels (weight slice packed per half, dx written at the half's channel offset)
            h = cin // 2
            self.tcs_wd_slices[l.name] = [
                (t.zeros(lib.crn_tc5s_packed_floats(mid), dtype=t.float32, device=dev),
                 t.zeros(mid, h, 5, 5, 5, dtype=t.float32, device=dev), c0, h) for c0 in (0, h)]
    # ... and per eligible ConvTranspose3d(k=7, s=2) layer a packed copy for the forward (csrc/conv_tc5.cu, KT=4)
    self.tct_w = {}
    self.tct_slices = {}
    for stage, cin, mid, t_out, skip_c, enc_c, g in self.dec_plan:
      l = self.L[f"stage_{stage}.t1"]
      if l.k == (7, 7, 7) and g >= 16 and g % 16 == 0 and mid % 4 == 0:
        fwd_ok = t_out <= 16                          # 8 * Cout accumulator columns <= 128
        # float4 class-channel gathers, N = Cin <= 64 (on small grids the C side takes one z-plane per work item)
        dgrad_ok = t_out % 4 == 0 and mid <= 64
        mk = lambda dg: t.zeros(lib.crn_tct_packed_floats(mid, t_out, dg), dtype=t.float32, device=dev)
        if fwd_ok or dgrad_ok:
          self.tct_w[l.name] = (mk(0) if fwd_ok else None, mk(1) if dgrad_ok else None)
        # wider outputs: the forward runs once per 16-channel slice of Cout (each slice = 128 accumulator columns)
        if not fwd_ok and t_out % 16 == 0 and t_out <= 64:
          self.tct_slices[l.name] = [
              (t.zeros(lib.crn_tct_packed_floats(mid, 16, 0), dtype=t.float32, device=dev),
               t.zeros(mid, 16, 7, 7, 7, dtype=t.float32, device=dev), co0) for co0 in range(0, t_out, 16)]
    # logits layer with more than 4 classes (SEMANTIC task, C = 15): the class-scatter kernels run on the channel count
    # padded to a multiple of 4 (zero weights / bias in the pad channels) and write / read channels-last ROWS
    # [B*128^3, cp]: float4 epilogues instead of per-channel planar scalars, tcgen05 dgrad (needs Cout % 4 == 0), and
    # the training step never materialises planar logits or logit gradients (the loss kernels take rows)
    self.rows_pad = {}
    stage, cin, mid, t_out, skip_c, enc_c, g = self.dec_plan[-1]
    l = self.L[f"stage_{stage}.t1"]
    cp = _r4(t_out)
    if USE_TC and USE_ROWS_LOGITS and l.k == (7, 7, 7) and 4 < t_out and cp <= 16 and g % 16 == 0 and mid % 4 == 0 and mid <= 64:
      self.rows_pad[l.name] = dict(
          cp=cp, w=t.zeros(mid, cp, 7, 7, 7, dtype=t.float32, device=dev), b=t.zeros(cp, dtype=t.float32, device=dev),
          fwd=t.zeros(lib.crn_tct_packed_floats(mid, cp, 0), dtype=t.float32, device=dev),
          dgrad=t.zeros(lib.crn_tct_packed_floats(mid, cp, 1), dtype=t.float32, device=dev))
      self.tct_w.pop(l.name, None)
    # transposed-conv dgrad with the jz taps stacked into N (csrc/conv_tc5s.cu, KT = 4): Cin <= 32, Cout % 4 == 0
    self.tcts_wd = {}
    if USE_TC and USE_TCTS:
      for stage, cin, mid, t_out, skip_c, enc_c, g in self.dec_plan:
        l = self.L[f"stage_{stage}.t1"]
        co = self.rows_pad[l.name]["cp"] if l.name in self.rows_pad else t_out
        if l.k == (7, 7, 7) and g % 16 == 0 and mid % 4 == 0 and mid <= 32 and co % 4 == 0 and (
            l.name in self.rows_pad or self.tct_w.get(l.name, (None, None))[1] is not None):
          self.tcts_wd[l.name] = (t.zeros(lib.crn_tcts_packed_floats(co), dtype=t.float32, device=dev), co)
        elif (l.k == (7, 7, 7) and g % 16 == 0 and mid % 4 == 0 and mid <= 32 and t_out == 2 and stage == 6
              and l.name not in self.rows_pad):
          # FG_BG logits layer: the class-channel gather reads the PLANAR logit gradient (float2 per channel plane)
          self.tcts_wd[l.name] = (t.zeros(lib.crn_tcts_packed_floats(2), dtype=t.float32, device=dev), 2)
    # forward of the FG_BG logits layer (Cout <= 2) with the jz taps stacked into N (crn_convt7_tcs_fwd)
    self.tctsf_w = {}
    if USE_TC and USE_TCTS:
      stage, cin, mid, t_out, skip_c, enc_c, g = self.dec_plan[-1]
      l = self.L[f"stage_{stage}.t1"]
      if l.k == (7, 7, 7) and t_out <= 2 and g % 16 == 0 and mid % 4 == 0:
        self.tctsf_w[l.name] = t.zeros(lib.crn_tctsf_packed_floats(mid), dtype=t.float32, device=dev)
    # wide layers (>= 32 channels on both sides): implicit-GEMM forward / dgrad (csrc/conv_gemm_tc.cu) and weight
    # gradient (csrc/conv_wgrad_tc.cu) on tcgen05
    self.gt_w = {}
    self.gt_wgrad = set()
    self.wl_ok = {}
    self.wl_split = {}
    self.gt_td = {}
    self.gt_tf = {}
    for l in self.layers:
      wide = min(l.cin, l.cout) >= 32 and l.cin % 4 == 0 and l.cout % 4 == 0 and l.src_cin == l.cin
      enc = l.name.startswith("encoder.") and l.name != "encoder.stage1.conv"
      dec_c1 = l.name.startswith("decoder.stage_") and l.name.endswith(".c1") and l.name not in self.tc_w
      # 4.t1 (64 -> 32, 343 taps): the per-tap gather is L2-bound at these narrow channel counts (3.0 ms vs 1.3 FFMA)
      dec_t1 = l.name in ("decoder.stage_2.t1", "decoder.stage_3.t1")
      if wide and not l.transposed and (enc or dec_c1):
        mk = lambda k, n: t.zeros(lib.crn_gemm_tc_packed_floats(k, n, l.taps), dtype=t.float32, device=dev)
        self.gt_w[l.name] = (mk(l.cin, l.cout), mk(l.cout, l.cin))     # strided dgrad = transposed-gather mode
        self.gt_wgrad.add(l.name)
      elif wide and l.transposed and dec_t1:
        self.gt_wgrad.add(l.name)
      # transposed-conv forward through the class-ordered transposed-gather mode of the implicit-GEMM kernel
      if (wide and l.transposed and l.name.startswith("decoder.stage_") and l.stride == 2
          and self.tct_w.get(l.name, (None, None))[0] is None and l.name not in self.tct_slices):
        self.gt_tf[l.name] = t.zeros(lib.crn_gemm_tc_packed_floats(l.cin, l.cout, l.taps), dtype=t.float32, device=dev)
      # transposed-conv dgrad = strided conv of dy: wide layers that the class-channel tcgen05 kernel does not take
      if (wide and l.transposed and l.name.startswith("decoder.stage_") and l.stride == 2
          and self.tct_w.get(l.name, (None, None))[1] is None):
        self.gt_td[l.name] = t.zeros(lib.crn_gemm_tc_packed_floats(l.cout, l.cin, l.taps), dtype=t.float32, device=dev)
    self._gt_sig = None
    self._pack_stream = None
    self._pack_ev = None
    self._pack_ev_bwd = None
    self.plans = {}
    self._ptr_sig = None
    self._ver_sig = None

  def _pack_gemm_tc(self, P, part):
    """Re-packs the tcgen05 implicit-GEMM weights with one launch per part: "fwd" = the buffers the forward pass reads
    (joined before the encoder blocks), "bwd" = the dgrad buffers (first read by the backward pass: their pack overlaps
    the forward pass).  Item lists are cached per parameter pointer set."""
    if not self.gt_w and not self.gt_td and not self.gt_tf:
      return
    names = sorted(self.gt_w)
    sig = tuple(P[n + ".weight"].data_ptr() for n in names + sorted(self.gt_td) + sorted(self.gt_tf))
    if sig != self._gt_sig:
      lay = {l.name: l for l in self.layers}
      entries = {"fwd": [], "bwd": []}
      for n in names:
        for dg, buf in enumerate(self.gt_w[n]):
          if buf is not None:
            entries["bwd" if dg else "fwd"].append((lay[n], P[n + ".weight"], dg, buf, False))
      for n in sorted(self.gt_td):      # (layer, weight, dgrad flag, buffer, swap): convT weight read as [Cout'=Cin][Cin'=Cout]
        entries["bwd"].append((lay[n], P[n + ".weight"], 0, self.gt_td[n], True))
      for n in sorted(self.gt_tf):      # convT forward: K = Cin, N = Cout, flipped taps (transposed-gather mode)
        entries["fwd"].append((lay[n], P[n + ".weight"], 1, self.gt_tf[n], True))
      self._gt_parts = {}
      for key, ent in entries.items():
        if not ent:
          continue
        items = (_lib.GemmTcPackItem * len(ent))()
        offs = (C.c_int64 * (len(ent) + 1))()
        tot = 0
        for i, (l, w, dg, buf, swap) in enumerate(ent):
          assert w.is_contiguous() and w.dtype == t.float32 and w.device == self.dev
          it = items[i]
          co, ci = (l.cin, l.cout) if swap else (l.cout, l.cin)
          it.src, it.dst, it.Cout, it.Cin, it.taps, it.dgrad = w.data_ptr(), buf.data_ptr(), co, ci, l.taps, dg
          offs[i] = tot
          tot += buf.numel() // 2
        offs[len(ent)] = tot
        self._gt_parts[key] = (self._to_dev(items, self.dev), self._to_dev(offs, self.dev), len(ent), tot)
      self._gt_sig = sig
    if part in self._gt_parts:
      items, offs, n, tot = self._gt_parts[part]
      _call("crn_gemm_tc_pack", items.data_ptr(), offs.data_ptr(), n, tot, _lib.stream_ptr())

  def tensors(self):
    """(name -> parameter, name -> buffer), cached."""
    if self._pcache is None:
      self._pcache = dict(self.model.named_parameters())
      self._bcache = dict(self.model.named_buffers())
    return self._pcache, self._bcache

  def invalidate(self):
    self._pcache = None
    self._ptr_sig = None
    self._gt_sig = None
    self.__dict__.pop("_ucache", None)

  @staticmethod
  def _to_dev(ctypes_array, dev):
    return t.frombuffer(bytearray(bytes(ctypes_array)), dtype=t.uint8).to(dev)

  def pack_weights(self):
    """One launch: all parameters -> tap-major fwd + dgrad arenas (skipped when unchanged)."""
    P, _ = self.tensors()
    ws = [P[l.name + ".weight"] for l in self.layers]
    ptr_sig = tuple(w.data_ptr() for w in ws)
    ver_sig = (self.weights_epoch,) + tuple(w._version for w in ws)
    if ptr_sig != self._ptr_sig:
      entries = []
      for l, w in zip(self.layers, ws):
        assert w.is_contiguous() and w.dtype == t.float32 and w.device == self.dev
        # layers served by the tcgen05 implicit-GEMM kernels do not read the FFMA arenas
        tc_f = USE_TC and l.name in self.gt_w
        tc_d = USE_TC and (self.gt_w.get(l.name, (None, None))[1] is not None or l.name in self.gt_td)
        if not (tc_f and tc_d):
          entries.append((l, w, tc_f, tc_d))
      items = (PackItem * len(entries))()
      offs = (C.c_int64 * (len(entries) + 1))()
      tot = 0
      for i, (l, w, tc_f, tc_d) in enumerate(entries):
        it = items[i]
        it.src = w.data_ptr()
        it.dst_fwd = None if tc_f else self.w_fwd.data_ptr() + 4 * l.off
        it.dst_dgrad = None if tc_d else self.w_dgrad.data_ptr() + 4 * l.off
        it.Cin, it.Cout, it.taps, it.CinP, it.CoutP = l.src_cin, l.cout, l.taps, l.cinp, l.coutp
        it.src_is_transposed = int(l.transposed)
        offs[i] = tot
        tot += l.size
      offs[len(entries)] = tot
      self._items_dev = self._to_dev(items, self.dev)
      self._offs_dev = self._to_dev(offs, self.dev)
      self._pack_n, self._pack_tot = len(entries), tot
      self._ptr_sig = ptr_sig
      self._ver_sig = None
    if ver_sig != self._ver_sig:
      _call("crn_pack_weights", self._items_dev.data_ptr(), self._offs_dev.data_ptr(), self._pack_n,
            self._pack_tot, _lib.stream_ptr())
      # the tcgen05 re-packs (~0.5 ms) are first needed by the encoder blocks / the decoder: they run on a side
      # stream concurrently with preprocessing, the stem and its BatchRenorm (Plan.forward joins before the blocks)
      main = t.cuda.current_stream()
      if self._pack_stream is None:
        self._pack_stream = t.cuda.Stream(device=self.dev)
      ev = t.cuda.Event()
      ev.record(main)
      self._pack_stream.wait_event(ev)
      with t.cuda.stream(self._pack_stream):
        self._pack_tc(P)
        self._pack_ev = t.cuda.Event()
        self._pack_ev.record(self._pack_stream)
        if USE_TC and self.layers:
          self._pack_gemm_tc(P, "bwd")
          self._pack_ev_bwd = t.cuda.Event()
          self._pack_ev_bwd.record(self._pack_stream)
      self._ver_sig = ver_sig

  def join_packs(self, fwd_only=False):
    """Main stream waits for the side-stream weight re-packs of this forward (no-op if none were launched).
    fwd_only: only for the buffers the forward pass reads; Plan.forward joins the dgrad packs when it ends."""
    if self._pack_ev is not None:
      t.cuda.current_stream().wait_event(self._pack_ev)
      self._pack_ev = None
    if not fwd_only and self._pack_ev_bwd is not None:
      t.cuda.current_stream().wait_event(self._pack_ev_bwd)
      self._pack_ev_bwd = None

  def _pack_tc(self, P):
    """tcgen05 weight re-packs (conv_tc5 / class-scatter / implicit-GEMM layouts) on the current stream."""
    if self.layers:
      for l in self.layers:
        if l.name in self.tc_wf:
          _call("crn_tc5_pack", P[l.name + ".weight"].data_ptr(), l.cout, l.cin, 0, self.tc_wf[l.name].data_ptr(),
                _lib.stream_ptr())
        if l.name in self.tc_w:
          w = P[l.name + ".weight"]
          wf, wd = self.tc_w[l.name]
          if l.name in self.tcs_w:
            _call("crn_tc5s_pack2", w.data_ptr(), l.cout, l.cin, 0, self.tcs_w[l.name].data_ptr(), _lib.stream_ptr())
          else:
            _call("crn_tc5_pack", w.data_ptr(), l.cout, l.cin, 0, wf.data_ptr(), _lib.stream_ptr())
          if l.name in self.tcs_wd:
            _call("crn_tc5s_pack2", w.data_ptr(), l.cout, l.cin, 1, self.tcs_wd[l.name].data_ptr(), _lib.stream_ptr())
          elif l.name in self.tcs_wd_slices:
            for buf, wslice, c0, h in self.tcs_wd_slices[l.name]:
              wslice.copy_(w[:, c0:c0 + h])
              _call("crn_tc5s_pack2", wslice.data_ptr(), l.cout, h, 1, buf.data_ptr(), _lib.stream_ptr())
          else:
            _call("crn_tc5_pack", w.data_ptr(), l.cout, l.cin, 1, wd.data_ptr(), _lib.stream_ptr())
        for wt, wslice, co0 in self.tct_slices.get(l.name, ()):
          wslice.copy_(P[l.name + ".weight"][:, co0:co0 + 16])
          _call("crn_tct_pack", wslice.data_ptr(), l.cin, 16, 0, wt.data_ptr(), _lib.stream_ptr())
        if l.name in self.rows_pad:
          rp = self.rows_pad[l.name]
          rp["w"][:, :l.cout].copy_(P[l.name + ".weight"])
          rp["b"][:l.cout].copy_(P[l.name + ".bias"])
          _call("crn_tct_pack", rp["w"].data_ptr(), l.cin, rp["cp"], 0, rp["fwd"].data_ptr(), _lib.stream_ptr())
          _call("crn_tct_pack", rp["w"].data_ptr(), l.cin, rp["cp"], 1, rp["dgrad"].data_ptr(), _lib.stream_ptr())
        if l.name in self.tctsf_w:
          _call("crn_tctsf_pack", P[l.name + ".weight"].data_ptr(), l.cin, l.cout, self.tctsf_w[l.name].data_ptr(),
                _lib.stream_ptr())
        if l.name in self.tcts_wd:
          wsrc = self.rows_pad[l.name]["w"] if l.name in self.rows_pad else P[l.name + ".weight"]
          buf, co = self.tcts_wd[l.name]
          _call("crn_tcts_pack", wsrc.data_ptr(), l.cin, co, buf.data_ptr(), _lib.stream_ptr())
        if l.name in self.tct_w:
          for dg, wt in enumerate(self.tct_w[l.name]):
            if wt is not None:
              _call("crn_tct_pack", P[l.name + ".weight"].data_ptr(), l.cin, l.cout, dg, wt.data_ptr(),
                    _lib.stream_ptr())
      self._pack_gemm_tc(P, "fwd")

  def chunk_layers(self, ci: int):
    """Conv layers of gradient chunk ci (GRAD_CHUNKS order)."""
    if getattr(self, "_chunk_layers", None) is None:
      self._chunk_layers = [[l for l in self.layers if grad_chunk_of(l.name) == c] for c in range(len(GRAD_CHUNKS))]
    return self._chunk_layers[ci]

  def unpack_wgrads(self, grads: Dict[str, t.Tensor], layers=None, key=-1):
    layers = self.layers if layers is None else layers
    sig = tuple(grads[l.name + ".weight"].data_ptr() for l in layers)
    cache = self.__dict__.setdefault("_ucache", {})
    ent = cache.get(key)
    if ent is None or ent[0] != sig:                 # item list is cached per destination set (graph-capture safe)
      items = (UnpackItem * len(layers))()
      offs = (C.c_int64 * (len(layers) + 1))()
      tot = 0
      for i, l in enumerate(layers):
        it = items[i]
        it.src_packed = self.dw.data_ptr() + 4 * l.off
        it.dst = grads[l.name + ".weight"].data_ptr()
        it.Cin, it.Cout, it.taps, it.CinP, it.CoutP = l.src_cin, l.cout, l.taps, l.cinp, l.coutp
        it.dst_is_transposed = int(l.transposed)
        offs[i] = tot
        tot += l.src_cin * l.cout * l.taps
      offs[len(layers)] = tot
      ent = cache[key] = (sig, self._to_dev(items, self.dev), self._to_dev(offs, self.dev), tot)
    _call("crn_unpack_wgrads", ent[1].data_ptr(), ent[2].data_ptr(), len(layers), ent[3], _lib.stream_ptr())

  def wf(self, l):
    return self.w_fwd.data_ptr() + 4 * l.off

  def wd(self, l):
    return self.w_dgrad.data_ptr() + 4 * l.off

  def dwp(self, l):
    return self.dw.data_ptr() + 4 * l.off

  def get_plan(self, batch: int, dev, need_grad: bool):
    self._ensure_device(dev)
    pool = self.plans.setdefault((batch, need_grad), [])
    for p in pool:
      if not p.busy:
        return p
    p = Plan(self, batch, need_grad)
    pool.append(p)
    return p


# ====================================================================== plan
class BRNOp:
  """One BatchRenorm instance bound to buffers inside a plan."""

  def __init__(self, plan, name, x: Buf, C_, relu_in, y: Buf, res: Optional[Buf] = None, relu_out=False,
               y_pre: Optional[Buf] = None):
    self.plan, self.name, self.x, self.C = plan, name, x, C_
    self.relu_in, self.relu_out, self.y, self.res, self.y_pre = relu_in, relu_out, y, res, y_pre
    self.coef = plan.f32(6 * C_)
    self.acc = plan.slot("fwd", 3 * C_)
    self.bacc = plan.slot("bwd", 2 * C_)
    self.dxsum = plan.slot("bwd", C_)
    # small maps: statistics + coefficients + apply (and reduce + dx) in ONE launch each (csrc/brn_fused.cu)
    self.fused = bool(USE_FUSED_BRN and _lib.lib().crn_brn_fused_supported(x.rows, C_) and x.cs % 4 == 0
                      and y.cs % 4 == 0)
    self.snap_index = plan.register_fused_brn(self) if self.fused else -1

  def fwd(self, training):
    P, Bf = self.plan.eng.tensors()
    st = _lib.stream_ptr()
    n, x = self.name, self.x
    if self.fused:
      _call("crn_brn_fwd_fused", x.p, x.rows, self.C, x.cs, 0, int(self.relu_in), P[n + ".weight"].data_ptr(),
            P[n + ".bias"].data_ptr(), Bf[n + ".running_mean"].data_ptr(), Bf[n + ".running_var"].data_ptr(),
            self.plan.nbt_snapshot.data_ptr() + 8 * self.snap_index if training else None, BRN_EPS, BRN_MOMENTUM,
            int(training), self.res.p if self.res is not None else None, int(self.relu_out), self.y.p, self.y.cs, 0,
            self.y_pre.p if self.y_pre is not None else None, self.coef.data_ptr(), st,
            hbm=("brn_fwd_fused", (12 + (4 if self.res is not None else 0) + (4 if self.y_pre is not None else 0))
                 * x.rows * self.C))
      return
    if training:
      _call("crn_brn_stats", x.p, x.rows, self.C, x.cs, 0, int(self.relu_in), self.acc.data_ptr(), st)
    _call("crn_brn_finalize", self.acc.data_ptr(), x.rows, self.C, P[n + ".weight"].data_ptr(),
          P[n + ".bias"].data_ptr(), Bf[n + ".running_mean"].data_ptr(), Bf[n + ".running_var"].data_ptr(),
          Bf[n + ".num_batches_tracked"].data_ptr(), BRN_EPS, BRN_MOMENTUM, int(training),
          self.coef.data_ptr(), st)
    _call("crn_brn_apply", x.p, x.rows, self.C, x.cs, 0, self.coef.data_ptr(),
          self.res.p if self.res is not None else None, int(self.relu_in), int(self.relu_out), self.y.p,
          self.y.cs, 0, self.y_pre.p if self.y_pre is not None else None, st)

  def bwd(self, training, grads, dy_ptr, dy_cs, dx_ptr, dx_cs, g_extra_ptr=None, g_store_ptr=None):
    """dy: gradient wrt this op's (activated) output.  The masked / combined
    gradient is written to g_store (defaults to in-place over dy when a mask or
    extra term exists).  dx: gradient wrt the op's input x.  Returns the Slot
    holding column sums of dx (= bias gradient of the conv that produced x)."""
    st = _lib.stream_ptr()
    n, x = self.name, self.x
    g_out = g_store_ptr
    if g_out is None and (self.relu_out or g_extra_ptr is not None):
      g_out = dy_ptr
    if self.fused:
      _call("crn_brn_bwd_fused", dy_ptr, dy_cs, 0, self.y.p if self.relu_out else None, g_extra_ptr, x.p, x.cs, 0,
            x.rows, self.C, self.coef.data_ptr(), int(self.relu_in), int(self.relu_out), int(training), g_out, dx_ptr,
            dx_cs, 0, 0, grads[n + ".weight"].data_ptr(), grads[n + ".bias"].data_ptr(), self.dxsum.data_ptr(), st,
            hbm=("brn_bwd_fused", 28 * x.rows * self.C))
      return self.dxsum
    _call("crn_brn_bwd_reduce", dy_ptr, dy_cs, 0, self.y.p if self.relu_out else None, g_extra_ptr, x.p,
          x.cs, 0, x.rows, self.C, self.coef.data_ptr(), int(self.relu_in), int(self.relu_out), g_out,
          self.bacc.data_ptr(), st)
    g = g_out if g_out is not None else dy_ptr
    _call("crn_brn_bwd_dx", g, dy_cs, 0, x.p, x.cs, 0, x.rows, self.C, self.coef.data_ptr(),
          self.bacc.data_ptr(), None, int(self.relu_in), int(training), dx_ptr, dx_cs, 0, 0,
          grads[n + ".weight"].data_ptr(), grads[n + ".bias"].data_ptr(), self.dxsum.data_ptr(), st)
    return self.dxsum


class Plan:
  """All buffers + the launch sequence for one batch size."""

  def __init__(self, eng: Engine, batch: int, need_grad: bool):
    self.eng, self.B, self.need_grad = eng, batch, need_grad
    self.dev = eng.dev
    self.busy = False
    self._n = {"fwd": 0, "bwd": 0}
    self.fused_brn = []    # BRNOp instances on the single-launch path, in construction order (encoder first)
    self.rows_cp = 0       # > 0: the logits layer works on channels-last rows of this pitch (see Engine.rows_pad)
    self._build()
    self.arena_fwd = t.zeros(max(self._n["fwd"], 1), dtype=t.float64, device=self.dev)
    self.arena_bwd = t.zeros(max(self._n["bwd"], 1), dtype=t.float64, device=self.dev)
    self.training = True
    self.launches = 0
    self.side_stream = t.cuda.Stream(device=self.dev)
    self.gen = 0
    # the module path (CoreNet.forward / autograd backward) replays captured CUDA graphs after GRAPH_WARMUP eager calls
    self._graphs = {}
    self.in_image = t.zeros(batch, 3, 256, 256, dtype=t.uint8, device=self.dev)
    self.in_v2s = t.zeros(batch, 4, 4, dtype=t.float32, device=self.dev)
    self.in_offs = t.zeros(batch, 3, dtype=t.float32, device=self.dev)
    self.in_glog = None
    self.gflat = None

  def f32(self, n):
    return t.zeros(n, dtype=t.float32, device=self.dev)

  def register_fused_brn(self, op) -> int:
    self.fused_brn.append(op)
    return len(self.fused_brn) - 1

  def _snapshot_nbt(self, lo, hi):
    """ONE launch: copies num_batches_tracked of the fused BatchRenorm instances [lo, hi) into the plan's snapshot
    array and increments the originals (batch_renorm.py:58); the pointer list is cached per buffer set."""
    if hi <= lo:
      return
    _, Bf = self.eng.tensors()
    ptrs = tuple(Bf[op.name + ".num_batches_tracked"].data_ptr() for op in self.fused_brn[lo:hi])
    cache = self.__dict__.setdefault("_nbt_cache", {})
    ent = cache.get((lo, hi))
    if ent is None or ent[0] != ptrs:
      arr = (C.c_int64 * len(ptrs))(*ptrs)
      ent = cache[(lo, hi)] = (ptrs, self.eng._to_dev(arr, self.dev))
    _call("crn_brn_nbt_snapshot", ent[1].data_ptr(), hi - lo, self.nbt_snapshot.data_ptr() + 8 * lo, _lib.stream_ptr())

  def slot(self, which, n):
    s = Slot(self, which, self._n[which], n)
    self._n[which] += n
    return s

  def buf(self, spatial, C_, cs=None, grad=None):
    return Buf(self.dev, self.B, spatial, C_, cs, self.need_grad if grad is None else grad)

  # ------------------------------------------------------------------ build
  def _build(self):
    from corenet_b200.model import resnet50
    eng, B, L = self.eng, self.B, self.eng.L
    R = 256
    self.img4 = self.buf((R, R), 4, grad=False)
    self.s1 = self.buf((128, 128), 64)
    self.s1a = self.buf((128, 128), 64)
    self.brn_stem = BRNOp(self, "encoder.stage1_part2.bn", self.s1, 64, False, self.s1a, relu_out=True)
    self.p1 = self.buf((64, 64), 64)
    self.p1_idx = t.zeros(self.p1.rows * 64, dtype=t.int8, device=self.dev)
    self.d_stem = L["stem"].desc(4, (1, R, R), 64, (1, 128, 128), B)
    self.blocks = []
    x, hw = self.p1, 64
    self.enc_pre = {}
    for sname, letters, (f1, f2, f3), stride in resnet50.STAGES:
      pre = None
      for i, letter in enumerate(letters):
        p = f"encoder.{sname}.{letter}."
        s = stride if i == 0 else 1
        ohw = hw // s
        blk = {"p": p, "x": x, "down": i == 0}
        a_c, a_y = self.buf((ohw, ohw), f1), self.buf((ohw, ohw), f1)
        b_c, b_y = self.buf((ohw, ohw), f2), self.buf((ohw, ohw), f2)
        c_c, out = self.buf((ohw, ohw), f3), self.buf((ohw, ohw), f3)
        pre = self.buf((ohw, ohw), f3) if i == len(letters) - 1 else None
        blk.update(a_c=a_c, a_y=a_y, b_c=b_c, b_y=b_y, c_c=c_c, out=out, pre=pre)
        blk["la"], blk["lb"], blk["lc"] = L[p + "op_a"], L[p + "op_b"], L[p + "op_c"]
        blk["d_a"] = blk["la"].desc(x.cs, (1, hw, hw), a_c.cs, (1, ohw, ohw), B)
        blk["d_b"] = blk["lb"].desc(a_y.cs, (1, ohw, ohw), b_c.cs, (1, ohw, ohw), B)
        blk["d_c"] = blk["lc"].desc(b_y.cs, (1, ohw, ohw), c_c.cs, (1, ohw, ohw), B)
        blk["bn_a"] = BRNOp(self, p + "op_a.bn", a_c, f1, False, a_y, relu_out=True)
        blk["bn_b"] = BRNOp(self, p + "op_b.bn", b_c, f2, False, b_y, relu_out=True)
        if i == 0:
          s_c, s_y = self.buf((ohw, ohw), f3), self.buf((ohw, ohw), f3)
          blk.update(s_c=s_c, s_y=s_y, ls=L[p + "shortcut"])
          blk["d_s"] = blk["ls"].desc(x.cs, (1, hw, hw), s_c.cs, (1, ohw, ohw), B)
          blk["bn_s"] = BRNOp(self, p + "shortcut.bn", s_c, f3, False, s_y)
          res = s_y
        else:
          res = x
        blk["bn_c"] = BRNOp(self, p + "op_c.bn", c_c, f3, False, out, res=res, relu_out=True, y_pre=pre)
        self.blocks.append(blk)
        x, hw = out, ohw
      self.enc_pre[sname] = pre
    self.enc_out = x                      # [B, 8, 8, 2048] post-ReLU
    self.feat = Buf(self.dev, B, (), 2048, grad=self.need_grad)
    # ---- decoder
    lat = eng.lat
    self.latb = Buf(self.dev, B, (), lat + 3, grad=self.need_grad)
    self.z1 = Buf(self.dev, B, (), lat + 3, grad=self.need_grad)
    self.d_s0 = L["stage_0"].desc(2048, (1, 1, 1), self.latb.cs, (1, 1, 1), B)
    self.brn_s1 = BRNOp(self, "decoder.stage_1.b1", self.latb, lat + 3, True, self.z1)
    self.x1 = self.buf((4, 4, 4), 256)
    self.d_s1t = L["stage_1.t1"].desc(self.z1.cs, (1, 1, 1), 256, (4, 4, 4), B)
    self.stages = []
    cat = self.x1
    for stage, cin, mid, t_out, skip_c, enc_c, g in eng.dec_plan:
      st = {"stage": stage, "cat": cat, "g": g, "cin": cin, "mid": mid, "t_out": t_out, "skip_c": skip_c}
      z, c, z2 = self.buf((g, g, g), cin), self.buf((g, g, g), mid), self.buf((g, g, g), mid)
      st.update(z=z, c=c, z2=z2)
      st["bn1"] = BRNOp(self, f"decoder.stage_{stage}.b1", cat, cin, True, z)
      st["bn2"] = BRNOp(self, f"decoder.stage_{stage}.b2", c, mid, True, z2)
      st["lc"], st["lt"] = L[f"stage_{stage}.c1"], L[f"stage_{stage}.t1"]
      st["d_c"] = st["lc"].desc(z.cs, (g, g, g), c.cs, (g, g, g), B)
      g2 = 2 * g
      if stage < 6:
        nxt = self.buf((g2, g2, g2), t_out + skip_c)
        st["next"] = nxt
        st["d_t"] = st["lt"].desc(z2.cs, (g, g, g), nxt.cs, (g2, g2, g2), B)
        if skip_c:
          ls = L[f"rt_skip_{stage}"]
          src = self.enc_pre[ENC_STAGE_OF[enc_c]]
          hw = src.spatial[0]
          cmap = self.buf((hw, hw), skip_c)
          st.update(ls=ls, src=src, cmap=cmap, hw=hw)
          st["d_s"] = ls.desc(src.cs, (1, hw, hw), cmap.cs, (1, hw, hw), B, bias_n_stride=skip_c)
          res = eng.model.config.decoder.resolution
          st["scale"] = t.diag(t.tensor([res[0] / g2, res[1] / g2, res[2] / g2, 1.0], dtype=t.float32)).to(self.dev)
          st["sbias"] = t.zeros(B, skip_c, dtype=t.float32, device=self.dev)
          st["mat"] = t.zeros(B, 4, 4, dtype=t.float32, device=self.dev)
          st["ssum"] = t.zeros(B, skip_c, dtype=t.float32, device=self.dev)
          st["dwoff"] = t.zeros(skip_c, 3, dtype=t.float32, device=self.dev)
          if self.need_grad:
            nb = _lib.lib().crn_skip_lists_workspace_bytes(B, hw, hw, g2, g2, g2)
            st["lists_ws"] = t.empty(nb, dtype=t.uint8, device=self.dev)
            st["sorted_vox"] = t.empty(B * g2 ** 3, dtype=t.int32, device=self.dev)
            st["starts"] = t.empty(B * hw * hw + 1, dtype=t.int32, device=self.dev)
        cat = nxt
      else:
        cp = _r4(t_out)
        st["glog"] = t.zeros(B * g2 ** 3, cp, dtype=t.float32, device=self.dev) if self.need_grad else None
        st["d_t"] = st["lt"].desc(z2.cs, (g, g, g), cp, (g2, g2, g2), B, planar=True)
        st["d_t_bwd"] = st["lt"].desc(z2.cs, (g, g, g), cp, (g2, g2, g2), B, planar=False)
        st["rows"] = eng.rows_pad.get(st["lt"].name)
        if st["rows"] is not None:
          d16 = st["lt"].desc(z2.cs, (g, g, g), cp, (g2, g2, g2), B, planar=False)
          d16.Cout = d16.CoutP = cp
          st["d_t_rows"] = d16
          st["logits_rows"] = t.zeros(B * g2 ** 3, cp, dtype=t.float32, device=self.dev)
          self.rows_cp = cp
      self.stages.append(st)
    self.scratch64 = t.zeros(4096, dtype=t.float64, device=self.dev)
    self.nbt_snapshot = t.zeros(max(len(self.fused_brn), 1), dtype=t.int64, device=self.dev)
    self.n_fused_enc = sum(1 for op in self.fused_brn if op.name.startswith("encoder."))
    self.offs = t.zeros(B, 3, dtype=t.float32, device=self.dev)
    res = eng.model.config.decoder.resolution
    self.logits = t.zeros(B, eng.model.config.decoder.num_output_channels, *res, dtype=t.float32, device=self.dev)

  # ------------------------------------------------------------------ forward
  @t.no_grad()
  def forward(self, image: t.Tensor, v2s: t.Tensor, offsets: t.Tensor, training: bool,
              want_features: bool = False, pack: bool = True, run_encoder: bool = True,
              rows_logits: bool = False) -> t.Tensor:
    """Enqueues the forward pass on the current stream and returns the plan-owned logits buffer
    float32[B,C,D,H,W] (overwritten by the next forward of this plan).  pack=False: the caller has already
    re-packed the weights (graph replays pack outside the captured region when the weights did not change);
    run_encoder=False: reuse the encoder activations of the previous forward of this plan (same image; the
    encoder does not depend on the sample offsets -- super_resolution.py:92-126)."""
    global _ENG
    eng, B = self.eng, self.B
    _ENG = eng
    P, _ = eng.tensors()
    st = _lib.stream_ptr()
    self.training = training
    bias = lambda l: P[l.name + ".bias"].data_ptr()
    if pack:
      eng.pack_weights()
    if run_encoder:
      if training:
        self.arena_fwd.zero_()
      self._forward_encoder(image, training, P, bias, st)
    else:
      assert not training, "run_encoder=False is an inference-only path"
      eng.join_packs()
    if want_features:
      eng.join_packs()
      return None
    out = self._forward_decoder(v2s, offsets, training, P, bias, st, rows_logits)
    eng.join_packs()                 # the dgrad re-packs ran beside the forward pass
    return out

  def _forward_encoder(self, image, training, P, bias, st):
    eng, B = self.eng, self.B
    # ---- encoder
    if training:
      self._snapshot_nbt(0, self.n_fused_enc)
    _call("crn_preprocess_image", image.data_ptr(), B, 256, 256, self.img4.p, st)
    conv_call("fwd", eng.L["stem"], self.d_stem, self.img4.p, eng.wf(eng.L["stem"]), bias(eng.L["stem"]),
              self.s1.p, 0, st)
    self.brn_stem.fwd(training)
    _call("crn_maxpool_fwd", self.s1a.p, B, 128, 128, 64, self.p1.p, self.p1_idx.data_ptr(), st)
    eng.join_packs(fwd_only=True)
    for blk in self.blocks:
      x = blk["x"]
      if blk["down"]:
        conv_call("fwd", blk["ls"], blk["d_s"], x.p, eng.wf(blk["ls"]), bias(blk["ls"]), blk["s_c"].p, 0, st)
        blk["bn_s"].fwd(training)
      conv_call("fwd", blk["la"], blk["d_a"], x.p, eng.wf(blk["la"]), bias(blk["la"]), blk["a_c"].p, 0, st)
      blk["bn_a"].fwd(training)
      conv_call("fwd", blk["lb"], blk["d_b"], blk["a_y"].p, eng.wf(blk["lb"]), bias(blk["lb"]),
                blk["b_c"].p, 0, st)
      blk["bn_b"].fwd(training)
      conv_call("fwd", blk["lc"], blk["d_c"], blk["b_y"].p, eng.wf(blk["lc"]), bias(blk["lc"]),
                blk["c_c"].p, 0, st)
      blk["bn_c"].fwd(training)
    _call("crn_spatial_mean_fwd", self.enc_out.p, B, 64, 2048, self.feat.p, st)

  def _forward_decoder(self, v2s, offsets, training, P, bias, st, rows_logits=False):
    eng, B = self.eng, self.B
    # ---- decoder
    if training:
      self._snapshot_nbt(self.n_fused_enc, len(self.fused_brn))
    L = eng.L
    lat = eng.lat
    conv_call("fwd", L["stage_0"], self.d_s0, self.feat.p, eng.wf(L["stage_0"]), bias(L["stage_0"]),
              self.latb.p, 0, st)
    self.latb.v[:, lat:lat + 3].copy_(offsets)
    self.brn_s1.fwd(training)
    conv_call("fwd", L["stage_1.t1"], self.d_s1t, self.z1.p, eng.wf(L["stage_1.t1"]), bias(L["stage_1.t1"]),
              self.x1.p, 0, st)
    # per-scale layer matrices (reconstruction_decoder.py:111-115) and offsets
    self.offs.copy_(offsets)
    # per-scale layer matrices v2s @ scale(res / g) (reconstruction_decoder.py:111-115)
    skips = [sd for sd in self.stages if sd["stage"] < 6 and sd["skip_c"]]
    for sd in skips:
      t.matmul(v2s, sd["scale"], out=sd["mat"])
    if self.need_grad and DETERMINISTIC_SKIP_BWD:
      # voxel lists per pixel for the atomics-free skip backward: they depend on (v2s, offsets) only and are built on
      # the side stream while the decoder runs
      main, side = t.cuda.current_stream(), self.side_stream
      ev = t.cuda.Event()
      ev.record(main)
      side.wait_event(ev)
      with t.cuda.stream(side):
        for sd in skips:
          g2, hw = 2 * sd["g"], sd["hw"]
          _call("crn_skip_build_lists", B, hw, hw, sd["mat"].data_ptr(), self.offs.data_ptr(), g2, g2, g2,
                sd["lists_ws"].data_ptr(), sd["lists_ws"].numel(), sd["sorted_vox"].data_ptr(),
                sd["starts"].data_ptr(), side.cuda_stream, hbm=(f"skip_lists g={g2}", 40 * B * g2 ** 3))
        lists_ev = t.cuda.Event()
        lists_ev.record(side)
    logits = None
    for sd in self.stages:
      g = sd["g"]
      sd["bn1"].fwd(training)
      if USE_TC and sd["lc"].name in eng.tcs_w:
        conv5_tcs_call(sd["lc"], sd["d_c"], sd["z"].p, eng.tcs_w[sd["lc"].name].data_ptr(), bias(sd["lc"]), sd["c"].p,
                       eng.tc_status.data_ptr(), st)
      elif USE_TC and sd["lc"].name in eng.tc_w:
        conv5_tc_call("fwd", sd["lc"], sd["d_c"], sd["z"].p, eng.tc_w[sd["lc"].name][0].data_ptr(), bias(sd["lc"]),
                      sd["c"].p, eng.tc_status.data_ptr(), st)
      elif USE_TC and sd["lc"].name in eng.tc_wf:
        conv5_tc_call("fwd", sd["lc"], sd["d_c"], sd["z"].p, eng.tc_wf[sd["lc"].name].data_ptr(), bias(sd["lc"]),
                      sd["c"].p, eng.tc_status.data_ptr(), st)
      else:
        conv_call("fwd", sd["lc"], sd["d_c"], sd["z"].p, eng.wf(sd["lc"]), bias(sd["lc"]), sd["c"].p, 0, st)
      sd["bn2"].fwd(training)
      if sd["stage"] < 6:
        nxt = sd["next"]
        if USE_TC and eng.tct_w.get(sd["lt"].name, (None, None))[0] is not None:
          convt7_tc_call(sd["lt"], sd["d_t"], sd["z2"].p, eng.tct_w[sd["lt"].name][0].data_ptr(), bias(sd["lt"]), nxt.p,
                         eng.tc_status.data_ptr(), st)
        elif USE_TC and sd["lt"].name in eng.tct_slices:
          if "d_t_slices" not in sd:
            sd["d_t_slices"] = []
            for _, _, co0 in eng.tct_slices[sd["lt"].name]:
              ds = ConvDesc.from_buffer_copy(bytes(sd["d_t"]))
              ds.Cout, ds.y_co = 16, sd["d_t"].y_co + co0
              sd["d_t_slices"].append(ds)
          for (wt, _, co0), ds in zip(eng.tct_slices[sd["lt"].name], sd["d_t_slices"]):
            convt7_tc_call(sd["lt"], ds, sd["z2"].p, wt.data_ptr(), bias(sd["lt"]) + 4 * co0, nxt.p,
                           eng.tc_status.data_ptr(), st)
        else:
          conv_call("fwd", sd["lt"], sd["d_t"], sd["z2"].p, eng.wf(sd["lt"]), bias(sd["lt"]), nxt.p, 0, st)
        if sd["skip_c"]:
          ls, src, cmap, hw = sd["ls"], sd["src"], sd["cmap"], sd["hw"]
          w = P[ls.name + ".weight"]
          # the 3 offset channels are constant per scene -> exact per-scene bias
          sbias, mat = sd["sbias"], sd["mat"]
          t.addmm(P[ls.name + ".bias"], self.offs, w[:, ls.cin:, 0, 0].t(), out=sbias)
          conv_call("fwd", ls, sd["d_s"], src.p, eng.wf(ls), sbias.data_ptr(), cmap.p, 0, st)
          g2 = 2 * g
          _call("crn_skip_sample_fwd", cmap.p, B, hw, hw, sd["skip_c"], cmap.cs, mat.data_ptr(),
                self.offs.data_ptr(), g2, g2, g2, nxt.p, nxt.cs, sd["t_out"], st)
      elif sd["rows"] is not None:
        rp, rows = sd["rows"], sd["logits_rows"]
        convt7_tc_call(sd["lt"], sd["d_t_rows"], sd["z2"].p, rp["fwd"].data_ptr(), rp["b"].data_ptr(), rows.data_ptr(),
                       eng.tc_status.data_ptr(), st, acct=sd["d_t_bwd"])
        _call("crn_status_poison", eng.tc_status.data_ptr(), rows.data_ptr(), rows.numel() // B, B, st)
        if rows_logits:
          logits = rows
        else:
          logits = self.logits
          _call("crn_rows_to_planar", rows.data_ptr(), B, sd["t_out"], (2 * g) ** 3, rp["cp"], logits.data_ptr(), st)
      else:
        logits = self.logits
        if sd["lt"].name in eng.tctsf_w:
          convt7_tc_call(sd["lt"], sd["d_t"], sd["z2"].p, eng.tctsf_w[sd["lt"].name].data_ptr(), bias(sd["lt"]),
                         logits.data_ptr(), eng.tc_status.data_ptr(), st, fn="crn_convt7_tcs_fwd")
        elif USE_TC and eng.tct_w.get(sd["lt"].name, (None, None))[0] is not None:
          convt7_tc_call(sd["lt"], sd["d_t"], sd["z2"].p, eng.tct_w[sd["lt"].name][0].data_ptr(), bias(sd["lt"]),
                         logits.data_ptr(), eng.tc_status.data_ptr(), st)
        else:
          conv_call("fwd", sd["lt"], sd["d_t"], sd["z2"].p, eng.wf(sd["lt"]), bias(sd["lt"]),
                    logits.data_ptr(), 0, st)
    if self.need_grad and DETERMINISTIC_SKIP_BWD:
      t.cuda.current_stream().wait_event(lists_ev)
    # a tcgen05 barrier timeout (status word != 0) must not go unnoticed: NaN in every scene's logits
    if logits is self.logits:
      _call("crn_status_poison", eng.tc_status.data_ptr(), logits.data_ptr(), logits[0].numel(), B, st)
    return logits

  # ------------------------------------------------------------------ backward
  @t.no_grad()
  def backward(self, grad_logits: t.Tensor, grads: Dict[str, t.Tensor], chunk_cb=None, comm_stream=None,
               grad_is_rows: bool = False):
    """grads: name -> zero/empty tensor per parameter (filled here).  grad_is_rows: the caller has written the
    logit gradient as channels-last rows straight into this plan's `glog_rows()` buffer (rows-mode plans only).

    chunk_cb(i): data-parallel hook.  The parameters are finished in three chunks in backward order -- GRAD_CHUNKS:
    0 = decoder, 1 = encoder stage4+5, 2 = the rest -- and as soon as every gradient of chunk i is final its
    un-pack / bias-gather launches and chunk_cb(i) (the all-reduce + optimiser update of that slice of the flat
    buffer) are enqueued on comm_stream, overlapping the rest of the backward pass (the reference: DDP's bucketed
    all-reduce, pipeline.py:199-200,229).  Without a callback everything is finished by one launch set at the end."""
    global _ENG
    eng, B = self.eng, self.B
    _ENG = eng
    P, _ = eng.tensors()
    st = _lib.stream_ptr()
    L = eng.L
    tr = self.training
    self.arena_bwd.zero_()
    eng.dw.zero_()

    # bias gradients = fp64 column sums in accumulator slots: collected here, moved by ONE batched launch at the end
    bias_jobs = []
    mark = [0]

    def bias_from(slot: Slot, name, lo=0, n=None):
      n = n if n is not None else grads[name].numel()
      bias_jobs.append((slot.data_ptr() + 8 * lo, grads[name].data_ptr(), n))

    def finish(ci, final=False):
      """Bias gather + weight-gradient un-pack (+ the skip convs' offset columns) of chunk ci; see the docstring."""
      if chunk_cb is None and not final:
        return
      jobs = tuple(bias_jobs[mark[0]:])
      mark[0] = len(bias_jobs)
      layers = eng.layers if chunk_cb is None else eng.chunk_layers(ci)
      where = main if chunk_cb is None else comm_stream
      if chunk_cb is not None:
        where.wait_stream(main)
      if side is not None:
        where.wait_stream(side)
      with t.cuda.stream(where):
        if jobs:
          self._gather_bias(ci if chunk_cb is not None else -1, jobs)
        eng.unpack_wgrads(grads, layers, ci if chunk_cb is not None else -1)
        if chunk_cb is None or ci == 0:
          # offset-channel columns of the skip compress convs (the GEMM ran without them)
          for sd in self.stages:
            if sd["stage"] < 6 and sd["skip_c"]:
              ls = sd["ls"]
              t.matmul(sd["ssum"].t(), self.offs, out=sd["dwoff"])
              grads[ls.name + ".weight"][:, ls.cin:, 0, 0].copy_(sd["dwoff"])
        if chunk_cb is not None:
          chunk_cb(ci)
      if chunk_cb is not None and final:
        main.wait_stream(comm_stream)

    # Weight gradients only feed the final un-pack, so they run on a side stream concurrently with the dgrad /
    # BatchRenorm chain (fork: event after dy is final; join: before crn_unpack_wgrads).  The small encoder launches
    # (16-160 CTAs) no longer serialise behind each other; inside a captured graph this becomes a parallel branch.
    main = t.cuda.current_stream()
    side = self.side_stream if WGRAD_SIDE_STREAM else None

    def wgrad(l, d, x_ptr, dy_ptr):
      if side is None:
        conv_call("wgrad", l, d, x_ptr, dy_ptr, eng.dwp(l), st)
        return
      ev = t.cuda.Event()
      ev.record(main)
      side.wait_event(ev)
      with t.cuda.stream(side):
        conv_call("wgrad", l, d, x_ptr, dy_ptr, eng.dwp(l), side.cuda_stream)

    def dgrad(l, d, dy_ptr, dx_ptr, acc=0):
      conv_call("dgrad", l, d, dy_ptr, eng.wd(l), dx_ptr, acc, st)

    # ---- decoder, last stage first
    grad_logits = grad_logits.contiguous()
    for sd in reversed(self.stages):
      g, stage = sd["g"], sd["stage"]
      lt, lc = sd["lt"], sd["lc"]
      if stage == 6:
        g2 = 2 * g
        S = g2 ** 3
        cp = sd["glog"].shape[1]
        if grad_is_rows:
          assert sd["rows"] is not None and grad_logits.data_ptr() == sd["glog"].data_ptr()
          _call("crn_colsum", sd["glog"].data_ptr(), B * S, sd["t_out"], cp, 0, grads[lt.name + ".bias"].data_ptr(),
                self.scratch64.data_ptr(), st)
        else:
          _call("crn_planar_to_rows", grad_logits.data_ptr(), B, sd["t_out"], S, cp, sd["glog"].data_ptr(), st)
          _call("crn_colsum_planar", grad_logits.data_ptr(), B, sd["t_out"], S,
                grads[lt.name + ".bias"].data_ptr(), self.scratch64.data_ptr(), st)
        dy_ptr, d_t = sd["glog"].data_ptr(), sd["d_t_bwd"]
      else:
        nxt = sd["next"]
        dy_ptr, d_t = nxt.gp, sd["d_t"]
        if sd["skip_c"]:
          ls, src, cmap, hw = sd["ls"], sd["src"], sd["cmap"], sd["hw"]
          g2 = 2 * g
          if DETERMINISTIC_SKIP_BWD:
            _call("crn_skip_sample_bwd_sorted", nxt.gp, nxt.cs, sd["t_out"], B, hw, hw, sd["skip_c"], cmap.cs,
                  sd["sorted_vox"].data_ptr(), sd["starts"].data_ptr(), cmap.gp, st,
                  hbm=(f"skip_bwd g={g2}", 4 * sd["skip_c"] * B * (g2 ** 3 + hw * hw) + 4 * B * g2 ** 3))
          else:
            cmap.g.zero_()
            _call("crn_skip_sample_bwd", nxt.gp, nxt.cs, sd["t_out"], B, hw, hw, sd["skip_c"], cmap.cs,
                  sd["mat"].data_ptr(), self.offs.data_ptr(), g2, g2, g2, cmap.gp, st)
          wgrad(ls, sd["d_s"], src.p, cmap.gp)
          dgrad(ls, sd["d_s"], cmap.gp, src.gp, 0)
          # per-scene sums of d(cmap) -> bias and offset-channel weight grads (tiny glue)
          ssum = sd["ssum"]
          _call("crn_spatial_mean_fwd", cmap.gp, B, hw * hw, cmap.cs, ssum.data_ptr(), st)
          ssum.mul_(float(hw * hw))
      wgrad(lt, d_t, sd["z2"].p, dy_ptr)
      if lt.name in eng.tcts_wd:
        d_dg, dy_dg = (sd["d_t_rows"] if (stage == 6 and sd["rows"] is not None) else d_t), dy_ptr
        if stage == 6 and sd["rows"] is None:
          d_dg, dy_dg = sd["d_t"], grad_logits.data_ptr()      # planar gradient of the two FG_BG logits
        convt7_tc_dgrad_call(lt, d_dg, dy_dg, eng.tcts_wd[lt.name][0].data_ptr(), sd["z2"].gp,
                             eng.tc_status.data_ptr(), st, acct=d_t, fn="crn_convt7_tcs_dgrad")
      elif stage == 6 and sd["rows"] is not None:
        convt7_tc_dgrad_call(lt, sd["d_t_rows"], dy_ptr, sd["rows"]["dgrad"].data_ptr(), sd["z2"].gp,
                             eng.tc_status.data_ptr(), st, acct=d_t)
      elif USE_TC and eng.tct_w.get(lt.name, (None, None))[1] is not None:
        convt7_tc_dgrad_call(lt, d_t, dy_ptr, eng.tct_w[lt.name][1].data_ptr(), sd["z2"].gp, eng.tc_status.data_ptr(), st)
      else:
        dgrad(lt, d_t, dy_ptr, sd["z2"].gp, 0)
      dxs = sd["bn2"].bwd(tr, grads, sd["z2"].gp, sd["z2"].cs, sd["c"].gp, sd["c"].cs)
      bias_from(dxs, lc.name + ".bias")
      wgrad(lc, sd["d_c"], sd["z"].p, sd["c"].gp)
      if USE_TC and lc.name in eng.tcs_wd:
        conv5_tcs_call(lc, sd["d_c"], sd["c"].gp, eng.tcs_wd[lc.name].data_ptr(), None, sd["z"].gp,
                       eng.tc_status.data_ptr(), st, kind=1)
      elif USE_TC and lc.name in eng.tcs_wd_slices:
        if "d_c_slices" not in sd:
          sd["d_c_slices"] = []
          for buf, wslice, c0, h in eng.tcs_wd_slices[lc.name]:
            dsl = type(sd["d_c"]).from_buffer_copy(sd["d_c"])
            dsl.Cin = dsl.CinP = h
            dsl.x_co = sd["d_c"].x_co + c0
            sd["d_c_slices"].append(dsl)
        for (buf, wslice, c0, h), dsl in zip(eng.tcs_wd_slices[lc.name], sd["d_c_slices"]):
          conv5_tcs_call(lc, dsl, sd["c"].gp, buf.data_ptr(), None, sd["z"].gp, eng.tc_status.data_ptr(), st, kind=1)
      elif USE_TC and lc.name in eng.tc_w:
        conv5_tc_call("dgrad", lc, sd["d_c"], sd["c"].gp, eng.tc_w[lc.name][1].data_ptr(), None, sd["z"].gp,
                      eng.tc_status.data_ptr(), st)
      else:
        dgrad(lc, sd["d_c"], sd["c"].gp, sd["z"].gp, 0)
      cat = sd["cat"]
      dxs = sd["bn1"].bwd(tr, grads, sd["z"].gp, sd["z"].cs, cat.gp, cat.cs)
      sd["dxs_cat"] = dxs
    # bias grads of the transposed convs = column sums of the next stage's cat gradient
    for i, sd in enumerate(self.stages):
      if sd["stage"] < 6:
        nxt_sd = self.stages[i + 1]
        bias_from(nxt_sd["dxs_cat"], sd["lt"].name + ".bias", 0, sd["t_out"])
        if sd["skip_c"]:
          ls = sd["ls"]
          t.sum(sd["ssum"], 0, out=grads[ls.name + ".bias"])
    bias_from(self.stages[0]["dxs_cat"], L["stage_1.t1"].name + ".bias")
    # ---- stage_1 / stage_0
    l1 = L["stage_1.t1"]
    wgrad(l1, self.d_s1t, self.z1.p, self.x1.gp)
    dgrad(l1, self.d_s1t, self.x1.gp, self.z1.gp, 0)
    dxs = self.brn_s1.bwd(tr, grads, self.z1.gp, self.z1.cs, self.latb.gp, self.latb.cs)
    l0 = L["stage_0"]
    bias_from(dxs, l0.name + ".bias", 0, eng.lat)
    wgrad(l0, self.d_s0, self.feat.p, self.latb.gp)
    dgrad(l0, self.d_s0, self.latb.gp, self.feat.gp, 0)
    finish(0)
    _call("crn_spatial_mean_bwd", self.feat.gp, B, 64, 2048, self.enc_out.gp, 0, st)
    # ---- encoder, last block first
    for blk in reversed(self.blocks):
      x = blk["x"]
      out, pre = blk["out"], blk["pre"]
      la, lb, lc = blk["la"], blk["lb"], blk["lc"]
      res_g = blk["s_y"].gp if blk["down"] else x.gp
      # masked (and skip-combined) gradient doubles as the residual branch's gradient
      dxs = blk["bn_c"].bwd(tr, grads, out.gp, out.cs, blk["c_c"].gp, blk["c_c"].cs,
                            g_extra_ptr=pre.gp if pre is not None else None, g_store_ptr=res_g)
      bias_from(dxs, lc.name + ".bias")
      wgrad(lc, blk["d_c"], blk["b_y"].p, blk["c_c"].gp)
      dgrad(lc, blk["d_c"], blk["c_c"].gp, blk["b_y"].gp, 0)
      dxs = blk["bn_b"].bwd(tr, grads, blk["b_y"].gp, blk["b_y"].cs, blk["b_c"].gp, blk["b_c"].cs)
      bias_from(dxs, lb.name + ".bias")
      wgrad(lb, blk["d_b"], blk["a_y"].p, blk["b_c"].gp)
      dgrad(lb, blk["d_b"], blk["b_c"].gp, blk["a_y"].gp, 0)
      dxs = blk["bn_a"].bwd(tr, grads, blk["a_y"].gp, blk["a_y"].cs, blk["a_c"].gp, blk["a_c"].cs)
      bias_from(dxs, la.name + ".bias")
      wgrad(la, blk["d_a"], x.p, blk["a_c"].gp)
      if blk["down"]:
        ls = blk["ls"]
        dxs = blk["bn_s"].bwd(tr, grads, blk["s_y"].gp, blk["s_y"].cs, blk["s_c"].gp, blk["s_c"].cs)
        bias_from(dxs, ls.name + ".bias")
        wgrad(ls, blk["d_s"], x.p, blk["s_c"].gp)
        dgrad(ls, blk["d_s"], blk["s_c"].gp, x.gp, 0)
        dgrad(la, blk["d_a"], blk["a_c"].gp, x.gp, 1)
      else:
        dgrad(la, blk["d_a"], blk["a_c"].gp, x.gp, 1)
      if blk["p"] == "encoder.stage4.a.":
        finish(1)
    # ---- stem
    _call("crn_maxpool_bwd", self.p1.gp, self.p1_idx.data_ptr(), B, 128, 128, 64, self.s1a.gp, st)
    dxs = self.brn_stem.bwd(tr, grads, self.s1a.gp, self.s1a.cs, self.s1.gp, self.s1.cs)
    stem = L["stem"]
    bias_from(dxs, stem.name + ".bias")
    wgrad(stem, self.d_stem, self.img4.p, self.s1.gp)
    finish(2, final=True)

  def _gather_bias(self, key, jobs):
    """ONE launch moves the fp64 column sums of `jobs` into the bias gradients (item list cached per pointer set)."""
    eng = self.eng
    cache = self.__dict__.setdefault("_bias_cache", {})
    ent = cache.get(key)
    if ent is None or ent[0] != jobs:
      items = (_lib.F64CopyItem * len(jobs))()
      offs = (C.c_int64 * (len(jobs) + 1))()
      tot = 0
      for i, (src, dst, n) in enumerate(jobs):
        items[i].src, items[i].dst = src, dst
        offs[i] = tot
        tot += n
      offs[len(jobs)] = tot
      ent = cache[key] = (jobs, eng._to_dev(items, self.dev), eng._to_dev(offs, self.dev), tot)
    _call("crn_gather_f64_to_f32", ent[1].data_ptr(), ent[2].data_ptr(), len(jobs), ent[3], _lib.stream_ptr())

  def glog_rows(self) -> t.Tensor:
    """The rows-layout logit-gradient buffer float32[B*128^3, rows_cp] the backward pass reads (rows-mode plans)."""
    return self.stages[-1]["glog"]

  # ------------------------------------------------------------------ graph-replayed entry points (module path)
  def _sig(self):
    """Everything a captured graph bakes in: parameter / buffer addresses and the kernel routing switches."""
    P, Bf = self.eng.tensors()
    return (tuple(p.data_ptr() for p in P.values()), tuple(b.data_ptr() for b in Bf.values()), USE_TC, USE_TC5S,
            WGRAD_SIDE_STREAM, PRECISION)

  def _graph_state(self, key):
    sig = self._sig()
    gs = self._graphs.get(key)
    if gs is None or gs["sig"] != sig:
      gs = self._graphs[key] = {"sig": sig, "calls": 0, "graph": None}
    return gs

  def run_forward(self, image, v2s, offsets, training, run_encoder=True):
    """forward() through a CUDA graph: inputs are copied into the plan's static buffers; the first GRAPH_WARMUP calls
    per (mode, parameter set) run eagerly (lazy initialisation must not happen under capture), the next one captures,
    later ones replay.  The weight re-pack stays outside the graph: it is skipped when the weights did not change."""
    eng = self.eng
    if not graphs_enabled():
      return self.forward(image, v2s, offsets, training, run_encoder=run_encoder)
    self.training = training
    if run_encoder:
      self.in_image.copy_(image)
    self.in_v2s.copy_(v2s)
    self.in_offs.copy_(offsets)
    args = (self.in_image, self.in_v2s, self.in_offs, training)
    gs = self._graph_state(("fwd", training, run_encoder))
    if gs["graph"] is None:
      if gs["calls"] < GRAPH_WARMUP:
        gs["calls"] += 1
        return self.forward(*args, run_encoder=run_encoder)
      eng.pack_weights()
      eng.join_packs()
      g = t.cuda.CUDAGraph()
      with t.cuda.graph(g, capture_error_mode="thread_local"):
        self.forward(*args, pack=False, run_encoder=run_encoder)
      gs["graph"] = g
    eng.pack_weights()
    eng.join_packs()
    gs["graph"].replay()
    return self.logits

  def run_backward(self, grad_logits, names):
    """backward() through a CUDA graph; returns name -> gradient (fresh tensors: views of one flat copy)."""
    eng = self.eng
    P, _ = eng.tensors()
    if self.gflat is None or self._gnames != names:
      sizes = [P[n].numel() for n in names]
      self.gflat = t.zeros(sum(sizes), dtype=t.float32, device=self.dev)
      self._gnames, self._gviews, off = list(names), {}, 0
      for n, sz in zip(names, sizes):
        self._gviews[n] = (off, sz, P[n].shape)
        off += sz
      self.gstatic = {n: self.gflat[o:o + sz].view(shp) for n, (o, sz, shp) in self._gviews.items()}
      self.in_glog = t.zeros_like(self.logits)
      self._graphs = {k: v for k, v in self._graphs.items() if k[0] != "bwd"}

    def enqueue(glog):
      self.gflat.zero_()
      self.backward(glog, self.gstatic)

    if not graphs_enabled():
      enqueue(grad_logits)
    else:
      self.in_glog.copy_(grad_logits)
      gs = self._graph_state(("bwd", self.training))
      if gs["graph"] is None and gs["calls"] < GRAPH_WARMUP:
        gs["calls"] += 1
        enqueue(self.in_glog)
      else:
        if gs["graph"] is None:
          g = t.cuda.CUDAGraph()
          with t.cuda.graph(g, capture_error_mode="thread_local"):
            enqueue(self.in_glog)
          gs["graph"] = g
        gs["graph"].replay()
    out = self.gflat.clone()
    return {n: out[o:o + sz].view(shp) for n, (o, sz, shp) in self._gviews.items()}

  # ------------------------------------------------------------------ features (NCHW copies)
  def features_nchw(self):
    B = self.B
    def nchw(b: Buf):
      hw = b.spatial[0]
      return b.v.view(B, hw, hw, b.cs)[..., :b.C].permute(0, 3, 1, 2).contiguous()
    return (nchw(self.s1), nchw(self.enc_pre["stage2"]), nchw(self.enc_pre["stage3"]),
            nchw(self.enc_pre["stage4"]), nchw(self.enc_pre["stage5"]), self.feat.v[:, :2048].clone())


# ====================================================================== autograd bridge
class _Lease:
  """Marks a plan busy between a grad-enabled forward and its backward.  If the autograd node dies without a
  backward (an eval loop without no_grad, an exception) the lease is released by its destructor, so the plan is
  reused instead of a new ~0.5 GB plan being built on every such call."""

  def __init__(self, plan):
    plan.busy = True
    plan.gen += 1
    self.plan, self.gen = plan, plan.gen

  def release(self):
    plan, self.plan = self.plan, None
    if plan is not None and plan.gen == self.gen:
      plan.busy = False

  __del__ = release


class _CoreNetFn(t.autograd.Function):
  @staticmethod
  def forward(ctx, model, need_grad, image, v2s, offsets, *params):
    eng = get_engine(model)
    plan = eng.get_plan(image.shape[0], image.device, need_grad)
    logits = plan.run_forward(image, v2s, offsets, model.training)
    if need_grad:
      ctx.lease = _Lease(plan)
      ctx.model = model
    ctx.nparams = len(params)
    # the plan owns its logits buffer (the next forward overwrites it): hand out a copy
    return logits.clone()

  @staticmethod
  def backward(ctx, grad_logits):
    lease, model = ctx.lease, ctx.model
    plan = lease.plan
    if plan is None:
      raise RuntimeError("corenet_b200: backward through the same CoreNet forward twice is not supported")
    names = [n for n, _ in model.named_parameters()]
    try:
      grads = plan.run_backward(grad_logits, names)
    finally:
      lease.release()
    return (None, None, None, None, None) + tuple(grads[n] for n in names)


def get_engine(model) -> Engine:
  eng = model.__dict__.get("_crn_engine")
  if eng is None:
    eng = Engine(model)
    model.__dict__["_crn_engine"] = eng
  return eng


def _check_inputs(image, v2s, offsets):
  if not image.is_cuda:
    raise RuntimeError("corenet_b200.CoreNet runs on CUDA only (no CPU fallback); move the module and "
                       "its inputs to a B200")
  if image.dtype != t.uint8 or image.dim() != 4 or image.shape[1] != 3:
    raise AssertionError("image must be uint8[B,3,H,W]")
  if tuple(image.shape[2:]) != (256, 256):
    raise ValueError("the encoder feature pyramid (256 -> 8) is fixed: image must be 256x256")


def corenet_forward(model, image: t.Tensor, v2s: t.Tensor, offsets: t.Tensor) -> t.Tensor:
  _check_inputs(image, v2s, offsets)
  b = image.shape[0]
  v2s = v2s.to(dtype=t.float32)
  offsets = offsets.to(dtype=t.float32)
  assert v2s.shape == (b, 4, 4) and offsets.shape == (b, 3)
  params = list(model.parameters())
  need_grad = t.is_grad_enabled() and any(p.requires_grad for p in params)
  return _CoreNetFn.apply(model, need_grad, image.contiguous(), v2s.contiguous(), offsets.contiguous(),
                          *params)


def corenet_multi_offset_pmf(model, image: t.Tensor, v2s: t.Tensor, grid_offsets: t.Tensor) -> t.Tensor:
  """Class probabilities for several sample offsets of the same images: float32[n_off, B, 3] offsets ->
  float32[n_off, B, C, D, H, W].  The encoder does not depend on the offsets (super_resolution.py:92-126 of the
  reference re-runs it mult^3 times): in eval mode it runs ONCE and only the decoder + skip connections + softmax
  are replayed per offset.  In train mode BatchRenorm statistics are batch statistics, so the reference's loop is
  kept as is."""
  from corenet_b200 import ops
  _check_inputs(image, v2s, grid_offsets)
  b, n_off = image.shape[0], grid_offsets.shape[0]
  v2s = v2s.to(dtype=t.float32).contiguous()
  grid_offsets = grid_offsets.to(dtype=t.float32)
  assert v2s.shape == (b, 4, 4) and grid_offsets.shape[1:] == (b, 3)
  with t.no_grad():
    if model.training:
      return t.stack([ops.softmax_channels(model(image, v2s, o)) for o in grid_offsets], 0)
    eng = get_engine(model)
    plan = eng.get_plan(b, image.device, False)
    c = model.config.decoder.num_output_channels
    out = t.empty((n_off, b, c) + tuple(model.config.decoder.resolution), dtype=t.float32, device=image.device)
    for i in range(n_off):
      logits = plan.run_forward(image.contiguous(), v2s, grid_offsets[i], False, run_encoder=(i == 0))
      _call("crn_softmax_planar", logits.data_ptr(), b, c, logits[0, 0].numel(), out[i].data_ptr(),
            _lib.stream_ptr())
    return out


def encoder_forward(encoder, image_f32: t.Tensor):
  raise NotImplementedError(
      "standalone encoder forward: use CoreNet.forward / CoreNet.encode_features (the engine fuses the "
      "Caffe preprocessing into its staging kernel and takes the uint8 image)")
