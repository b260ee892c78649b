"""Tensor-level wrappers of the C-ABI ops (the standalone operator surface).

Each function takes/returns torch CUDA tensors in the REFERENCE's layouts
(NCHW / NCDHW, planar logits), converts to the engine's channels-last rows where
needed, and launches the kernels on the current stream.  They raise on CPU
tensors: there is no CPU fallback.
"""
import ctypes as C
from typing import Optional, Tuple

import torch as t

from corenet_b200 import _lib
from corenet_b200._lib import ConvDesc

_call = _lib.call


def _r4(c):
  return (c + 3) // 4 * 4


def _need_cuda(*xs):
  for x in xs:
    if x is not None and not x.is_cuda:
      raise ValueError("corenet_b200 ops need CUDA tensors (there is no CPU fallback)")


def _to_rows(x: t.Tensor) -> Tuple[t.Tensor, int, int]:
  """N C ... -> (rows x C contiguous, rows, C)."""
  c = x.shape[1]
  perm = [0] + list(range(2, x.dim())) + [1]
  r = x.permute(perm).contiguous().view(-1, c)
  return r, r.shape[0], c


def _from_rows(r: t.Tensor, shape) -> t.Tensor:
  n, c = shape[0], shape[1]
  sp = list(shape[2:])
  inv = [0, len(sp) + 1] + list(range(1, len(sp) + 1))
  return r.view([n] + sp + [c]).permute(inv).contiguous()


# ----------------------------------------------------------------------------
# BatchRenorm (model/batch_renorm.py:33-62)
# ----------------------------------------------------------------------------
class _BRNFn(t.autograd.Function):
  @staticmethod
  def forward(ctx, x, weight, bias, rm, rv, nbt, training, eps, momentum):
    _need_cuda(x, weight, bias, rm, rv, nbt)
    st = _lib.stream_ptr()
    rows_t, rows, c = _to_rows(x)
    coef = t.zeros(6 * c, dtype=t.float32, device=x.device)
    acc = t.zeros(3 * c, dtype=t.float64, device=x.device)
    if training:
      _call("crn_brn_stats", rows_t.data_ptr(), rows, c, c, 0, 0, acc.data_ptr(), st)
    _call("crn_brn_finalize", acc.data_ptr(), rows, c, weight.data_ptr(), bias.data_ptr(), rm.data_ptr(),
          rv.data_ptr(), nbt.data_ptr(), float(eps), float(momentum), int(training), coef.data_ptr(), st)
    y = t.empty_like(rows_t)
    _call("crn_brn_apply", rows_t.data_ptr(), rows, c, c, 0, coef.data_ptr(), None, 0, 0, y.data_ptr(), c, 0,
          None, st)
    ctx.save_for_backward(rows_t, coef)
    ctx.training, ctx.shape = training, x.shape
    return _from_rows(y, x.shape)

  @staticmethod
  def backward(ctx, gy):
    rows_t, coef = ctx.saved_tensors
    st = _lib.stream_ptr()
    g, rows, c = _to_rows(gy)
    acc = t.zeros(2 * c, dtype=t.float64, device=g.device)
    _call("crn_brn_bwd_reduce", g.data_ptr(), c, 0, None, None, rows_t.data_ptr(), c, 0, rows, c,
          coef.data_ptr(), 0, 0, None, acc.data_ptr(), st)
    dx = t.empty_like(g)
    dw = t.empty(c, dtype=t.float32, device=g.device)
    db = t.empty(c, dtype=t.float32, device=g.device)
    _call("crn_brn_bwd_dx", g.data_ptr(), c, 0, rows_t.data_ptr(), c, 0, rows, c, coef.data_ptr(),
          acc.data_ptr(), None, 0, int(ctx.training), dx.data_ptr(), c, 0, 0, dw.data_ptr(), db.data_ptr(),
          None, st)
    return _from_rows(dx, ctx.shape), dw, db, None, None, None, None, None, None


def batch_renorm(x, weight, bias, running_mean, running_var, num_batches_tracked, training, eps, momentum):
  return _BRNFn.apply(x, weight, bias, running_mean, running_var, num_batches_tracked, training, eps,
                      momentum)


# ----------------------------------------------------------------------------
# Generic convolution through the C-ABI (used by the standalone skip module and tests)
# ----------------------------------------------------------------------------
def pack_weight(w: t.Tensor, transposed: bool, src_cin: Optional[int] = None):
  """PyTorch conv / convT / linear weight -> (Wf, Wd, taps, CinP, CoutP) packed arenas."""
  from corenet_b200._lib import PackItem
  _need_cuda(w)
  w = w.contiguous()
  if transposed:
    cin, cout = w.shape[0], w.shape[1]
  else:
    cout, cin = w.shape[0], w.shape[1]
  taps = 1
  for s in w.shape[2:]:
    taps *= s
  cinp, coutp = _r4(cin), _r4(cout)
  n = taps * cinp * coutp
  wf = t.zeros(n, dtype=t.float32, device=w.device)
  wd = t.zeros(n, dtype=t.float32, device=w.device)
  items = (PackItem * 1)()
  it = items[0]
  it.src, it.dst_fwd, it.dst_dgrad = w.data_ptr(), wf.data_ptr(), wd.data_ptr()
  it.Cin, it.Cout, it.taps, it.CinP, it.CoutP, it.src_is_transposed = cin, cout, taps, cinp, coutp, int(transposed)
  offs = (C.c_int64 * 2)(0, n)
  items_d = t.frombuffer(bytearray(bytes(items)), dtype=t.uint8).to(w.device)
  offs_d = t.frombuffer(bytearray(bytes(offs)), dtype=t.uint8).to(w.device)
  _call("crn_pack_weights", items_d.data_ptr(), offs_d.data_ptr(), 1, n, _lib.stream_ptr())
  return wf, wd, taps, cinp, coutp


def unpack_wgrad(dw_packed: t.Tensor, like: t.Tensor, transposed: bool):
  from corenet_b200._lib import UnpackItem
  if transposed:
    cin, cout = like.shape[0], like.shape[1]
  else:
    cout, cin = like.shape[0], like.shape[1]
  taps = like.numel() // (cin * cout)
  out = t.empty_like(like, memory_format=t.contiguous_format)
  items = (UnpackItem * 1)()
  it = items[0]
  it.src_packed, it.dst = dw_packed.data_ptr(), out.data_ptr()
  it.Cin, it.Cout, it.taps, it.CinP, it.CoutP, it.dst_is_transposed = cin, cout, taps, _r4(cin), _r4(cout), int(transposed)
  n = like.numel()
  offs = (C.c_int64 * 2)(0, n)
  items_d = t.frombuffer(bytearray(bytes(items)), dtype=t.uint8).to(like.device)
  offs_d = t.frombuffer(bytearray(bytes(offs)), dtype=t.uint8).to(like.device)
  _call("crn_unpack_wgrads", items_d.data_ptr(), offs_d.data_ptr(), 1, n, _lib.stream_ptr())
  return out


def make_desc(n, cin, cout, idims, odims, k, stride, pad, transposed, x_cs, y_cs, planar=False,
              bias_n_stride=0) -> ConvDesc:
  d = ConvDesc()
  d.N, d.Cin, d.Cout = n, cin, cout
  d.iD, d.iH, d.iW = idims
  d.oD, d.oH, d.oW = odims
  d.kD, d.kH, d.kW = k
  d.stride, d.pad, d.transposed = stride, pad, int(transposed)
  d.x_cs, d.x_co, d.y_cs, d.y_co = x_cs, 0, y_cs, 0
  d.CinP, d.CoutP = _r4(cin), _r4(cout)
  d.y_planar, d.bias_n_stride = int(planar), bias_n_stride
  return d


def gemm_tc_pack(weights, dgrads):
  """Packs conv parameters [Cout][Cin][taps...] for crn_conv_gemm_tc with ONE launch; returns one tensor per item."""
  import torch as _t
  dev = weights[0].device
  lib = _lib.lib()
  n = len(weights)
  items = (_lib.GemmTcPackItem * n)()
  offs = (C.c_int64 * (n + 1))()
  outs, tot = [], 0
  for i, (w, dg) in enumerate(zip(weights, dgrads)):
    assert w.is_cuda and w.is_contiguous() and w.dtype == _t.float32
    cout, cin = w.shape[0], w.shape[1]
    taps = w[0, 0].numel()
    K, N = (cout, cin) if dg else (cin, cout)
    nfl = lib.crn_gemm_tc_packed_floats(K, N, taps)
    o = _t.zeros(nfl, dtype=_t.float32, device=dev)
    outs.append(o)
    it = items[i]
    it.src, it.dst, it.Cout, it.Cin, it.taps, it.dgrad = w.data_ptr(), o.data_ptr(), cout, cin, taps, int(dg)
    offs[i] = tot
    tot += nfl // 2
  offs[n] = tot
  items_d = _t.frombuffer(bytearray(bytes(items)), dtype=_t.uint8).to(dev)
  offs_d = _t.frombuffer(bytearray(bytes(offs)), dtype=_t.uint8).to(dev)
  _call("crn_gemm_tc_pack", items_d.data_ptr(), offs_d.data_ptr(), n, tot, _lib.stream_ptr())
  return outs


def _dims3(sp):
  sp = list(sp)
  return tuple([1] * (3 - len(sp)) + sp)


class _ConvFn(t.autograd.Function):
  """conv / conv_transpose in the reference's NC(D)HW layout through the C-ABI."""

  @staticmethod
  def forward(ctx, x, w, b, stride, pad, transposed, output_padding):
    _need_cuda(x, w, b)
    st = _lib.stream_ptr()
    n = x.shape[0]
    k = _dims3(w.shape[2:])
    idims = _dims3(x.shape[2:])
    if transposed:
      cin, cout = w.shape[0], w.shape[1]
      odims = tuple(1 if (i == 1 and kk == 1) else (i - 1) * stride - 2 * (pad if kk > 1 else 0) + kk + output_padding
                    for i, kk in zip(idims, k))
    else:
      cout, cin = w.shape[0], w.shape[1]
      odims = tuple(1 if (i == 1 and kk == 1) else (i + 2 * (pad if kk > 1 else 0) - kk) // stride + 1
                    for i, kk in zip(idims, k))
    xr, rows, _ = _to_rows(x)
    cs_in, cs_out = _r4(cin), _r4(cout)
    if cs_in != cin:
      xp = t.zeros(rows, cs_in, dtype=t.float32, device=x.device)
      xp[:, :cin] = xr
      xr = xp
    wf, wd, taps, cinp, coutp = pack_weight(w, transposed)
    orows = n * odims[0] * odims[1] * odims[2]
    y = t.zeros(orows, cs_out, dtype=t.float32, device=x.device)
    d = make_desc(n, cin, cout, idims, odims, k, stride, pad, transposed, cs_in, cs_out)
    _call("crn_conv_fwd", C.byref(d), xr.data_ptr(), wf.data_ptr(), b.data_ptr() if b is not None else None,
          y.data_ptr(), 0, st)
    ctx.save_for_backward(xr, w)
    ctx.d, ctx.wd, ctx.meta = d, wd, (n, cin, cout, idims, odims, x.shape, transposed, b is not None)
    oshape = [n, cout] + list(odims[3 - (x.dim() - 2):])
    return _from_rows(y[:, :cout].contiguous(), oshape)

  @staticmethod
  def backward(ctx, gy):
    xr, w = ctx.saved_tensors
    st = _lib.stream_ptr()
    n, cin, cout, idims, odims, xshape, transposed, has_b = ctx.meta
    d = ctx.d
    g, orows, _ = _to_rows(gy)
    cs_out = _r4(cout)
    if cs_out != cout:
      gp = t.zeros(orows, cs_out, dtype=t.float32, device=g.device)
      gp[:, :cout] = g
      g = gp
    dx = t.zeros_like(xr)
    _call("crn_conv_dgrad", C.byref(d), g.data_ptr(), ctx.wd.data_ptr(), dx.data_ptr(), 0, st)
    dwp = t.zeros(ctx.wd.numel(), dtype=t.float32, device=g.device)
    _call("crn_conv_wgrad", C.byref(d), xr.data_ptr(), g.data_ptr(), dwp.data_ptr(), st)
    dw = unpack_wgrad(dwp, w, transposed)
    db = g[:, :cout].sum(0) if has_b else None
    return _from_rows(dx[:, :cin].contiguous(), xshape), dw, db, None, None, None, None


def conv(x, w, b=None, stride=1, padding=0):
  return _ConvFn.apply(x, w, b, stride, padding, False, 0)


def conv_transpose(x, w, b=None, stride=1, padding=0, output_padding=0):
  return _ConvFn.apply(x, w, b, stride, padding, True, output_padding)


# ----------------------------------------------------------------------------
# Ray-traced skip (model/ray_traced_skip_connection.py:53-144)
# ----------------------------------------------------------------------------
class _SkipGatherFn(t.autograd.Function):
  @staticmethod
  def forward(ctx, cmap_nchw, res3d, matrix, offsets):
    _need_cuda(cmap_nchw, matrix, offsets)
    st = _lib.stream_ptr()
    b, c, h, w = cmap_nchw.shape
    assert c % 4 == 0, "compressed channel count must be a multiple of 4"
    rows, _, _ = _to_rows(cmap_nchw)
    gd, gh, gw = res3d
    out = t.empty(b * gd * gh * gw, c, dtype=t.float32, device=cmap_nchw.device)
    m = matrix.contiguous()
    o = offsets.contiguous()
    _call("crn_skip_sample_fwd", rows.data_ptr(), b, h, w, c, c, m.data_ptr(), o.data_ptr(), gd, gh, gw,
          out.data_ptr(), c, 0, st)
    ctx.save_for_backward(m, o)
    ctx.meta = (b, c, h, w, gd, gh, gw)
    return _from_rows(out, (b, c, gd, gh, gw))

  @staticmethod
  def backward(ctx, gy):
    m, o = ctx.saved_tensors
    b, c, h, w, gd, gh, gw = ctx.meta
    st = _lib.stream_ptr()
    g, _, _ = _to_rows(gy)
    dmap = skip_scatter_sorted(g, c, 0, b, h, w, c, m, o, (gd, gh, gw))
    return _from_rows(dmap, (b, c, h, w)), None, None, None


def skip_scatter_sorted(dout_rows, out_cs, out_co, b, h, w, c, matrix, offsets, res3d):
  """Deterministic backward of the gather: dmap rows [b*h*w, c] = per-pixel sums of the voxel gradients, voxels
  visited in sorted order (no atomics; ray_traced_skip_connection.py:135-142 under autograd is an atomic index_put)."""
  gd, gh, gw = res3d
  st = _lib.stream_ptr()
  dev = dout_rows.device
  nb = _lib.lib().crn_skip_lists_workspace_bytes(b, h, w, gd, gh, gw)
  ws = t.empty(nb, dtype=t.uint8, device=dev)
  sorted_vox = t.empty(b * gd * gh * gw, dtype=t.int32, device=dev)
  starts = t.empty(b * h * w + 1, dtype=t.int32, device=dev)
  _call("crn_skip_build_lists", b, h, w, matrix.data_ptr(), offsets.data_ptr(), gd, gh, gw, ws.data_ptr(), nb,
        sorted_vox.data_ptr(), starts.data_ptr(), st)
  dmap = t.empty(b * h * w, c, dtype=t.float32, device=dev)
  _call("crn_skip_sample_bwd_sorted", dout_rows.data_ptr(), out_cs, out_co, b, h, w, c, c, sorted_vox.data_ptr(),
        starts.data_ptr(), dmap.data_ptr(), st)
  return dmap


def sample_grid2d(grid2d, weight, bias, res3d, matrix, offsets):
  """compress conv (1x1) + project/truncate/gather."""
  compressed = conv(grid2d, weight, bias)
  return _SkipGatherFn.apply(compressed, tuple(int(v) for v in res3d), matrix.to(t.float32),
                             offsets.to(t.float32))


def skip_indices(b, hw, res3d, matrix, offsets):
  """int32[B,D,H,W] index iy*(w+2)+ix into the padded map, -1 behind the camera (test hook)."""
  _need_cuda(matrix, offsets)
  h, w = hw
  gd, gh, gw = res3d
  idx = t.empty(b, gd, gh, gw, dtype=t.int32, device=matrix.device)
  _call("crn_skip_indices", b, h, w, matrix.contiguous().data_ptr(), offsets.contiguous().data_ptr(), gd, gh,
        gw, idx.data_ptr(), _lib.stream_ptr())
  return idx


# ----------------------------------------------------------------------------
# Losses: corenet_b200/model/losses.py (the reference's module name); re-exported as part of the operator surface
# ----------------------------------------------------------------------------
from corenet_b200.model.losses import _LossFn, iou_fgbg, xent_times_iou_agnostic  # noqa: E402,F401


def softmax_channels(logits):
  _need_cuda(logits)
  logits = logits.contiguous()
  b, c = logits.shape[:2]
  s = logits[0, 0].numel()
  out = t.empty_like(logits)
  _call("crn_softmax_planar", logits.data_ptr(), b, c, s, out.data_ptr(), _lib.stream_ptr())
  return out


def argmax_confusion(logits, gt, cm=None):
  """confusion matrix cm[gt, argmax(logits)] (int64[C,C]) accumulated on device."""
  _need_cuda(logits, gt)
  logits, gt = logits.contiguous(), gt.contiguous()
  b, c = logits.shape[:2]
  s = logits[0, 0].numel()
  if cm is None:
    cm = t.zeros(c, c, dtype=t.int64, device=logits.device)
  _call("crn_argmax_confusion", logits.data_ptr(), gt.data_ptr(), int(gt.dtype == t.int64), b, c, s,
        cm.data_ptr(), _lib.stream_ptr())
  return cm


# ----------------------------------------------------------------------------
# fill_inside_voxels (cc/fill_voxels_gpu.cu) and voxelize_mesh (geometry/voxelization.py)
# ----------------------------------------------------------------------------
_KIND = {t.float32: (4, 1), t.float64: (8, 1), t.uint8: (1, 2), t.int8: (1, 0), t.int16: (2, 0),
         t.int32: (4, 0), t.int64: (8, 0)}


def fill_inside_voxels(grid: t.Tensor, inplace: bool = False) -> t.Tensor:
  if not grid.is_cuda:
    raise ValueError("Only CUDA tensors are supported by this OP")
  if grid.dim() != 4:
    raise ValueError("Expecting rank 4 tensor")
  if grid.dtype not in _KIND:
    raise ValueError(f"unsupported dtype {grid.dtype}")
  src = grid if grid.is_contiguous() else grid.contiguous()
  if inplace and src is not grid:
    raise ValueError("inplace=True needs a contiguous grid")
  out = grid if inplace else t.empty_like(src)
  n, d, h, w = grid.shape
  if grid.numel() == 0:
    return out
  es, kind = _KIND[grid.dtype]
  ws_bytes = _lib.lib().crn_fill_workspace_bytes(n, d, h, w)
  ws = t.empty(ws_bytes, dtype=t.uint8, device=grid.device)
  _call("crn_fill_inside", src.data_ptr(), out.data_ptr(), es, kind, n, d, h, w, ws.data_ptr(),
        _lib.stream_ptr())
  return out


def voxelize_mesh(triangles, tri_mesh, num_meshes, resolution, view2voxel, sub_grid_sampling,
                  image_resolution_multiplier, conservative_rasterization, projection_depth_multiplier):
  _need_cuda(triangles, tri_mesh, view2voxel)
  d, h, w = resolution
  r = int(round(max(w, h, d * projection_depth_multiplier) * image_resolution_multiplier))
  shape = (num_meshes, 2 * d + 1, 2 * h + 1, 2 * w + 1) if sub_grid_sampling else (num_meshes, d, h, w)
  grid = t.zeros(shape, dtype=t.float32, device=triangles.device)
  side = int(image_resolution_multiplier) if sub_grid_sampling else -1
  _call("crn_voxelize_mesh", triangles.contiguous().data_ptr(), tri_mesh.contiguous().data_ptr(),
        triangles.shape[0], view2voxel.contiguous().data_ptr(), num_meshes, d, h, w, r,
        int(projection_depth_multiplier), side, int(bool(conservative_rasterization)), grid.data_ptr(),
        _lib.stream_ptr())
  return grid


def merge_mesh_grids(mesh_grids, mesh_scene, labels, batch):
  _need_cuda(mesh_grids, mesh_scene, labels)
  m = mesh_grids.shape[0]
  vox = mesh_grids[0].numel()
  out = t.zeros((batch,) + tuple(mesh_grids.shape[1:]), dtype=t.int32, device=mesh_grids.device)
  _call("crn_merge_mesh_grids", mesh_grids.contiguous().data_ptr(), mesh_scene.contiguous().data_ptr(),
        labels.contiguous().data_ptr(), m, vox, out.data_ptr(), _lib.stream_ptr())
  return out


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
  _need_cuda(p, g, m, v)
  _call("crn_adam_step", p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2,
        eps, step, grad_scale, _lib.stream_ptr())
