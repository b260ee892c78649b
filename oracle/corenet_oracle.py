"""CPU oracle for the CoReNet forward/backward hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain PyTorch fp32 CPU ops, what the reference
(google-research/corenet @ 1ba76ac) computes on the path named by
BASELINE.json:north_star.  It is the checker for `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py`.  Nothing under `corenet_b200/` imports it.

Pinning: `oracle/make_golden.py` (run in the build container, where
`/root/reference` exists) imports the *real* reference, runs it on seeded
inputs, checks that this restatement reproduces it, and commits compact
fixtures to `tests/golden/`.  `tests/test_oracle_golden.py` replays those
fixtures against this file without the reference being present.

Every function cites the reference file:line it follows (paths relative to
`/root/reference/src/corenet`).  The model is a pure function of a
`state_dict` with the reference's key names, so the same weights drive the
oracle and the CUDA path.
"""
from __future__ import annotations

import math
from typing import Dict, List, NamedTuple, Optional, Tuple

import torch as t
import torch.nn.functional as F

State = Dict[str, t.Tensor]

BRN_EPS = 1e-3       # model/resnet50.py:64 and reconstruction_decoder.py:44 pass eps=0.001
BRN_MOMENTUM = 0.01  # model/batch_renorm.py:20


# --------------------------------------------------------------------------
# geometry/transformations.py (only what the hot path uses)
# --------------------------------------------------------------------------
def scale(v) -> t.Tensor:
  """transformations.py:25-38."""
  v = t.as_tensor(v, dtype=t.float32)
  return t.diag(t.cat([v, v.new_ones([1])], dim=0))


def translate(v) -> t.Tensor:
  """transformations.py:41-60 (batched translation matrices)."""
  v = t.as_tensor(v, dtype=t.float32)
  n = v.shape[-1]
  m = t.eye(n + 1, dtype=t.float32).expand(*v.shape[:-1], n + 1, n + 1).clone()
  m[..., :n, n] = v
  return m


def look_at_rh(eye, center, up) -> t.Tensor:
  """transformations.py:206-225."""
  eye = t.as_tensor(eye, dtype=t.float32)
  center = t.as_tensor(center, dtype=t.float32)
  up = t.as_tensor(up, dtype=t.float32)
  f = F.normalize(center - eye, dim=-1)
  s = F.normalize(t.linalg.cross(f, up), dim=-1)
  u = t.linalg.cross(s, f)
  return t.tensor([
      [s[0], s[1], s[2], -t.dot(s, eye)],
      [u[0], u[1], u[2], -t.dot(u, eye)],
      [-f[0], -f[1], -f[2], t.dot(f, eye)],
      [0, 0, 0, 1]], dtype=t.float32)


def perspective_rh(fov_y, aspect, z_near, z_far) -> t.Tensor:
  """transformations.py:250-270."""
  fov_y = t.as_tensor(fov_y, dtype=t.float32)
  aspect = t.as_tensor(aspect, dtype=t.float32)
  z_near = t.as_tensor(z_near, dtype=t.float32)
  z_far = t.as_tensor(z_far, dtype=t.float32)
  th = t.tan(fov_y / 2)
  return t.tensor([
      [1.0 / (aspect * th), 0, 0, 0],
      [0, 1.0 / th, 0, 0],
      [0, 0, -(z_far + z_near) / (z_far - z_near),
       -(2 * z_far * z_near) / (z_far - z_near)],
      [0, 0, -1, 0]], dtype=t.float32)


def ortho_lh(left, right, bottom, top, z_near, z_far) -> t.Tensor:
  """transformations.py:273-294."""
  l, r, b, tp, n, f = [float(x) for x in (left, right, bottom, top, z_near, z_far)]
  return t.tensor([
      [2 / (r - l), 0, 0, -(r + l) / (r - l)],
      [0, 2 / (tp - b), 0, -(tp + b) / (tp - b)],
      [0, 0, 2 / (f - n), -(f + n) / (f - n)],
      [0, 0, 0, 1]], dtype=t.float32)


def transform_points_homogeneous(points: t.Tensor, matrix: t.Tensor, w: float) -> t.Tensor:
  """transformations.py:108-136: pad with w, einsum("bnm,bvm->bvn")."""
  points = t.constant_pad_nd(points, [0, 1], value=w)
  return t.einsum("bnm,bvm->bvn", matrix, points)


def dataset_camera() -> t.Tensor:
  """The single fixed camera of the CoReNet datasets
  (doc/data_format_and_coordinate_systems.md:103-110)."""
  return perspective_rh(math.pi * 60 / 180, 1, 1e-4, 10) @ look_at_rh(
      [.5, .5, -1.3666666 + .5], [.5, .5, .5], [0, -1, 0])


def default_v2s(batch: int, res: int = 128) -> t.Tensor:
  """voxel->screen matrix used by pipeline.py:220 with batched_example.py:156."""
  cam = dataset_camera()
  return (cam @ scale([res] * 3).inverse())[None].expand(batch, 4, 4).contiguous()


# --------------------------------------------------------------------------
# model/batch_renorm.py:33-62
# --------------------------------------------------------------------------
def batch_renorm(x: t.Tensor, st: State, prefix: str, training: bool,
                 new_buffers: Optional[State] = None) -> t.Tensor:
  weight, bias = st[prefix + "weight"], st[prefix + "bias"]
  running_mean, running_var = st[prefix + "running_mean"], st[prefix + "running_var"]
  nt = st[prefix + "num_batches_tracked"]
  view = [1, x.shape[1]] + [1] * (x.dim() - 2)
  _v = lambda v: v.view(view)
  running_std = (running_var + BRN_EPS).sqrt()
  if training:
    d_max = (5.0 * (nt - 5000) / (25000 - 5000)).clamp(0.0, 5.0)
    r_max = 1.0 + (2.0 * (nt - 5000) / (40000 - 5000)).clamp(0.0, 2.0)
    reduce_dims = [i for i in range(x.dim()) if i != 1]
    b_mean = x.mean(reduce_dims)
    b_var = x.var(reduce_dims, unbiased=False)
    b_std = (b_var + BRN_EPS).sqrt()
    r = (b_std.detach() / running_std).clamp(1 / r_max, r_max)
    d = ((b_mean.detach() - running_mean) / running_std)
    d = t.max(t.min(d, d_max), -d_max)
    x = (x - _v(b_mean)) / _v(b_std) * _v(r) + _v(d)
    if new_buffers is not None:
      unbiased_var = b_var.detach() * x.shape[1] / (x.shape[1] - 1)   # quirk: channel count
      new_buffers[prefix + "running_var"] = running_var + BRN_MOMENTUM * (unbiased_var - running_var)
      new_buffers[prefix + "running_mean"] = running_mean + BRN_MOMENTUM * (b_mean.detach() - running_mean)
      new_buffers[prefix + "num_batches_tracked"] = nt + 1
  else:
    x = (x - _v(running_mean)) / _v(running_std)
  return _v(weight) * x + _v(bias)


# --------------------------------------------------------------------------
# model/resnet50.py
# --------------------------------------------------------------------------
def preprocess_image_caffe(image: t.Tensor) -> t.Tensor:
  """resnet50.py:189-204 (adds the BGR mean, bug-for-bug)."""
  assert image.dtype == t.uint8 and image.dim() == 4 and image.shape[1] == 3
  image = image.to(t.float32).flip(1)
  return image + image.new_tensor([103.939, 116.779, 123.68])[None, :, None, None]


class Features(NamedTuple):
  stage1_64x128x128: t.Tensor
  stage2_256x64x64: t.Tensor
  stage3_512x32x32: t.Tensor
  stage4_1024x16x16: t.Tensor
  stage5_2048x8x8: t.Tensor
  global_average_2048: t.Tensor


def _conv_bn(x, st, p, training, nb, stride=1, padding=0):
  x = F.conv2d(x, st[p + "conv.weight"], st[p + "conv.bias"], stride=stride, padding=padding)
  return batch_renorm(x, st, p + "bn.", training, nb)


def _rec(taps, key, v):
  if taps is not None:
    taps[key] = v
  return v


def _identity_block(x, st, p, training, nb, taps=None):
  """resnet50.py:72-83."""
  inp = x
  x = _rec(taps, p + "a_y", _conv_bn(x, st, p + "op_a.", training, nb).relu())
  x = _rec(taps, p + "b_y", _conv_bn(x, st, p + "op_b.", training, nb, padding=1).relu())
  x = _conv_bn(x, st, p + "op_c.", training, nb) + inp
  return _rec(taps, p + "out", x.relu()), x


def _downscale_block(x, st, p, training, nb, stride, taps=None):
  """resnet50.py:110-115."""
  s = _conv_bn(x, st, p + "shortcut.", training, nb, stride=stride)
  x = _rec(taps, p + "a_y", _conv_bn(x, st, p + "op_a.", training, nb, stride=stride).relu())
  x = _rec(taps, p + "b_y", _conv_bn(x, st, p + "op_b.", training, nb, padding=1).relu())
  x = (_conv_bn(x, st, p + "op_c.", training, nb) + s).relu()
  return _rec(taps, p + "out", x)


ENCODER_STAGES = (("stage2", "abc", 1), ("stage3", "abcd", 2),
                  ("stage4", "abcdef", 2), ("stage5", "abc", 2))


def resnet50_features(st: State, image_f32: t.Tensor, training: bool,
                      nb: Optional[State] = None, prefix: str = "encoder.", taps=None) -> Features:
  """resnet50.py:176-186."""
  p = prefix
  x = F.pad(image_f32, [3, 3, 3, 3])
  x = stage1 = F.conv2d(x, st[p + "stage1.conv.weight"], st[p + "stage1.conv.bias"], stride=2)
  x = batch_renorm(x, st, p + "stage1_part2.bn.", training, nb).relu()
  x = F.max_pool2d(F.pad(x, [1, 1, 1, 1]), kernel_size=3, stride=2)
  outs = []
  for name, blocks, stride in ENCODER_STAGES:
    x = _downscale_block(x, st, f"{p}{name}.a.", training, nb, stride, taps)
    pre = None
    for b in blocks[1:]:
      x, pre = _identity_block(x, st, f"{p}{name}.{b}.", training, nb, taps)
    outs.append(pre)
  return Features(stage1, outs[0], outs[1], outs[2], outs[3], x.mean(dim=(2, 3)))


# --------------------------------------------------------------------------
# model/ray_traced_skip_connection.py:53-144
# --------------------------------------------------------------------------
def sample_grid2d_indices(batch: int, res3d: Tuple[int, int, int], hw: Tuple[int, int],
                          matrix: t.Tensor, offsets: t.Tensor):
  """Index part of SampleGrid2d.forward (:91-123,138-142): returns
  (ix, iy) into the 1-pixel padded map, int64[B,D,H,W], and the in-front mask."""
  gd, gh, gw = res3d
  height, width = hw
  zz, yy, xx = t.meshgrid([t.arange(0, gd, dtype=t.float32), t.arange(0, gh, dtype=t.float32),
                           t.arange(0, gw, dtype=t.float32)], indexing="ij")
  centers = t.stack([xx, yy, zz], dim=-1)
  centers = centers[None].expand(batch, gd, gh, gw, 3).contiguous()
  centers = centers + offsets[:, None, None, None, :]
  centers = centers.reshape([batch, -1, 3])
  proj = transform_points_homogeneous(centers, matrix, w=1).reshape([batch, gd, gh, gw, 4])
  depth = proj[..., 2]
  proj = proj[..., :3] / proj[..., 3:4]
  proj = proj[..., :2] / 2 + 0.5
  wh = proj.new_tensor([[[[[width, height]]]]], dtype=t.float32)
  pix = (proj * wh).to(t.int64)
  ix, iy = pix.unbind(-1)
  ix = (ix + 1).clamp(0, width + 1)
  iy = (iy + 1).clamp(0, height + 1)
  return ix, iy, depth >= 0


def sample_grid2d(grid2d: t.Tensor, weight: t.Tensor, bias: t.Tensor, res3d, matrix: t.Tensor,
                  offsets: t.Tensor, outside_value: float = 0.0) -> t.Tensor:
  compressed = F.conv2d(grid2d, weight, bias)
  b, c, h, w = compressed.shape
  # the index arithmetic is fp32 in the reference whatever the dtype of the features
  ix, iy, front = sample_grid2d_indices(b, tuple(res3d), (h, w), matrix.float(), offsets.float())
  padded = t.constant_pad_nd(compressed, [1, 1, 1, 1], value=outside_value)
  bb = t.arange(b, dtype=t.int64)[:, None, None, None].expand_as(ix)
  result = padded[bb, :, iy, ix].permute([0, 4, 1, 2, 3])
  return t.where(front[:, None].expand(result.shape), result,
                 t.ones_like(result) * outside_value)


# --------------------------------------------------------------------------
# model/reconstruction_decoder.py
# --------------------------------------------------------------------------
def _apply_skip(st, p, x3d, src2d, stage, v2s, offsets, resolution):
  """reconstruction_decoder.py:97-117."""
  key = f"{p}rt_skip_{stage}.compress_channels.weight"
  if key not in st:
    return x3d
  o = offsets.to(src2d.dtype)[:, :, None, None].expand(src2d.shape[0], 3, *src2d.shape[2:])
  src2d = t.cat([src2d, o], 1)
  r1 = t.tensor(x3d.shape[2:], dtype=t.float32)
  r2 = t.tensor(resolution, dtype=t.float32)
  layer_matrix = v2s.matmul(scale(r2 / r1).to(v2s.dtype))
  skip = sample_grid2d(src2d, st[key], st[f"{p}rt_skip_{stage}.compress_channels.bias"],
                       x3d.shape[2:], layer_matrix, offsets)
  return t.cat([x3d, skip], dim=1)


def decoder_forward(st: State, f: Features, v2s: t.Tensor, offsets: t.Tensor, training: bool,
                    nb: Optional[State] = None, prefix: str = "decoder.",
                    resolution=(128, 128, 128), taps: Optional[dict] = None) -> t.Tensor:
  """reconstruction_decoder.py:119-152 (ctor :32-95).  Only the 128^3 /
  last_upscale_factor=2 geometry exists."""
  p = prefix
  bn = lambda x, name: batch_renorm(x, st, p + name + ".", training, nb)
  W = lambda name: st[p + name + ".weight"]
  Bz = lambda name: st[p + name + ".bias"]
  rec = (lambda k, v: taps.__setitem__(k, v)) if taps is not None else (lambda k, v: None)

  x = F.linear(f.global_average_2048, W("stage_0"), Bz("stage_0"))
  rec("stage_0", x)
  x = t.cat([x, offsets.to(x.dtype)], 1)[:, :, None, None, None]
  ir = tuple(r // 32 for r in resolution)
  x = F.conv_transpose3d(bn(x.relu(), "stage_1.b1"), W("stage_1.t1"), Bz("stage_1.t1"), stride=ir)
  rec("stage_1", x)
  x = _apply_skip(st, p, x, f.stage5_2048x8x8, 1, v2s, offsets, resolution)
  x = F.conv3d(bn(x.relu(), "stage_2.b1"), W("stage_2.c1"), Bz("stage_2.c1"), padding=1)
  rec("stage_2.c1", x)
  x = F.conv_transpose3d(bn(x.relu(), "stage_2.b2"), W("stage_2.t1"), Bz("stage_2.t1"),
                         stride=2, padding=1, output_padding=1)
  rec("stage_2", x)
  for stage, src in ((3, f.stage5_2048x8x8), (4, f.stage4_1024x16x16),
                     (5, f.stage3_512x32x32), (6, f.stage2_256x64x64)):
    x = _apply_skip(st, p, x, src, stage - 1, v2s, offsets, resolution)
    rec(f"cat_{stage}", x)
    x = F.conv3d(bn(x.relu(), f"stage_{stage}.b1"), W(f"stage_{stage}.c1"), Bz(f"stage_{stage}.c1"),
                 padding=2)
    rec(f"stage_{stage}.c1", x)
    x = F.conv_transpose3d(bn(x.relu(), f"stage_{stage}.b2"), W(f"stage_{stage}.t1"),
                           Bz(f"stage_{stage}.t1"), stride=2, padding=3, output_padding=1)
    rec(f"stage_{stage}", x)
  return x


def corenet_forward(st: State, image: t.Tensor, v2s: t.Tensor, offsets: t.Tensor, training: bool,
                    new_buffers: Optional[State] = None, taps: Optional[dict] = None,
                    dtype: Optional[t.dtype] = None) -> t.Tensor:
  """model/core_net.py:36-43.  dtype=float64 (with a float64 state) gives the "exact" answer used to
  calibrate how much of a mismatch is fp32 rounding noise of the reference itself."""
  x = preprocess_image_caffe(image)
  if dtype is not None:
    x = x.to(dtype)
  f = resnet50_features(st, x, training, new_buffers, taps=taps)
  if taps is not None:
    for k, v in f._asdict().items():
      taps[k] = v
  return decoder_forward(st, f, v2s, offsets, training, new_buffers, taps=taps)


# --------------------------------------------------------------------------
# model/losses.py
# --------------------------------------------------------------------------
def iou_agnostic(gt, logits, weights=None):
  """losses.py:19-61."""
  b, c, d, h, w = logits.shape
  g = F.one_hot(gt, c).to(t.float32).permute([0, 4, 1, 2, 3])[:, 1:]
  pr = logits.softmax(dim=1)[:, 1:]
  fw = t.where(g == 0, t.ones_like(g), t.ones_like(g) * (c - 1.0))
  if weights is not None:
    fw = fw * weights[:, None]
  inter = (t.min(g, pr) * fw).sum(dim=[1, 2, 3, 4])
  union = (t.max(g, pr) * fw).sum(dim=[1, 2, 3, 4])
  iou = inter / t.where(union == 0, t.ones_like(union), union)
  return 1 - iou.mean()


def iou_fgbg(gt, logits, weights=None):
  """losses.py:64-114."""
  b, c, d, h, w = logits.shape
  g = F.one_hot(gt, c).to(t.float32).permute([0, 4, 1, 2, 3])[:, 1:].sum(1)
  pr = logits.softmax(dim=1)[:, 1:].sum(1)
  g = t.min(g, g.new_tensor(1.0))
  inter, union = t.min(g, pr), t.max(g, pr)
  if weights is not None:
    inter, union = inter * weights, union * weights
  inter = inter.reshape([b, -1]).sum(1)
  union = union.reshape([b, -1]).sum(1)
  iou = inter / t.where(union == 0, t.ones_like(union), union)
  return 1 - iou.mean()


def xent(gt, logits, weights=None):
  """losses.py:117-141."""
  loss = F.cross_entropy(logits, gt, reduction="none")
  if weights is not None:
    loss = loss * weights
  return loss.mean()


def xent_times_iou_agnostic(gt, logits, weights=None):
  """losses.py:144-160."""
  return (1 + iou_agnostic(gt, logits, weights)) * (1 + xent(gt, logits, weights))


# --------------------------------------------------------------------------
# voxel_metrics.py:33-58 / evaluation_results.py:262-266 (IoU oracle)
# --------------------------------------------------------------------------
def confusion_matrix(pred: t.Tensor, gt: t.Tensor, num_classes: int) -> t.Tensor:
  """cm[gt, pred] counts, int64[num_classes, num_classes]."""
  idx = gt.reshape(-1).to(t.int64) * num_classes + pred.reshape(-1).to(t.int64)
  return t.bincount(idx, minlength=num_classes * num_classes).reshape(num_classes, num_classes)


def iou_per_class(cm: t.Tensor) -> t.Tensor:
  """voxel_metrics.py:61-97,118-135: IoU = TP / (TP + FP + FN), NaN for a class without a true positive
  (nan_tp_div; the reference's own test expects 0 there and fails as shipped -- SURVEY F10)."""
  cm = cm.to(t.float64)
  tp = cm.diag()
  iou = tp / (cm.sum(0) + cm.sum(1) - tp)
  return t.where(tp == 0, t.full_like(iou, math.nan), iou)


def mean_iou(cm: t.Tensor) -> float:
  """evaluation_results.py:262-266: mean over the non-void classes of a pandas frame, i.e. NaN classes are skipped."""
  iou = iou_per_class(cm)[1:]
  ok = ~iou.isnan()
  return float(iou[ok].mean()) if bool(ok.any()) else math.nan


# --------------------------------------------------------------------------
# super_resolution.py:66-112 (host logic of the inference plug point)
# --------------------------------------------------------------------------
def native_offsets(mult: int, grid_offsets: t.Tensor) -> t.Tensor:
  zz, yy, xx = t.meshgrid([t.arange(mult)] * 3, indexing="ij")
  offs = (t.stack([xx, yy, zz], -1) / mult).reshape([-1, 3])
  return offs[:, None] + grid_offsets[None, :] / mult


def interleave_pmfs(pmfs: t.Tensor, mult: int) -> t.Tensor:
  _, b, c, d, h, w = pmfs.shape
  pmfs = pmfs.reshape([mult, mult, mult, b, c, d, h, w]).permute([3, 4, 5, 0, 6, 1, 7, 2])
  return pmfs.reshape([b, c, mult * d, mult * h, mult * w])
