"""Training losses with the reference's names and argument order.

`iou_fgbg` (FG_BG task) and `xent_times_iou_agnostic` (SEMANTIC task) are what
`TrainPipeline` selects (src/corenet/pipeline.py:154-158 of the reference); they
follow src/corenet/model/losses.py:64-114 and :144-160, fused into one pass
over the logits forward (per-scene fp64 sums -> loss and per-scene coefficients)
and one pass backward (csrc/loss.cu: crn_loss_sums / crn_loss_finalize / crn_loss_bwd).
Logits are the reference's planar float32[B, C, D, H, W], labels int64 / int32 [B, D, H, W].
"""
import torch as t

from corenet_b200 import _lib

_call = _lib.call


def _need_cuda(*xs):
  for x in xs:
    if x is not None and not x.is_cuda:
      raise ValueError("corenet_b200 ops need CUDA tensors (there is no CPU fallback)")


class _LossFn(t.autograd.Function):
  @staticmethod
  def forward(ctx, gt, logits, mode):
    _need_cuda(gt, logits)
    assert logits.dtype == t.float32 and logits.dim() == 5
    b, c, d, h, w = logits.shape
    assert gt.shape == (b, d, h, w) and gt.dtype in (t.int64, t.int32)
    st = _lib.stream_ptr()
    logits = logits.contiguous()
    gt = gt.contiguous()
    s = d * h * w
    sums = t.empty(4 * b, dtype=t.float64, device=logits.device)
    loss = t.empty(1, dtype=t.float32, device=logits.device)
    coef = t.empty(2 * b + 1, dtype=t.float32, device=logits.device)
    is64 = int(gt.dtype == t.int64)
    _call("crn_loss_sums", logits.data_ptr(), gt.data_ptr(), is64, b, c, s, mode, sums.data_ptr(), st)
    _call("crn_loss_finalize", sums.data_ptr(), b, c, s, mode, loss.data_ptr(), coef.data_ptr(), st)
    ctx.save_for_backward(logits, gt, coef)
    ctx.mode = mode
    return loss[0]

  @staticmethod
  def backward(ctx, g):
    logits, gt, coef = ctx.saved_tensors
    b, c, d, h, w = logits.shape
    st = _lib.stream_ptr()
    gs = g.reshape(1).to(t.float32).contiguous()
    dl = t.empty_like(logits)
    _call("crn_loss_bwd", logits.data_ptr(), gt.data_ptr(), int(gt.dtype == t.int64), b, c, d * h * w,
          ctx.mode, coef.data_ptr(), gs.data_ptr(), dl.data_ptr(), st)
    return None, dl, None


def iou_fgbg(gt_volume, logits, weights=None):
  if weights is not None:
    raise NotImplementedError("per-voxel loss weights are not used by the training pipeline "
                              "(pipeline.py:228) and are not implemented")
  return _LossFn.apply(gt_volume, logits, 0)


def xent_times_iou_agnostic(gt_volume, logits, weights=None):
  if weights is not None:
    raise NotImplementedError("per-voxel loss weights are not implemented")
  return _LossFn.apply(gt_volume, logits, 1)


def iou_agnostic(gt_volume, logits, weights=None):
  """losses.py:19-61 of the reference: class-agnostic soft IoU loss (the first factor of xent_times_iou_agnostic)."""
  if weights is not None:
    raise NotImplementedError("per-voxel loss weights are not implemented")
  return _LossFn.apply(gt_volume, logits, 2)


def xent(gt_volume, logits, weights=None):
  """losses.py:117-141 of the reference: mean cross entropy (the second factor of xent_times_iou_agnostic)."""
  if weights is not None:
    raise NotImplementedError("per-voxel loss weights are not implemented")
  return _LossFn.apply(gt_volume, logits, 3)


def xent_times_iou_fgbg(gt_volume, logits, weights=None):
  """losses.py:163-178 of the reference (not selected by any shipped config): two fused passes, combined by autograd."""
  return (1 + iou_fgbg(gt_volume, logits, weights)) * (1 + xent(gt_volume, logits, weights))
