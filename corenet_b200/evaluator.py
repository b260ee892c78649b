"""Evaluation step (BASELINE config h7; SURVEY §8 rows f1, d): eval-mode forward -> class probabilities ->
argmax -> confusion matrix, accumulated on the device.

Host-side mirror of the inner loop of EvalPipeline.run_eval (src/corenet/pipeline.py:318-327 of the reference:
`pmf = inference_fn(...)`; `quantitative_results.add_batch(pmf, batch)`), of QuantitativeResults
(evaluation_results.py:240-266: confusion matrix accumulated over batches, reduced over ranks, mean IoU over the
non-void classes) and of compute_voxel_metrics' IoU (voxel_metrics.py:61-107).  One batch = one replayed CUDA graph:
weight re-pack is skipped (weights do not change), forward, fused softmax, fused argmax + confusion counts.
"""
from typing import Optional

import torch as t

from corenet_b200 import _lib
from corenet_b200 import engine as engine_lib

_call = engine_lib._call     # C-ABI launch (bracketed with CUDA events under engine.PROFILE)


class Evaluator:
  def __init__(self, model, num_classes: Optional[int] = None, process_group=None, use_graph: bool = True):
    """num_classes: size of the confusion matrix (dataset classes incl. void).  Defaults to the number of output
    channels (SEMANTIC task); a larger value selects the FG_BG form where labels are scaled by the scene's class."""
    self.model = model
    self.c = model.config.decoder.num_output_channels
    self.k = num_classes or self.c
    self.pg = process_group
    self.use_graph = use_graph
    self.eng = engine_lib.get_engine(model)
    self._graphs = {}
    self._copy_stream = None
    self._prefetched = None
    self.graph_launches = 0
    self.confusion_matrix = None
    self.pmf = None

  # ------------------------------------------------------------------ one batch on the current stream
  def _body(self, image, v2s, offsets, gt, labels, pack=True):
    st = _lib.stream_ptr()
    b = image.shape[0]
    plan = self.eng.get_plan(b, image.device, False)
    logits = plan.forward(image, v2s, offsets, False, pack=pack, rows_logits=True)
    rows = plan.rows_cp          # > 0: channels-last rows straight from the logits layer's epilogue (C > 4)
    s = self.pmf[0, 0].numel()
    _call("crn_softmax_l", logits.data_ptr(), rows, b, self.c, s, self.pmf.data_ptr(), st)
    _call("crn_argmax_confusion_l", logits.data_ptr(), rows, gt.data_ptr(), int(gt.dtype == t.int64), b, self.c, s,
          labels.data_ptr() if labels is not None else None, self.k, self.confusion_matrix.data_ptr(), st)

  def _gstate(self, image, v2s, offsets, gt, labels):
    if self.model.training:
      raise RuntimeError("Evaluator needs model.eval() (the reference evaluates with eval-mode BatchRenorm)")
    if gt.dtype not in (t.int32, t.int64):
      raise AssertionError(f"ground-truth grid must be int32 or int64 (voxel_metrics.py:46), got {gt.dtype}")
    if (labels is None) != (self.k == self.c):
      raise ValueError("scene labels are needed exactly when num_classes != num_output_channels (FG_BG task)")
    dev = next(self.model.parameters()).device
    key = (tuple(image.shape), tuple(gt.shape), gt.dtype, labels is not None, engine_lib.PRECISION)
    gs = self._graphs.get(key)
    if gs is None:
      xs = [(image, None), (v2s, t.float32), (offsets, t.float32), (gt, None)]
      if labels is not None:
        xs.append((labels, t.int32))
      mk = lambda: [t.empty(x.shape, dtype=d or x.dtype, device=dev) for x, d in xs]
      gs = {"calls": 0, "graph": None, "in": mk(), "stage": mk(), "ready": None, "consumed": None}
      self._graphs[key] = gs
    if self.confusion_matrix is None:
      self.confusion_matrix = t.zeros(self.k, self.k, dtype=t.int64, device=dev)
    b = image.shape[0]
    shape = (b, self.c) + tuple(self.model.config.decoder.resolution)
    if self.pmf is None or tuple(self.pmf.shape) != shape:
      self.pmf = t.empty(shape, dtype=t.float32, device=dev)
    return gs

  def prefetch(self, image, v2s, offsets, gt, labels=None) -> None:
    """Host->device copy of the NEXT batch on a copy stream (pinned host memory), overlapping the running batch."""
    gs = self._gstate(image, v2s, offsets, gt, labels)
    if self._copy_stream is None:
      self._copy_stream = t.cuda.Stream(device=gs["in"][0].device)
    cs = self._copy_stream
    if gs["consumed"] is not None:
      cs.wait_event(gs["consumed"])
    with t.cuda.stream(cs):
      for dst, src in zip(gs["stage"], [x for x in (image, v2s, offsets, gt, labels) if x is not None]):
        dst.copy_(src, non_blocking=True)
      gs["ready"] = t.cuda.Event()
      gs["ready"].record(cs)
    self._prefetched = gs

  def add_batch(self, image=None, v2s=None, offsets=None, gt=None, labels=None) -> t.Tensor:
    """Evaluates one batch and adds its counts to `confusion_matrix` (int64[K, K], rows = ground truth, columns =
    prediction).  Returns the class probabilities float32[B, C, D, H, W] of the batch (a buffer owned by the
    evaluator, overwritten by the next call).  Without arguments it consumes the prefetched batch."""
    if image is None:
      gs = self._prefetched
      assert gs is not None, "add_batch() without arguments needs a preceding prefetch()"
      self._prefetched = None
      main = t.cuda.current_stream()
      main.wait_event(gs["ready"])
      for dst, src in zip(gs["in"], gs["stage"]):
        dst.copy_(src, non_blocking=True)
      gs["consumed"] = t.cuda.Event()
      gs["consumed"].record(main)
    else:
      gs = self._gstate(image, v2s, offsets, gt, labels)
      for dst, src in zip(gs["in"], [x for x in (image, v2s, offsets, gt, labels) if x is not None]):
        dst.copy_(src, non_blocking=True)
    ins = list(gs["in"]) + [None] * (5 - len(gs["in"]))
    if not self.use_graph or engine_lib.PROFILE is not None:
      self._body(*ins)
      return self.pmf
    if gs["graph"] is None:
      if gs["calls"] < 2:
        gs["calls"] += 1
        self._body(*ins)
        return self.pmf
      self.eng.pack_weights()
      self.eng.join_packs()
      n0 = _lib.lib().crn_launch_count()
      g = t.cuda.CUDAGraph()
      with t.cuda.graph(g, capture_error_mode="thread_local"):
        self._body(*ins, pack=False)
      self.graph_launches = int(_lib.lib().crn_launch_count() - n0)
      gs["graph"] = g
    self.eng.pack_weights()           # no-op unless the weights changed since the last batch
    self.eng.join_packs()
    gs["graph"].replay()
    return self.pmf

  # ------------------------------------------------------------------ metrics
  def compute_metrics(self) -> t.Tensor:
    """Sums the confusion matrix over ranks (evaluation_results.py:251-253) and returns it."""
    if t.distributed.is_available() and t.distributed.is_initialized():
      t.distributed.all_reduce(self.confusion_matrix, op=t.distributed.ReduceOp.SUM, group=self.pg)
    return self.confusion_matrix

  def tfpn(self):
    """(tp, tn, fp, fn) float64[K] per class from the confusion matrix (voxel_metrics.py:61-97: rows = ground truth,
    columns = prediction)."""
    cm = self.confusion_matrix.to(t.float64)
    tp = cm.diag()
    fp = cm.sum(0) - tp
    fn = cm.sum(1) - tp
    return tp, cm.sum() - tp - fp - fn, fp, fn

  def tfpn_fg(self):
    """Class-agnostic foreground / background counts (voxel_metrics.py:100-107): scalars tp, tn, fp, fn."""
    cm = self.confusion_matrix.to(t.float64)
    return cm[1:, 1:].sum(), cm[0, 0], cm[0, 1:].sum(), cm[1:, 0].sum()

  @staticmethod
  def _nan_tp_div(tp, y):
    # voxel_metrics.py:118-120: a class without a single true positive gets NaN (and drops out of the mean below),
    # not 0 -- the shipped behaviour of the reference (its own test expects 0 and fails)
    return t.where(tp == 0, t.full_like(tp, float("nan")), tp / y)

  def metrics(self) -> dict:
    """iou / precision / recall, float64[K + 1]: one entry per class plus the class-agnostic "__global__" entry last
    (evaluation_results.py:188-210 builds the same table as a DataFrame)."""
    out = {}
    (tp, _, fp, fn), (gtp, _, gfp, gfn) = self.tfpn(), self.tfpn_fg()
    tp, fp, fn = t.cat([tp, gtp[None]]), t.cat([fp, gfp[None]]), t.cat([fn, gfn[None]])
    out["iou"] = self._nan_tp_div(tp, tp + fp + fn)
    out["precision"] = self._nan_tp_div(tp, tp + fp)
    out["recall"] = self._nan_tp_div(tp, tp + fn)
    return out

  def iou_per_class(self) -> t.Tensor:
    """TP / (TP + FP + FN) per class, NaN where TP == 0 (voxel_metrics.py:118-135), float64[K]."""
    return self.metrics()["iou"][:-1]

  def mean_iou(self) -> float:
    """Mean IoU over the non-void classes (evaluation_results.py:262-266: a pandas mean, which skips the NaN classes;
    NaN when no class has a true positive)."""
    iou = self.iou_per_class()[1:]
    ok = ~iou.isnan()
    return float(iou[ok].mean()) if bool(ok.any()) else float("nan")

  def check_status(self) -> None:
    if int(self.eng.tc_status) != 0:
      raise RuntimeError("corenet_b200: a tcgen05 kernel reported an mbarrier timeout during evaluation")
