"""Data-parallel training step (SURVEY §8 rows e, f2): one process per GPU, scenes sharded by rank,
ONE NCCL all-reduce over a flat fp32 gradient buffer, one fused Adam launch.

Host-side mirror of TrainPipeline._process_batch (src/corenet/pipeline.py:215-240 of the reference:
zero_grad -> model -> loss -> backward [DDP all-reduce] -> Adam step) and of the fixed-seed index
shard of distributed.DistributedSampler (src/corenet/distributed.py:204-230).
"""
from typing import Optional

import torch as t

from corenet_b200 import _lib
from corenet_b200 import engine as engine_lib

_call = _lib.call


def shard_indices(num_items: int, rank: int, world: int, pad: bool = True):
  """Rank-strided shard of range(num_items) (distributed.py:204-224): rank r gets r, r+world, ...;
  with pad=True the tail is padded by wrapping so that every rank gets the same count."""
  idx = list(range(num_items))
  if pad and num_items % world:
    idx += idx[:world - num_items % world]
  return idx[rank::world]


def allreduce_flat_grad(flat_grad: t.Tensor, world: int, group=None) -> float:
  """The path's single exchange step: in-place sum-all-reduce of the flat gradient buffer (NCCL on the
  GPU box, gloo in the CPU tests).  Returns the scale (1/world) the optimiser kernel must apply to get
  DDP's rank average."""
  if world > 1:
    t.distributed.all_reduce(flat_grad, op=t.distributed.ReduceOp.SUM, group=group)
  return 1.0 / world


def flatten_parameters(model: t.nn.Module):
  """Re-points every parameter at a view of ONE flat fp32 buffer (names/shapes/state_dict unchanged)."""
  params = list(model.parameters())
  total = sum(p.numel() for p in params)
  flat = t.empty(total, dtype=t.float32, device=params[0].device)
  off = 0
  views = []
  for p in params:
    n = p.numel()
    v = flat[off:off + n].view(p.shape)
    v.copy_(p.data)
    p.data = v
    views.append((off, n))
    off += n
  return flat, views


class Trainer:
  """Owns the flat parameter / gradient / Adam-moment buffers of a CoreNet and runs train steps."""

  def __init__(self, model, lr: float = 4e-4, eps: float = 1e-4, betas=(0.9, 0.999),
               loss: str = "iou_fgbg", process_group=None, use_graph: bool = True):
    self.model = model
    # A step is ~700 kernel launches; enqueuing them from Python costs as much as running them, so after two
    # eager warm-up steps per input signature the whole step (weight re-pack, forward, loss, backward, and on a
    # single GPU the Adam update) is captured once in a CUDA graph and replayed.
    self.use_graph = use_graph
    self._graphs = {}
    self._copy_stream = None
    self._prefetched = None
    self.graph_launches = 0      # kernels of this library inside one captured step
    self.lr, self.eps, self.betas = lr, eps, betas
    self.mode = {"iou_fgbg": 0, "xent_times_iou_agnostic": 1}[loss]
    self.pg = process_group
    self.world = t.distributed.get_world_size(process_group) if t.distributed.is_initialized() else 1
    self.flat, self.views = flatten_parameters(model)
    eng = engine_lib.get_engine(model)
    eng.invalidate()
    self.eng = eng
    self.grad = t.zeros_like(self.flat)
    self.m = t.zeros_like(self.flat)
    self.v = t.zeros_like(self.flat)
    self.step_count = 0
    self.step_dev = t.zeros(1, dtype=t.int32, device=self.flat.device)
    self.names = [n for n, _ in model.named_parameters()]
    self.grads = {}
    for (off, n), (name, p) in zip(self.views, model.named_parameters()):
      self.grads[name] = self.grad[off:off + n].view(p.shape)
    self._loss_bufs = {}

  def _bufs(self, b, c, dev):
    key = (b, c)
    if key not in self._loss_bufs:
      self._loss_bufs[key] = dict(
          sums=t.empty(4 * b, dtype=t.float64, device=dev), loss=t.empty(1, dtype=t.float32, device=dev),
          coef=t.empty(2 * b + 1, dtype=t.float32, device=dev),
          dlogits=t.empty(b, c, 128, 128, 128, dtype=t.float32, device=dev))
    return self._loss_bufs[key]

  def _fwd_bwd(self, image, v2s, offsets, gt):
    """Enqueues forward, loss and backward on the current stream; gradients land in the flat buffer."""
    model, eng = self.model, self.eng
    st = _lib.stream_ptr()
    b = image.shape[0]
    plan = eng.get_plan(b, image.device, True)
    logits = plan.forward(image, v2s, offsets, model.training)
    c = logits.shape[1]
    s = logits[0, 0].numel()
    lb = self._bufs(b, c, image.device)
    is64 = int(gt.dtype == t.int64)
    _call("crn_loss_sums", logits.data_ptr(), gt.data_ptr(), is64, b, c, s, self.mode, lb["sums"].data_ptr(), st)
    _call("crn_loss_finalize", lb["sums"].data_ptr(), b, c, s, self.mode, lb["loss"].data_ptr(),
          lb["coef"].data_ptr(), st)
    _call("crn_loss_bwd", logits.data_ptr(), gt.data_ptr(), is64, b, c, s, self.mode, lb["coef"].data_ptr(),
          None, lb["dlogits"].data_ptr(), st)
    plan.backward(lb["dlogits"], self.grads)
    return lb["loss"]

  def _adam(self, scale):
    _call("crn_adam_step_dev", self.flat.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
          self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.step_dev.data_ptr(), scale,
          _lib.stream_ptr())
    self.eng._ver_sig = None   # the fused Adam kernel changed the weights: re-pack on the next forward

  def _eager_step(self, image, v2s, offsets, gt):
    loss = self._fwd_bwd(image, v2s, offsets, gt)
    scale = allreduce_flat_grad(self.grad, self.world, self.pg)
    self._adam(scale)
    return loss

  def _gstate(self, image, v2s, offsets, gt):
    key = (tuple(image.shape), tuple(gt.shape), gt.dtype, self.flat.device, self.model.training)
    gs = self._graphs.get(key)
    if gs is None:
      mk = lambda: [t.empty(x.shape, dtype=d or x.dtype, device=self.flat.device)
                    for x, d in ((image, None), (v2s, t.float32), (offsets, t.float32), (gt, None))]
      gs = {"calls": 0, "graph": None, "in": mk(), "stage": mk(), "ready": None, "consumed": None}
      self._graphs[key] = gs
    return gs

  def prefetch(self, image: t.Tensor, v2s: t.Tensor, offsets: t.Tensor, gt: t.Tensor) -> None:
    """Starts the host->device copy of the NEXT step's inputs on a copy stream (into staging buffers), so that it
    overlaps the step that is currently running -- what the reference's DataLoader prefetch + .cuda() do
    (pipeline.py:215-223).  The following `step()` (called without arguments) consumes them."""
    gs = self._gstate(image, v2s, offsets, gt)
    if self._copy_stream is None:
      self._copy_stream = t.cuda.Stream(device=self.flat.device)
    cs = self._copy_stream
    if gs["consumed"] is not None:
      cs.wait_event(gs["consumed"])            # the previous step has copied the staging buffers out
    with t.cuda.stream(cs):
      for dst, src in zip(gs["stage"], (image, v2s, offsets, gt)):
        dst.copy_(src, non_blocking=True)
      gs["ready"] = t.cuda.Event()
      gs["ready"].record(cs)
    self._prefetched = gs

  def step(self, image: Optional[t.Tensor] = None, v2s: Optional[t.Tensor] = None,
           offsets: Optional[t.Tensor] = None, gt: Optional[t.Tensor] = None) -> t.Tensor:
    """One optimisation step on this rank's scenes; returns the (device) loss scalar.  In graph mode the inputs
    may live in (pinned) host memory: they are copied straight into the graph's static device buffers.  Called
    without arguments it consumes the batch started by `prefetch()`."""
    self.step_count += 1
    if image is None:
      gs = self._prefetched
      assert gs is not None, "step() without arguments needs a preceding prefetch()"
      self._prefetched = None
      main = t.cuda.current_stream()
      main.wait_event(gs["ready"])
      for dst, src in zip(gs["in"], gs["stage"]):
        dst.copy_(src, non_blocking=True)        # device-to-device, ~10 us
      gs["consumed"] = t.cuda.Event()
      gs["consumed"].record(main)
      if not self.use_graph or engine_lib.PROFILE is not None:
        return self._eager_step(*gs["in"])
    else:
      if not self.use_graph or engine_lib.PROFILE is not None:
        return self._eager_step(image, v2s, offsets, gt)
      gs = self._gstate(image, v2s, offsets, gt)
      for dst, src in zip(gs["in"], (image, v2s, offsets, gt)):
        dst.copy_(src, non_blocking=True)
    if gs["graph"] is None:
      if gs["calls"] < 2:                       # eager warm-up: lazy initialisation must not happen under capture
        gs["calls"] += 1
        return self._eager_step(*gs["in"])
      n0 = _lib.lib().crn_launch_count()
      self.eng._ver_sig = None
      g = t.cuda.CUDAGraph()
      with t.cuda.graph(g):
        gs["loss"] = self._fwd_bwd(*gs["in"])
        if self.world == 1:
          self._adam(1.0)
      self.graph_launches = int(_lib.lib().crn_launch_count() - n0)
      gs["graph"] = g
    gs["graph"].replay()
    if self.world > 1:
      scale = allreduce_flat_grad(self.grad, self.world, self.pg)
      self._adam(scale)
    return gs["loss"]
