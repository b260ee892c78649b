"""Data-parallel training step (SURVEY §8 rows e, f2): one process per GPU, scenes sharded by rank,
the gradient exchange as NCCL all-reduces over slices of ONE flat fp32 gradient buffer that overlap the
backward pass, fused Adam per slice.

Host-side mirror of TrainPipeline._process_batch (src/corenet/pipeline.py:215-240 of the reference:
zero_grad -> model -> loss -> backward [DDP all-reduce] -> Adam step), of DDP's construction-time
parameter/buffer broadcast (pipeline.py:199-200) and of the fixed-seed index shard of
distributed.DistributedSampler (src/corenet/distributed.py:204-230).
"""
from typing import List, Optional

import torch as t

from corenet_b200 import _lib
from corenet_b200 import engine as engine_lib

_call = engine_lib._call     # C-ABI launch (bracketed with CUDA events under engine.PROFILE)

SAMPLER_SEED = 0x1234       # distributed.py:216


def shard_indices(num_items: int, rank: int, world: int, pad: bool = True) -> List[int]:
  """Indices of `range(num_items)` that rank `rank` of `world` processes iterates over: the algorithm of the
  reference's DistributedSampler (distributed.py:204-224) -- a randperm seeded with 0x1234, zero-padded at the
  end to a multiple of `world` when pad=True (training; evaluation does not pad), split into contiguous blocks
  [rank*total//world, (rank+1)*total//world)."""
  total = (num_items + world - 1) // world * world if pad else num_items
  g = t.Generator()
  g.manual_seed(SAMPLER_SEED)
  indices = t.randperm(num_items, generator=g)
  indices = t.constant_pad_nd(indices, [0, total - indices.shape[0]])
  start = rank * total // world
  end = (rank + 1) * total // world
  return indices[start:end].tolist()


def allreduce_flat_grad(flat_grad: t.Tensor, world: int, group=None) -> float:
  """The path's exchange step on one (slice of the) flat gradient buffer: in-place sum-all-reduce (NCCL on the
  GPU box, gloo in the CPU tests).  Returns the scale (1/world) the optimiser kernel must apply to get DDP's
  rank average."""
  if world > 1:
    t.distributed.all_reduce(flat_grad, op=t.distributed.ReduceOp.SUM, group=group)
  return 1.0 / world


def broadcast_from_rank0(tensors, world: int, group=None) -> None:
  """What DistributedDataParallel does at construction (pipeline.py:199-200): every rank starts from rank 0's
  parameters and buffers."""
  if world > 1:
    for x in tensors:
      t.distributed.broadcast(x, src=0, group=group)


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


def grad_chunk_ranges(names, views):
  """[(lo, hi)] of the flat buffer per engine.GRAD_CHUNKS entry (each chunk is one contiguous range)."""
  ranges = []
  for ci in range(len(engine_lib.GRAD_CHUNKS)):
    idx = [i for i, n in enumerate(names) if engine_lib.grad_chunk_of(n) == ci]
    assert idx == list(range(idx[0], idx[-1] + 1)), "gradient chunk is not contiguous in parameter order"
    ranges.append((views[idx[0]][0], views[idx[-1]][0] + views[idx[-1]][1]))
  return ranges


class Trainer:
  """Owns the flat parameter / gradient / Adam-moment buffers of a CoreNet and runs train steps."""

  STATUS_EVERY = 1      # steps between (asynchronous, one step late) reads of the tcgen05 status word

  def __init__(self, model, lr: float = 4e-4, eps: float = 1e-4, betas=(0.9, 0.999),
               loss: str = "iou_fgbg", process_group=None, use_graph: bool = True,
               overlap_allreduce: bool = True, collective_in_graph: Optional[bool] = None):
    self.model = model
    # A step is ~600 kernel launches; enqueuing them from Python costs as much as running them, so after two
    # eager warm-up steps per input signature the whole step (weight re-pack, forward, loss, backward, the
    # gradient all-reduces when the backend can be captured, Adam) is captured once in a CUDA graph and replayed.
    self.use_graph = use_graph
    self._graphs = {}
    self._copy_stream = None
    self._prefetched = None
    self.graph_launches = 0      # kernels of this library inside one captured step
    self.lr, self.eps, self.betas = lr, eps, betas
    self.mode = {"iou_fgbg": 0, "xent_times_iou_agnostic": 1}[loss]
    self.pg = process_group
    dist_on = t.distributed.is_available() and t.distributed.is_initialized()
    self.world = t.distributed.get_world_size(process_group) if dist_on else 1
    self.backend = t.distributed.get_backend(process_group) if dist_on else None
    # NCCL collectives are stream-ordered and capturable; gloo's are host-synchronous
    self.collective_in_graph = (self.backend == "nccl") if collective_in_graph is None else collective_in_graph
    # the chunked, overlapped exchange needs stream-ordered collectives whenever the step is graph-captured
    self.overlap_allreduce = (overlap_allreduce and self.world > 1
                              and (self.collective_in_graph or not use_graph))
    self.flat, self.views = flatten_parameters(model)
    eng = engine_lib.get_engine(model)
    eng.invalidate()
    self.eng = eng
    dev = self.flat.device
    broadcast_from_rank0([self.flat] + [b for _, b in model.named_buffers()], self.world, process_group)
    self.grad = t.zeros_like(self.flat)
    self.m = t.zeros_like(self.flat)
    self.v = t.zeros_like(self.flat)
    self.step_count = 0
    self.step_dev = t.zeros(1, dtype=t.int32, device=dev)
    self.names = [n for n, _ in model.named_parameters()]
    self.grads = {}
    for (off, n), (name, p) in zip(self.views, model.named_parameters()):
      self.grads[name] = self.grad[off:off + n].view(p.shape)
    self.chunks = grad_chunk_ranges(self.names, self.views)
    self._comm_stream = t.cuda.Stream(device=dev) if self.overlap_allreduce else None
    self._loss_bufs = {}
    # tcgen05 status word: copied to pinned host memory after every STATUS_EVERY-th step and checked, without
    # blocking, at the start of the next one
    self._status_host = t.zeros(1, dtype=t.int32).pin_memory()
    self._status_ev = None

  # ------------------------------------------------------------------ failure surfacing
  def check_status(self, wait: bool = False) -> None:
    """Raises if a tcgen05 kernel reported an mbarrier timeout (the guarded Adam kernel has skipped the update of
    that step, the logits / loss of that step are NaN)."""
    ev = self._status_ev
    if ev is None:
      return
    if wait:
      ev.synchronize()
    if ev.query():
      self._status_ev = None
      if int(self._status_host[0]) != 0:
        raise RuntimeError("corenet_b200: a tcgen05 kernel reported an mbarrier timeout; the step was skipped "
                           "(weights unchanged). Reset with trainer.eng.tc_status.zero_() after fixing the cause.")

  def _post_status_read(self):
    if self.step_count % self.STATUS_EVERY == 0 and self.eng.dev is not None:
      self._status_host.copy_(self.eng.tc_status, non_blocking=True)
      self._status_ev = t.cuda.Event()
      self._status_ev.record()

  # ------------------------------------------------------------------ one step, enqueued on the current stream
  def _bufs(self, b, c, dev, planar_grad=True):
    key = (b, c)
    if key not in self._loss_bufs:
      self._loss_bufs[key] = dict(
          sums=t.empty(4 * b, dtype=t.float64, device=dev), loss=t.empty(1, dtype=t.float32, device=dev),
          coef=t.empty(2 * b + 1, dtype=t.float32, device=dev),
          dlogits=(t.empty((b, c) + tuple(self.model.config.decoder.resolution), dtype=t.float32, device=dev)
                   if planar_grad else None))
    return self._loss_bufs[key]

  def _adam(self, lo, hi, bump, scale):
    """Guarded fused Adam over flat[lo:hi) on the current stream."""
    n = hi - lo
    _call("crn_adam_step_guarded", self.flat.data_ptr() + 4 * lo, self.grad.data_ptr() + 4 * lo,
          self.m.data_ptr() + 4 * lo, self.v.data_ptr() + 4 * lo, n, self.lr, self.betas[0], self.betas[1],
          self.eps, self.step_dev.data_ptr(), int(bump), scale, self.eng.tc_status.data_ptr(), _lib.stream_ptr())

  def _chunk_cb(self, ci):
    """Called by the engine on the comm stream as soon as chunk ci of the flat gradient is final."""
    lo, hi = self.chunks[ci]
    scale = allreduce_flat_grad(self.grad[lo:hi], self.world, self.pg)
    self._adam(lo, hi, ci == 0, scale)

  def _step_body(self, image, v2s, offsets, gt, with_update: bool):
    """forward + loss + backward (+ all-reduce + Adam when with_update)."""
    model, eng = self.model, self.eng
    st = _lib.stream_ptr()
    b = image.shape[0]
    plan = eng.get_plan(b, image.device, True)
    # rows-mode plans (C > 4) hand over / take the logits and their gradient as channels-last rows
    rows = plan.rows_cp
    logits = plan.forward(image, v2s, offsets, model.training, rows_logits=True)
    c = model.config.decoder.num_output_channels
    s = 1
    for r in model.config.decoder.resolution:
      s *= r
    lb = self._bufs(b, c, image.device, planar_grad=not rows)
    dlogits = plan.glog_rows() if rows else lb["dlogits"]
    is64 = int(gt.dtype == t.int64)
    _call("crn_loss_sums_l", logits.data_ptr(), rows, gt.data_ptr(), is64, b, c, s, self.mode, lb["sums"].data_ptr(),
          st)
    _call("crn_loss_finalize", lb["sums"].data_ptr(), b, c, s, self.mode, lb["loss"].data_ptr(),
          lb["coef"].data_ptr(), st)
    _call("crn_loss_bwd_l", logits.data_ptr(), rows, gt.data_ptr(), is64, b, c, s, self.mode, lb["coef"].data_ptr(),
          None, dlogits.data_ptr(), rows, st)
    bw = dict(grad_is_rows=bool(rows))
    if with_update and self.overlap_allreduce:
      plan.backward(dlogits, self.grads, chunk_cb=self._chunk_cb, comm_stream=self._comm_stream, **bw)
    else:
      plan.backward(dlogits, self.grads, **bw)
      if with_update:
        self._update_all()
    return lb["loss"]

  def _update_all(self):
    scale = allreduce_flat_grad(self.grad, self.world, self.pg)
    self._adam(0, self.flat.numel(), True, scale)

  def _weights_changed(self):
    # the fused Adam kernel writes the parameters through raw pointers (tensor._version does not move, graph
    # replays run no Python): the engine's re-pack signature is keyed on this counter
    self.eng.weights_epoch += 1

  def _update_in_body(self) -> bool:
    """Whether all-reduce + Adam are enqueued by _step_body (and therefore captured with it) or follow it eagerly
    (host-synchronous collectives, i.e. gloo, cannot be captured)."""
    return self.world == 1 or self.collective_in_graph or not self.use_graph

  def _eager_step(self, image, v2s, offsets, gt):
    in_body = self._update_in_body()
    loss = self._step_body(image, v2s, offsets, gt, in_body)
    if not in_body:
      self._update_all()
    self._weights_changed()
    return loss

  # ------------------------------------------------------------------ graph / prefetch plumbing
  @staticmethod
  def _check_inputs(image, v2s, offsets, gt):
    if gt.dtype not in (t.int32, t.int64):
      raise AssertionError(f"ground-truth grid must be int32 or int64 (model/losses.py:36), got {gt.dtype}")
    if image.dtype != t.uint8:
      raise AssertionError("image must be uint8[B,3,H,W]")

  def _gstate(self, image, v2s, offsets, gt):
    self._check_inputs(image, v2s, offsets, gt)
    key = (tuple(image.shape), tuple(gt.shape), gt.dtype, self.flat.device, self.model.training,
           engine_lib.PRECISION)
    gs = self._graphs.get(key)
    if gs is None:
      mk = lambda: [t.empty(x.shape, dtype=d or x.dtype, device=self.flat.device)
                    for x, d in ((image, None), (v2s, t.float32), (offsets, t.float32), (gt, None))]
      gs = {"calls": 0, "graph": None, "in": mk(), "stage": mk(), "ready": None, "consumed": None}
      self._graphs[key] = gs
    return gs

  def prefetch(self, image: t.Tensor, v2s: t.Tensor, offsets: t.Tensor, gt) -> None:
    """Starts the host->device copy of the NEXT step's inputs on a copy stream (into staging buffers), so that it
    overlaps the step that is currently running -- what the reference's DataLoader prefetch + .cuda() do
    (pipeline.py:215-223).  `gt` may be a callable returning the device grid: it is invoked with the copy stream
    current, which puts the device-resident ground-truth pipeline (data.batched_example.voxelize: rasterise + fill +
    label merge, the reference's voxelize_batch, pipeline.py:126-150) of step k+1 next to step k.
    The following `step()` (called without arguments) consumes the batch."""
    if self._copy_stream is None:
      self._copy_stream = t.cuda.Stream(device=self.flat.device)
    cs = self._copy_stream
    if callable(gt):
      with t.cuda.stream(cs):
        gt = gt()
    gs = self._gstate(image, v2s, offsets, gt)
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
    """One optimisation step on this rank's scenes; returns the (device) loss scalar.  Inputs may live in (pinned)
    host memory: they are always copied into the step's static device buffers first.  Called without arguments it
    consumes the batch started by `prefetch()`."""
    self.check_status()
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
    else:
      gs = self._gstate(image, v2s, offsets, gt)
      for dst, src in zip(gs["in"], (image, v2s, offsets, gt)):
        dst.copy_(src, non_blocking=True)
    loss = self._run(gs)
    self._post_status_read()
    return loss

  def _run(self, gs):
    if not self.use_graph or engine_lib.PROFILE is not None:
      return self._eager_step(*gs["in"])
    if gs["graph"] is None:
      if gs["calls"] < 2:                       # eager warm-up: lazy initialisation must not happen under capture
        gs["calls"] += 1
        return self._eager_step(*gs["in"])
      in_graph = self._update_in_body()
      n0 = _lib.lib().crn_launch_count()
      self.eng._ver_sig = None                  # the captured step re-packs the weights unconditionally
      g = t.cuda.CUDAGraph()
      with t.cuda.graph(g, capture_error_mode="thread_local"):
        gs["loss"] = self._step_body(*gs["in"], in_graph)
      self.graph_launches = int(_lib.lib().crn_launch_count() - n0)
      gs["graph"], gs["update_in_graph"] = g, in_graph
    gs["graph"].replay()
    if not gs["update_in_graph"]:
      self._update_all()
    self._weights_changed()
    return gs["loss"]
