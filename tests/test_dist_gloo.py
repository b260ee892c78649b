"""World-size-2 (gloo, CPU) coverage of the data-parallel host logic: scene sharding and the single
gradient exchange (flat buffer all-reduce + 1/world scaling = DDP's averaging, pipeline.py:199-200)."""
import os

import torch as t
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from corenet_b200.trainer import shard_indices, allreduce_flat_grad, broadcast_from_rank0
  idx = shard_indices(10, rank, world)
  w = t.full((3,), float(rank))          # DDP's construction-time broadcast: rank 0's values win
  broadcast_from_rank0([w], world)
  assert float(w.sum()) == 0.0
  # every rank's "gradient" = sum over its scenes of a known per-scene vector
  per_scene = t.arange(10, dtype=t.float32)[:, None] * t.ones(1, 5)
  flat = per_scene[idx].sum(0)
  scale = allreduce_flat_grad(flat, world)
  avg = flat * scale
  all_idx = [None] * world
  dist.all_gather_object(all_idx, idx)
  if rank == 0:
    out.put((all_idx, avg.tolist()))
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_shard_and_gradient_exchange():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29500 + os.getpid() % 1000
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  all_idx, avg = q.get(timeout=120)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert sorted(all_idx[0] + all_idx[1]) == list(range(10)) and len(all_idx[0]) == len(all_idx[1]) == 5
  expected = float(sum(range(10))) / 2       # DDP averages the rank sums
  assert all(abs(v - expected) < 1e-6 for v in avg)


def _chunk_worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from corenet_b200 import configuration
  from corenet_b200.model.core_net import CoreNet
  from corenet_b200.trainer import allreduce_flat_grad, broadcast_from_rank0, flatten_parameters, grad_chunk_ranges
  t.manual_seed(rank)                         # ranks start from DIFFERENT weights ...
  model = CoreNet(configuration.default_config(2))
  flat, views = flatten_parameters(model)
  names = [n for n, _ in model.named_parameters()]
  bufs = [b for _, b in model.named_buffers()]
  broadcast_from_rank0([flat] + bufs, world)   # ... and rank 0's win, parameters and BatchRenorm buffers alike
  ranges = grad_chunk_ranges(names, views)
  g = t.Generator().manual_seed(100 + rank)
  grad = t.randn(flat.numel(), generator=g)
  whole = grad.clone()
  scale = allreduce_flat_grad(whole, world)
  for lo, hi in ranges:                       # the order the backward pass finishes them: decoder, stage4+5, rest
    assert allreduce_flat_grad(grad[lo:hi], world) == scale
  cm = t.full((3, 3), rank + 1, dtype=t.int64)  # Evaluator.compute_metrics: confusion matrices add up over ranks
  dist.all_reduce(cm, op=dist.ReduceOp.SUM)
  res = dict(ranges=ranges, total=flat.numel(), same=bool(t.equal(grad, whole)), scale=scale,
             psum=float(flat.double().sum()), first=float(dict(model.named_parameters())[names[0]].flatten()[0]),
             rm=float(sum(b.double().sum() for b in bufs)), cm=int(cm[0, 0]))
  gathered = [None] * world
  dist.all_gather_object(gathered, res)
  if rank == 0:
    out.put(gathered)
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_chunked_exchange_equals_whole_buffer_and_rank0_broadcast():
  """The three gradient slices (engine.GRAD_CHUNKS) tile the flat buffer, all-reducing them one by one is the
  all-reduce of the whole buffer (what Trainer._chunk_cb relies on), and the construction-time broadcast leaves both
  ranks with rank 0's parameters and buffers (pipeline.py:199-200)."""
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 30500 + os.getpid() % 1000
  procs = [ctx.Process(target=_chunk_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  r0, r1 = q.get(timeout=300)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  ranges = r0["ranges"]
  assert sorted(ranges) == [ranges[2], ranges[1], ranges[0]], "flat order is [rest | stage4+5 | decoder]"
  assert sorted(ranges)[0][0] == 0 and sorted(ranges)[-1][1] == r0["total"]
  assert all(a[1] == b[0] for a, b in zip(sorted(ranges), sorted(ranges)[1:])), "slices must tile the buffer"
  assert r0["same"] and r1["same"] and r0["scale"] == 0.5
  assert r0["psum"] == r1["psum"] and r0["first"] == r1["first"] and r0["rm"] == r1["rm"]
  assert r0["cm"] == r1["cm"] == 3
