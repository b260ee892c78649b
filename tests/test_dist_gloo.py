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
