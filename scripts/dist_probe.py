"""Two-rank NCCL probe of the Trainer's exchange variants (run under torchrun; each variant is bounded by a
faulthandler dump so a hang shows where it is):  python -m torch.distributed.run --nproc-per-node 2 scripts/dist_probe.py V
V = a: eager step, single all-reduce after backward;  b: eager, chunked overlapped;  c: graph + all-reduce in graph
(no overlap);  d: graph + chunked overlapped in graph (default Trainer)."""
import faulthandler
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch as t
import torch.distributed as dist

import bench
from corenet_b200 import configuration
from corenet_b200.model.core_net import CoreNet
from corenet_b200.trainer import Trainer


def main():
  v = sys.argv[1]
  rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
  faulthandler.dump_traceback_later(70, exit=True)
  t.cuda.set_device(local)
  dev = t.device("cuda", local)
  dist.init_process_group("nccl", device_id=dev)
  t.manual_seed(0)
  model = CoreNet(configuration.default_config(2)).to(dev).train()
  kw = {"a": dict(use_graph=False, overlap_allreduce=False), "b": dict(use_graph=False, overlap_allreduce=True),
        "c": dict(use_graph=True, overlap_allreduce=False), "d": dict(use_graph=True, overlap_allreduce=True)}[v]
  tr = Trainer(model, **kw)
  print(f"[{v}] rank {rank}: trainer built (broadcast done)", flush=True)
  d_in = [x.to(dev) for x in bench.synthetic_batch(2, rank)]
  for i in range(5):
    t0 = time.time()
    loss = tr.step(*d_in)
    t.cuda.synchronize()
    print(f"[{v}] rank {rank}: step {i} loss {float(loss):.5f} {1e3 * (time.time() - t0):.0f} ms graph={tr.graph_launches}",
          flush=True)
  dist.barrier()
  t0 = time.time()
  for i in range(10):
    tr.step(*d_in)
  t.cuda.synchronize()
  print(f"[{v}] rank {rank}: 10 steps {1e2 * (time.time() - t0):.2f} ms/step", flush=True)
  dist.barrier()
  os._exit(0)     # destroy_process_group() blocks while graphs with captured NCCL kernels are alive


if __name__ == "__main__":
  main()
