"""CPU-side enqueue time of one training step (GPU idle at the start, no sync inside) vs its GPU time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
from bench import synthetic_batch
from corenet_b200 import configuration
from corenet_b200.model.core_net import CoreNet
from corenet_b200.trainer import Trainer
dev = t.device("cuda", 0)
t.manual_seed(0)
model = CoreNet(configuration.default_config(2)).to(dev).train()
tr = Trainer(model, lr=4e-4, eps=1e-4, loss="iou_fgbg")
d_in = [x.to(dev) for x in synthetic_batch(4, 0)]
for _ in range(3):
  tr.step(*d_in)
t.cuda.synchronize()
for _ in range(3):
  e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
  t0 = time.perf_counter()
  e0.record()
  tr.step(*d_in)
  e1.record()
  t1 = time.perf_counter()
  t.cuda.synchronize()
  print(f"cpu enqueue {1e3 * (t1 - t0):.2f} ms   gpu {e0.elapsed_time(e1):.2f} ms", flush=True)
