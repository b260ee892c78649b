"""bench.py -- voxels/sec forward+backward at 128^3 (BASELINE.json metric), 1..8 B200 of one node.

  python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path
  torchrun ... bench.py --gpus N --steps K --warmup W           # one rank per GPU (weak scaling)
  python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU path (oracle port)

A step = one training pass of the hot path over one batch of synthetic scenes:
forward (ResNet-50 encoder, ray-traced skips, 3-D decoder -> 128^3 logits), loss
(iou_fgbg), backward, [gradient all-reduce when N>1], Adam.  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

VOX = 128 ** 3
WORKLOAD = ("h7/h5 single-object 128^3, C=2 (FG_BG, iou_fgbg), train-mode forward+loss+backward+Adam, "
            "{b} scenes/GPU (reference per-GPU batch, generate_configs.py:52), 256x256 uint8 image")


def peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    d = json.load(open(p))
    return d, "measured"
  return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
  """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
  Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
       "clocks_event_reasons.sw_power_cap")

  def __init__(self, index):
    self.index, self.proc, self.lines = index, None, []

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits", "-lms", "100"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.lines.append(line.strip())

  def stop(self):
    if not self.proc:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], 0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for l in self.lines:
      f = [x.strip() for x in l.split(",")]
      if len(f) < 7:
        continue
      try:
        sm.append(float(f[0])); mx = max(mx, float(f[1]))
      except ValueError:
        continue
      for n, v in zip(names, f[3:7]):
        if v.lower().startswith("active"):
          reasons.add(n)
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
            "samples": len(sm)}


def synthetic_batch(b, seed, classes=2):
  """Synthetic scenes of the shapes the reference feeds the model."""
  import math
  import torch as t
  from corenet_b200.geometry import transformations as tt
  g = t.Generator().manual_seed(seed)
  image = t.randint(0, 256, (b, 3, 256, 256), dtype=t.uint8, generator=g)
  cam = tt.perspective_rh(math.pi * 60 / 180, 1, 1e-4, 10) @ tt.look_at_rh(
      [.5, .5, -1.3666666 + .5], [.5, .5, .5], [0, -1, 0])
  v2s = (cam @ tt.scale([128.0] * 3).inverse())[None].expand(b, 4, 4).contiguous()
  offsets = t.full((b, 3), 0.5)
  # GT occupancy: union of 1-3 random boxes / spheres per scene, labels 1..classes-1
  zz, yy, xx = t.meshgrid([t.arange(128)] * 3, indexing="ij")
  gt = t.zeros(b, 128, 128, 128, dtype=t.int32)
  for i in range(b):
    for _ in range(int(t.randint(1, 4, (1,), generator=g))):
      c = t.randint(32, 96, (3,), generator=g)
      r = t.randint(12, 32, (3,), generator=g)
      lab = int(t.randint(1, classes, (1,), generator=g))
      if int(t.randint(0, 2, (1,), generator=g)):
        msk = ((zz - c[0]).abs() <= r[0]) & ((yy - c[1]).abs() <= r[1]) & ((xx - c[2]).abs() <= r[2])
      else:
        msk = ((zz - c[0]) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2) <= int(r[0]) ** 2
      gt[i][msk] = lab
  return image, v2s, offsets, gt


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
  """The reference's own CPU implementation of the path (oracle port of its PyTorch modules),
  all host threads, bounded sample: one scene per step."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  import torch as t
  from oracle import corenet_oracle as O
  from corenet_b200 import configuration
  from corenet_b200.model.core_net import CoreNet
  cores = os.cpu_count() or 1
  t.set_num_threads(cores)
  t.manual_seed(0)
  m = CoreNet(configuration.default_config(2))
  params = {k: v.detach().clone().requires_grad_(True) for k, v in m.named_parameters()}
  state = {k: v.clone() for k, v in m.state_dict().items()}
  state.update(params)
  opt = t.optim.Adam(list(params.values()), lr=4e-4, eps=1e-4)
  b = 1
  image, v2s, offsets, gt = synthetic_batch(b, 0)
  gt = gt.to(t.int64)

  def step():
    opt.zero_grad()
    nb = {}
    logits = O.corenet_forward(state, image, v2s, offsets, True, nb)
    loss = O.iou_fgbg(gt, logits)
    loss.backward()
    opt.step()
    for k, v in nb.items():
      state[k] = v
    return loss.item()

  for _ in range(args.warmup):
    step()
  t0 = time.perf_counter()
  for _ in range(args.steps):
    step()
  dt = time.perf_counter() - t0
  value = b * VOX * args.steps / dt
  line = {"impl": "reference", "metric": "voxels/sec fwd+bwd @128^3", "value": value, "unit": "voxels/s",
          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": WORKLOAD.format(b=args.batch), "host_threads": cores},
          "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": cores, "kind": "port",
                           "sample": f"{args.steps} steps x 1 scene (bounded sample of the {args.batch}-scene batch)"},
          "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
  print(json.dumps(line))


# ------------------------------------------------------------------------------------------ native arm
def run_native(args):
  import torch as t
  import torch.distributed as dist
  from corenet_b200 import _lib, configuration, engine
  from corenet_b200.model.core_net import CoreNet
  from corenet_b200.trainer import Trainer
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  t.cuda.set_device(local)
  dev = t.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  _lib.lib()      # raises if the CUDA library is missing: no fallback
  t.manual_seed(0)
  model = CoreNet(configuration.default_config(2)).to(dev).train()
  trainer = Trainer(model, lr=4e-4, eps=1e-4, loss="iou_fgbg")
  b = args.batch
  image, v2s, offsets, gt = synthetic_batch(b, rank)
  h_in = [x.pin_memory() for x in (image, v2s, offsets, gt)]
  d_in = [x.to(dev) for x in h_in]
  h2d = sum(x.numel() * x.element_size() for x in h_in)
  flush = t.empty(256 * 1024 * 1024, dtype=t.uint8, device=dev)   # > 126 MB L2

  def barrier():
    if world > 1:
      dist.barrier()
    t.cuda.synchronize()

  def timed(fn, steps, profile=False):
    """max-over-ranks device time of `steps` calls of fn (CUDA events on the launch stream)."""
    evs = []
    barrier()
    for _ in range(steps):
      flush.zero_()                      # L2 flush between timed iterations (outside the events)
      e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
      e0.record()
      fn()
      e1.record()
      evs.append((e0, e1))
    barrier()
    ms = sum(a.elapsed_time(b_) for a, b_ in evs)
    tt_ = t.tensor([ms], dtype=t.float64, device=dev)
    if world > 1:
      dist.all_reduce(tt_, op=dist.ReduceOp.MAX)
    return float(tt_.item())

  dev_step = lambda: trainer.step(*d_in)
  loss_host = t.empty(1, dtype=t.float32).pin_memory()

  def e2e_step():
    # a training loop prefetches: the H2D copy of the NEXT step's inputs (pinned host memory -> staging buffers, copy
    # stream) is issued right after this step is enqueued and overlaps it; every timed step contains one full H2D
    # of a batch and the synchronous D2H read of its loss
    loss = trainer.step()                           # consumes the prefetched batch (D2D into the graph's inputs)
    trainer.prefetch(*h_in)                         # H2D of the next step's inputs
    loss_host.copy_(loss, non_blocking=False)       # D2H read of this step's result

  for _ in range(max(args.warmup, 3)):
    dev_step()
  sampler = ClockSampler(local)
  sampler.start()
  n0 = _lib.lib().crn_launch_count()
  ms = timed(dev_step, args.steps)
  # kernels of this library per step: counted at graph capture (replays do not pass through the C-ABI counter)
  launches = trainer.graph_launches * args.steps + (_lib.lib().crn_launch_count() - n0)
  trainer.prefetch(*h_in)
  for _ in range(2):
    e2e_step()
  ms_e2e = timed(e2e_step, args.steps)
  clocks = sampler.stop()                           # sampled across both timed regions (device-resident and e2e)
  trainer.step()                                    # drain the last prefetched batch
  # per-kernel roofline pass: the same step enqueued eagerly with CUDA events around every convolution launch
  # (a replayed graph has no per-launch events); same inputs, same process, right after the timed region
  engine.PROFILE = []
  engine.WGRAD_SIDE_STREAM = False          # kernels timed alone (in the graph the weight gradients run concurrently)
  timed(dev_step, args.steps)
  prof, engine.PROFILE = engine.PROFILE, None
  engine.WGRAD_SIDE_STREAM = True
  tc_status = int(engine.get_engine(model).tc_status)
  if tc_status != 0:
    raise RuntimeError("tcgen05 conv kernel reported a barrier timeout: results are invalid")
  total_vox = world * b * VOX * args.steps
  value = total_vox / (ms * 1e-3)
  e2e_value = total_vox / (ms_e2e * 1e-3)

  if rank == 0:
    pk, pk_src = peaks()
    # roofline of the dominant kernel family (conv launches timed live with CUDA events)
    fam = {}
    for kind, name, macs, e0, e1 in prof:
      k = {"wgrad": "wgrad_kernels(ffma)", "wgrad_tc": "wgrad_tc_kernel(tcgen05)", "fwd_gt": "gemm_tc_kernel(tcgen05)",
           "dgrad_gt": "gemm_tc_kernel(tcgen05)", "fwd_tc": "conv_tc5_kernel(tcgen05)",
           "dgrad_tc": "conv_tc5_kernel(tcgen05)"}.get(kind, "conv_fwd_dgrad_kernels(ffma)")
      f = fam.setdefault(k, [0.0, 0.0, 0])
      f[0] += e0.elapsed_time(e1); f[1] += 2.0 * macs; f[2] += 1
    if args.layers:
      per = {}
      for kind, name, macs, e0, e1 in prof:
        q = per.setdefault((name, kind), [0.0, 0.0])
        q[0] += e0.elapsed_time(e1) / args.steps; q[1] += 2.0 * macs / args.steps
      rows_ = sorted(((v[0], k[0], k[1], v[1] / (v[0] * 1e-3) / 1e12) for k, v in per.items()), reverse=True)
      with open(args.layers, "w") as f:
        for ms_, n_, k_, tf_ in rows_:
          f.write(f"{ms_:9.3f} ms  {tf_:7.2f} TFLOP/s  {k_:6s} {n_}\n")
    dom = max(fam.items(), key=lambda kv: kv[1][0])
    conv_ms = sum(v[0] for v in fam.values())
    achieved = dom[1][1] / (dom[1][0] * 1e-3) / 1e12
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    # dram bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/ncu_traffic.json:
    # {family: {"bytes_per_launch": ..., "source": ...}}), null if that family has no capture
    traffic, traffic_src = None, None
    try:
      tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
      if dom[0] in tj:
        traffic, traffic_src = tj[dom[0]]["bytes_per_launch"], tj[dom[0]]["source"]
    except Exception:
      pass
    roof = {"bound": "tensor", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": f"{pk_src} bf16 sustained",
            "launches": dom[1][2], "kernel_ms_per_step": dom[1][0] / args.steps,
            "all_conv_ms_per_step": conv_ms / args.steps,
            "families": {k: {"ms_per_step": v[0] / args.steps, "tflops": v[1] / (v[0] * 1e-3) / 1e12}
                         for k, v in fam.items()},
            "note": "algorithmic FLOPs = 2*MACs of each conv launch / CUDA-event time of that launch, from an eager "
                    "pass of the same step right after the timed region (the timed region replays a CUDA graph); "
                    "tcgen05 kernels run 3xTF32 (3 MMAs per product) for fp32-class accuracy"}
    line = {"metric": "voxels/sec fwd+bwd @128^3", "value": value, "unit": "voxels/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD.format(b=b), "global_batch": world * b, "parallelism": f"dp{world}",
                       "l2": "256 MiB buffer written between timed iterations",
                       "cuda_graph": bool(trainer.graph_launches),
                       "scenes_per_sec": value / VOX},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "voxels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps,
                    "note": "Trainer.step()/prefetch() with pinned HOST inputs: every timed step issues the full H2D of a "
                            "batch (copy stream, overlapping the running step like a prefetching data loader) and "
                            "synchronously reads the step's loss back"},
            "gpu_launches": int(launches), "roofline": roof}
    if world == 1 and not args.no_cpu_baseline:
      line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def cpu_baseline():
  """Oracle port of the reference timed on this box's host cores: bounded sample (1 scene/step)."""
  import torch as t
  from oracle import corenet_oracle as O
  from corenet_b200 import configuration
  from corenet_b200.model.core_net import CoreNet
  cores = os.cpu_count() or 1
  t.set_num_threads(cores)
  t.manual_seed(0)
  m = CoreNet(configuration.default_config(2))
  params = {k: v.detach().clone().requires_grad_(True) for k, v in m.named_parameters()}
  state = {k: v.clone() for k, v in m.state_dict().items()}
  state.update(params)
  image, v2s, offsets, gt = synthetic_batch(1, 0)
  gt = gt.to(t.int64)
  ts = []
  for i in range(3):
    t0 = time.perf_counter()
    for p in params.values():
      p.grad = None
    loss = O.iou_fgbg(gt, O.corenet_forward(state, image, v2s, offsets, True, {}))
    loss.backward()
    ts.append(time.perf_counter() - t0)
  dt = sorted(ts[1:])[0]
  return {"value": VOX / dt, "unit": "voxels/s", "cores": cores, "kind": "port",
          "sample": "1 scene fwd+loss+bwd, best of 2 after 1 warm-up (oracle/corenet_oracle.py, torch CPU fp32)"}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--batch", type=int, default=4, help="scenes per GPU")
  ap.add_argument("--impl", default="native", choices=["native", "reference"])
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--layers", default=None, help="write a per-layer conv timing table to this file")
  args = ap.parse_args()
  if args.impl == "reference":
    run_reference(args)
  else:
    run_native(args)


if __name__ == "__main__":
  main()
