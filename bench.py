"""bench.py -- the BASELINE.json metric (voxels/sec forward+backward at 128^3) and its sibling configurations,
1..8 B200 of one node.

  python bench.py [--workload h5] --gpus 1 --steps K --warmup W   # this repo's CUDA path (default: the headline)
  torchrun ... bench.py --gpus N --steps K --warmup W             # one rank per GPU (weak scaling)
  python bench.py --impl reference ...                            # the reference's CPU path (oracle port)
  python bench.py --impl reference-gpu ...                        # informational: the UNMODIFIED reference modules on
                                                                  # stock PyTorch-CUDA (staged copy in baseline/_ref)

Workloads (BASELINE.json `configs`; every run prints ONE JSON line):
  h5    train step, C=2 (FG_BG, iou_fgbg), 4 scenes/GPU: forward + loss + backward + all-reduce + Adam   [headline]
  m7    train step, C=15 (SEMANTIC, xent_times_iou_agnostic), 4 scenes/GPU, per-scale skip kernels profiled
  m9    m7 + on-the-fly ground truth: voxelise + fill + label merge of 3 meshes/scene on a side stream
  h7    evaluation, C=2, 8 scenes/GPU: eval forward + softmax + argmax + confusion matrix / IoU
  fill  fill_inside_voxels on 12 x 128^3 float32 grids, next to the reference's own CUDA kernels (K1/K2)
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

VOX = 128 ** 3
WORKLOADS = {
    # name: (classes, loss, default scenes/GPU, meshes/scene, description)
    "h5": (2, "iou_fgbg", 4, 1,
           "h7/h5 single-object 128^3, C=2 (FG_BG, iou_fgbg), train-mode forward+loss+backward+Adam, "
           "{b} scenes/GPU (reference per-GPU batch, generate_configs.py:52), 256x256 uint8 image"),
    "m7": (15, "xent_times_iou_agnostic", 4, 2,
           "m7 pairs 128^3, C=15 (SEMANTIC, xent_times_iou_agnostic), ray-traced skip at all 4 decoder scales, "
           "train-mode forward+loss+backward+Adam, {b} scenes/GPU, 256x256 uint8 image"),
    "m9": (15, "xent_times_iou_agnostic", 4, 3,
           "m9 triplets 128^3, C=15, full train step with on-the-fly ground truth (voxelise 3 meshes/scene at "
           "1024^2 samples + fill_inside_voxels + label merge on a side stream) + forward+loss+backward+Adam, "
           "{b} scenes/GPU"),
    "h7": (2, None, 8, 1,
           "h7 single-object 128^3 evaluation, C=2, eval-mode forward + softmax + argmax + confusion matrix (IoU), "
           "{b} scenes/GPU (reference eval batch, generate_configs.py:118), 256x256 uint8 image"),
    "fill": (None, None, 12, 1,
             "fill_inside_voxels on {b} x 128^3 float32 grids (voxelised icosphere / cube shells + 5% random voxels) "
             "= one m9 batch of 4 scenes x 3 meshes"),
}
METRIC = {"h5": "voxels/sec fwd+bwd @128^3", "m7": "voxels/sec fwd+bwd @128^3", "m9": "voxels/sec fwd+bwd @128^3",
          "h7": "voxels/sec fwd+IoU @128^3", "fill": "voxels/sec fill_inside_voxels @128^3"}


def peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    return json.load(open(p)), "measured"
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


# ------------------------------------------------------------------------------------------ synthetic inputs
def synthetic_batch(b, seed, classes=2):
  """Synthetic scenes of the shapes the reference feeds the model."""
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


def icosphere(level=3):
  """Unit icosphere, float32[20 * 4^level, 3, 3] (closed mesh)."""
  import numpy as np
  p = (1 + 5 ** 0.5) / 2
  v = np.array([[-1, p, 0], [1, p, 0], [-1, -p, 0], [1, -p, 0], [0, -1, p], [0, 1, p], [0, -1, -p], [0, 1, -p],
                [p, 0, -1], [p, 0, 1], [-p, 0, -1], [-p, 0, 1]], np.float64)
  v /= np.linalg.norm(v, axis=1, keepdims=True)
  f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
                [6, 2, 10], [8, 6, 7], [9, 8, 1]])
  tri = v[f]
  for _ in range(level):
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    ab, bc, ca = [(x + y) / np.linalg.norm(x + y, axis=1, keepdims=True) for x, y in ((a, b), (b, c), (c, a))]
    tri = np.concatenate([np.stack(q, 1) for q in ((a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca))])
  return tri.astype(np.float32)


def synthetic_meshes(b, meshes_per_scene, seed, classes):
  """Closed meshes inside the unit cube (icospheres, 1280 triangles each): -> (triangles float32[T,3,3],
  per-scene int32 triangle counts, per-scene labels)."""
  import numpy as np
  import torch as t
  rng = np.random.default_rng(seed)
  base = icosphere(3)
  tris, ntri, labels = [], [], []
  for _ in range(b):
    cnt, lab = [], []
    for _ in range(meshes_per_scene):
      r = rng.uniform(0.08, 0.22)
      c = rng.uniform(0.25, 0.75, 3)
      s = rng.uniform(0.6, 1.0, 3)            # ellipsoid
      tris.append((base * (r * s) + c).astype(np.float32))
      cnt.append(len(base))
      lab.append(int(rng.integers(1, classes)))
    ntri.append(t.tensor(cnt, dtype=t.int32))
    labels.append(lab)
  return t.from_numpy(np.concatenate(tris)), ntri, labels


def fill_grids(n, dev):
  """n x 128^3 float32 surface grids: voxelised ellipsoid shells (closed: their inside must be filled) + 5% noise."""
  import torch as t
  from corenet_b200.geometry import voxelization
  tri, ntri, _ = synthetic_meshes(n, 1, 4, 2)
  from corenet_b200.geometry import transformations as tt
  g = voxelization.voxelize_mesh(tri, t.cat(ntri), (128, 128, 128), tt.scale([128.0] * 3), image_resolution_multiplier=8,
                                 cuda_device=dev.index)
  noise = t.rand(g.shape, generator=t.Generator(device=dev).manual_seed(5), device=dev) < 0.05
  return t.maximum(g, noise.to(g.dtype)).contiguous()


# ------------------------------------------------------------------------------------------ CPU reference
def _oracle_state(classes):
  import torch as t
  from corenet_b200 import configuration
  from corenet_b200.model.core_net import CoreNet
  t.manual_seed(0)
  m = CoreNet(configuration.default_config(classes))
  params = {k: v.detach().clone().requires_grad_(True) for k, v in m.named_parameters()}
  state = {k: v.clone() for k, v in m.state_dict().items()}
  state.update(params)
  return params, state


def _reference_cpu_model(classes):
  """(CoreNet, losses module, voxel_metrics module) of the UNMODIFIED reference from the staged copy
  baseline/_ref/src (baseline/stage_ref.py; git-ignored, travels to the GPU box), seeded like the oracle state, or
  None when no copy is staged (the oracle port, pinned bit-exact to the reference, is timed instead)."""
  import torch as t
  try:
    from baseline import ref_import
    if ref_import.import_reference() is None:
      return None
    from corenet import configuration as rc, voxel_metrics as rm
    from corenet.model import core_net, losses as rl
    cfg = rc.CoreNetConfig(decoder=rc.DecoderConfig(resolution=(128, 128, 128), num_output_channels=classes,
                                                    last_upscale_factor=2, latent_channels=64, skip_fraction=0.75))
    t.manual_seed(0)
    return core_net.CoreNet(cfg), rl, rm
  except Exception as e:                         # a broken staging must not take the arm down: fall back to the port
    print(f"bench: reference modules unavailable ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
    return None


def _cpu_reference(workload, steps, warmup, use_staged_reference=True):
  """The reference's own CPU implementation of the workload, all host threads, on a bounded sample (see `sample`).
  Model / losses / metrics: the UNMODIFIED reference modules from the staged copy baseline/_ref/src (kind
  "reference") when present, else the oracle port of them (kind "port", pinned bit-exact to the reference by
  oracle/make_golden.py); fill: the reference's fill_voxels_cpu.cc compiled in place (oracle/_ref, kind "reference")
  when present.  Returns (units/s, seconds/step, kind, sample text, host threads)."""
  import numpy as np
  import torch as t
  from oracle import corenet_oracle as O
  cores = os.cpu_count() or 1
  t.set_num_threads(cores)
  classes, loss_name, _, meshes, _ = WORKLOADS[workload]
  kind = "port"
  if workload == "fill":
    from oracle import build_ref, fill_voxels_oracle
    ref = build_ref.load()
    rng = np.random.default_rng(0)
    zz, yy, xx = np.meshgrid(*[np.arange(128)] * 3, indexing="ij")
    g = np.zeros((2, 128, 128, 128), np.float32)
    for i in range(2):
      d = np.sqrt((zz - 64) ** 2 + (yy - 60 - 4 * i) ** 2 + (xx - 64) ** 2)
      g[i] = ((d > 30) & (d < 32)) | (rng.random((128, 128, 128)) < 0.05)
    gt_ = t.from_numpy(g)
    if ref is not None:
      kind = "reference"
      fn = lambda: ref.fill_inside_voxels_cpu(gt_)
    else:
      fn = lambda: fill_voxels_oracle.fill_inside_voxels_oracle(g)
    units = 2 * VOX
    sample = "2 x 128^3 float32 grids per step (bounded sample of the 12-grid batch)"
  else:
    image, v2s, offsets, gt = synthetic_batch(1, 0, classes)
    gt = gt.to(t.int64)
    units = VOX
    ref_model = _reference_cpu_model(classes) if use_staged_reference else None
    if ref_model is not None:
      # the UNMODIFIED reference modules (staged copy baseline/_ref/src) on the host cores
      kind = "reference"
      model, ref_losses, ref_metrics = ref_model
    else:
      params, state = _oracle_state(classes)
    if workload == "h7":
      if ref_model is not None:
        model.eval()

        def fn():
          with t.no_grad():
            pmf = model(image, v2s, offsets).softmax(1)
            return ref_metrics.confusion_matrix(pmf.argmax(1), gt, classes)
      else:
        def fn():
          with t.no_grad():
            logits = O.corenet_forward(state, image, v2s, offsets, False)
            pmf = logits.softmax(1)
            return O.confusion_matrix(pmf.argmax(1), gt, classes)
      sample = "1 scene per step: eval forward + softmax + argmax + confusion (bounded sample of the 8-scene batch)"
    else:
      opt = t.optim.Adam(list(model.parameters() if ref_model is not None else params.values()), lr=4e-4, eps=1e-4)
      ref_fill = None
      if workload == "m9":
        from oracle import build_ref
        ref_fill = build_ref.load()
        zz, yy, xx = np.meshgrid(*[np.arange(128)] * 3, indexing="ij")
        shells = np.stack([((np.sqrt((zz - 64) ** 2 + (yy - 40 - 20 * i) ** 2 + (xx - 64) ** 2) > 18) &
                            (np.sqrt((zz - 64) ** 2 + (yy - 40 - 20 * i) ** 2 + (xx - 64) ** 2) < 20))
                           for i in range(3)]).astype(np.float32)
        shells_t = t.from_numpy(shells)

      if ref_model is not None:
        model.train()
        ref_loss = getattr(ref_losses, loss_name)

      def fn():
        if ref_fill is not None:                 # GT: the reference's CPU fill on the scene's 3 mesh grids
          ref_fill.fill_inside_voxels_cpu(shells_t)
        opt.zero_grad()
        if ref_model is not None:                # pipeline.py:224-230 of the reference
          loss = ref_loss(gt, model(image, v2s, offsets))
          loss.backward()
          opt.step()
          return
        nb = {}
        logits = O.corenet_forward(state, image, v2s, offsets, True, nb)
        loss = getattr(O, loss_name)(gt, logits)
        loss.backward()
        opt.step()
        for k, v in nb.items():
          state[k] = v
      sample = ("1 scene per step: train forward + loss + backward + Adam (bounded sample of the 4-scene batch)"
                + ("; GT = reference fill_inside_voxels_cpu on 3 pre-rasterised 128^3 shells (the GL rasteriser "
                   "has no CPU implementation)" if workload == "m9" and ref_fill is not None else ""))
  for _ in range(warmup):
    fn()
  t0 = time.perf_counter()
  for _ in range(steps):
    fn()
  dt = (time.perf_counter() - t0) / steps
  return units / dt, dt, kind, f"{steps} steps x " + sample, cores


def cpu_reference(workload, steps, warmup):
  """_cpu_reference with the staged reference modules, falling back to the oracle port if they fail under this torch."""
  try:
    return _cpu_reference(workload, steps, warmup, True)
  except Exception as e:
    print(f"bench: staged reference failed ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
    return _cpu_reference(workload, steps, warmup, False)


def run_reference(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  value, dt, kind, sample, cores = cpu_reference(args.workload, args.steps, args.warmup)
  b = args.batch or WORKLOADS[args.workload][2]
  line = {"impl": "reference", "metric": METRIC[args.workload], "value": value, "unit": "voxels/s",
          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt,
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "u8/f32" if args.workload == "fill" else "f32", "data": "synthetic",
          "config": {"workload": WORKLOADS[args.workload][4].format(b=b), "host_threads": cores},
          "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": cores, "kind": kind, "sample": sample},
          "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
  print(json.dumps(line))


def cpu_baseline(workload):
  value, dt, kind, sample, cores = cpu_reference(workload, 2, 1)
  return {"value": value, "unit": "voxels/s", "cores": cores, "kind": kind, "sample": sample}


# ------------------------------------------------------------------------------------------ reference on the GPU
def run_reference_gpu(args):
  """Informational same-box context: the UNMODIFIED reference modules (staged copy baseline/_ref) on stock
  PyTorch-CUDA (cuDNN), torch.optim.Adam, with cuDNN TF32 convolutions as torch ships them (allow_tf32 = True) and
  with fp32 convolutions; fill: the reference's own CUDA kernels.  Not the reference arm (that is the CPU path)."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  import torch as t
  from baseline import ref_import
  dev = t.device("cuda", 0)
  wl = args.workload
  classes, loss_name, bdef, meshes, desc = WORKLOADS[wl]
  b = args.batch or bdef
  out = {"impl": "reference-gpu", "metric": METRIC[wl], "unit": "voxels/s", "n_gpus": 1, "steps": args.steps,
         "warmup": args.warmup, "higher_is_better": True, "data": "synthetic",
         "config": {"workload": desc.format(b=b)}}

  def timed(fn):
    for _ in range(max(args.warmup, 3)):
      fn()
    t.cuda.synchronize()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
      fn()
    e1.record()
    t.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps

  if wl == "fill":
    mod = ref_import.reference_native_module()
    if mod is None:
      print(json.dumps({"impl": "reference-gpu", "unavailable": "baseline/_ref/corenet_cpp not staged"}))
      return
    grids = fill_grids(b, dev)
    ms = timed(lambda: mod.fill_inside_voxels_gpu(grids, False))
    out.update(value=b * VOX / (ms * 1e-3), ms_per_step=ms, dtype="f32",
               note="reference fill_voxels_gpu.cu K1/K2 compiled for sm_100a (baseline/stage_ref.py), clone + fill")
    print(json.dumps(out))
    return
  ref = ref_import.import_reference()
  if ref is None:
    print(json.dumps({"impl": "reference-gpu", "unavailable": "baseline/_ref/src not staged"}))
    return
  from corenet import configuration as rc
  from corenet.model import core_net, losses as rl
  cfg = rc.CoreNetConfig(decoder=rc.DecoderConfig(resolution=(128, 128, 128), num_output_channels=classes,
                                                  last_upscale_factor=2, latent_channels=64, skip_fraction=0.75))
  image, v2s, offsets, gt = [x.to(dev) for x in synthetic_batch(b, 0, classes)]
  gt = gt.to(t.int64)
  res = {}
  for tf32 in (True, False):
    t.backends.cudnn.allow_tf32 = tf32
    t.backends.cuda.matmul.allow_tf32 = False
    t.manual_seed(0)
    model = core_net.CoreNet(cfg).to(dev)
    if wl == "h7":
      model.eval()

      def fn():
        with t.no_grad():
          pmf = model(image, v2s, offsets).softmax(1)
          pred = pmf.argmax(1)
          idx = (gt * classes + pred).reshape(-1)
          return t.zeros(classes * classes, dtype=t.int32, device=dev).scatter_add(0, idx, t.ones_like(idx, dtype=t.int32))
    else:
      model.train()
      opt = t.optim.Adam(model.parameters(), lr=4e-4, eps=1e-4)
      loss_fn = getattr(rl, loss_name)

      def fn():
        opt.zero_grad()
        loss = loss_fn(gt, model(image, v2s, offsets))
        loss.backward()
        opt.step()
    ms = timed(fn)
    res["cudnn_tf32" if tf32 else "fp32"] = {"ms_per_step": ms, "value": b * VOX / (ms * 1e-3)}
    del model
    t.cuda.empty_cache()
  out.update(value=res["fp32"]["value"], ms_per_step=res["fp32"]["ms_per_step"], dtype="f32", variants=res,
             note="stock torch eager; m9's GL rasteriser cannot run here, so m9 is timed without its GT pipeline")
  print(json.dumps(out))


# ------------------------------------------------------------------------------------------ native arm
def run_native(args):
  import torch as t
  import torch.distributed as dist
  from corenet_b200 import _lib, configuration, engine
  from corenet_b200.model.core_net import CoreNet
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  t.cuda.set_device(local)
  dev = t.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  _lib.lib()      # raises if the CUDA library is missing: no fallback
  base_mode = args.precision
  engine.set_precision(base_mode)
  wl = args.workload
  classes, loss_name, bdef, meshes, desc = WORKLOADS[wl]
  b = args.batch or bdef
  steps, warmup = args.steps, max(args.warmup, 3)
  flush = t.empty(256 * 1024 * 1024, dtype=t.uint8, device=dev)   # > 126 MB L2

  def barrier():
    if world > 1:
      dist.barrier()
    t.cuda.synchronize()

  def timed(fn, n):
    """max-over-ranks device time of n calls of fn (CUDA events on the launch stream), L2 flushed in between."""
    evs = []
    barrier()
    for _ in range(n):
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

  pk, pk_src = peaks()
  sampler = ClockSampler(local)
  extra = {}
  model = None

  if wl == "fill":
    from corenet_b200 import ops
    grids = fill_grids(b, dev)
    host = grids.cpu().pin_memory()
    work = t.empty_like(grids)
    out_host = t.empty_like(host).pin_memory()
    ws = t.empty(_lib.lib().crn_fill_workspace_bytes(b, 128, 128, 128), dtype=t.uint8, device=dev)
    lib_call = _lib.call

    def dev_step():
      lib_call("crn_fill_inside", grids.data_ptr(), work.data_ptr(), 4, 1, b, 128, 128, 128, ws.data_ptr(),
               _lib.stream_ptr())

    def e2e_step():
      work.copy_(host, non_blocking=True)
      r = ops.fill_inside_voxels(work, inplace=True)         # the public op (cc.fill_voxels.fill_inside_voxels_gpu)
      out_host.copy_(r, non_blocking=False)
    h2d = d2h = host.numel() * 4
    for _ in range(warmup):
      dev_step()
    sampler.start()
    n0 = _lib.lib().crn_launch_count()
    ms = timed(dev_step, steps)
    launches = _lib.lib().crn_launch_count() - n0
    e2e_step()
    ms_e2e = timed(e2e_step, steps)
    clocks = sampler.stop()
    units = world * b * VOX
    achieved = 8.0 * b * VOX / (ms / steps * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "crn_fill_inside (pack + flood + unpack)", "achieved": achieved,
            "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": None,
            "peak_source": f"{pk_src} copy bandwidth",
            "note": "algorithmic bytes = read T + write T per voxel (8 B for float32), SURVEY 8d"}
    filled = int(work.sum().item())
    extra["check"] = {"voxels_set_after_fill": filled, "voxels_set_before": int(grids.sum().item())}
    mod = None
    try:
      from baseline import ref_import
      mod = ref_import.reference_native_module()
    except Exception:
      pass
    if mod is not None and rank == 0:
      ref_out = mod.fill_inside_voxels_gpu(grids, False)
      extra["check"]["bit_exact_vs_reference_gpu"] = bool(t.equal(ref_out, work))
      for _ in range(2):
        mod.fill_inside_voxels_gpu(grids, True)          # in place on an already filled grid = same traffic, no clone
      scratch = grids.clone()
      ms_ref = timed(lambda: mod.fill_inside_voxels_gpu(scratch.copy_(grids), True), steps)
      ms_copy = timed(lambda: scratch.copy_(grids), steps)
      ms_ref = max(ms_ref - ms_copy, 1e-6)
      extra["gpu_baseline"] = {
          "what": "the reference's own fill_voxels_gpu.cu (K1 union-find + K2 resolve, :96-132) compiled for sm_100a",
          "ms_per_step": ms_ref / steps, "value": b * VOX / (ms_ref / steps * 1e-3), "unit": "voxels/s",
          "achieved_gbs": 8.0 * b * VOX / (ms_ref / steps * 1e-3) / 1e9,
          "frac_of_hbm_peak": 8.0 * b * VOX / (ms_ref / steps * 1e-3) / 1e9 / pk["hbm_gbs"],
          "speedup_of_this_repo": (ms_ref / steps) / (ms / steps)}
  else:
    from corenet_b200.trainer import Trainer
    from corenet_b200.evaluator import Evaluator
    t.manual_seed(0)
    model = CoreNet(configuration.default_config(classes)).to(dev)
    image, v2s, offsets, gt = synthetic_batch(b, rank, classes)
    h_in = [x.pin_memory() for x in (image, v2s, offsets, gt)]
    d_in = [x.to(dev) for x in h_in]
    loss_host = t.empty(1, dtype=t.float32).pin_memory()
    if wl == "h7":
      model.eval()
      runner = Evaluator(model)
      cm_host = t.empty(classes, classes, dtype=t.int64).pin_memory()
      dev_step = lambda: runner.add_batch(*d_in)

      def e2e_step():
        runner.add_batch()                                # consumes the prefetched batch
        runner.prefetch(*h_in)                            # H2D of the next batch on the copy stream
        cm_host.copy_(runner.confusion_matrix, non_blocking=False)
      prime = lambda: runner.prefetch(*h_in)
      drain = lambda: runner.add_batch()
      h2d, d2h = sum(x.numel() * x.element_size() for x in h_in), classes * classes * 8
    else:
      model.train()
      runner = Trainer(model, lr=4e-4, eps=1e-4, loss=loss_name)
      h2d, d2h = sum(x.numel() * x.element_size() for x in h_in), 4
      if wl == "m9":
        from corenet_b200.data import batched_example as be
        tri, ntri, labels = synthetic_meshes(b, meshes, 100 + rank, classes)
        tri_h = tri.pin_memory()
        tri_d = tri.to(dev)
        content = be.VoxelContentSemanticLabel(labels)

        def gt_fn(src):
          # the reference's voxelize_batch (pipeline.py:126-150): rasterise (multiplier 8 -> 1024^2 samples), fill,
          # label * occupancy, max over the scene's meshes -> int32[B,128,128,128]; enqueued on the copy stream
          return lambda: be.voxelize(src, ntri, offsets, (128, 128, 128), content, image_resolution_multiplier=8,
                                     conservative_rasterization=False)[1]
        d_args = d_in[:3] + [gt_fn(tri_d)]
        h_args = h_in[:3] + [gt_fn(tri_h)]
        h2d = sum(x.numel() * x.element_size() for x in h_in[:3]) + tri_h.numel() * 4

        def dev_step():
          loss = runner.step()
          runner.prefetch(*d_args)                        # next step's GT pipeline overlaps this step
          return loss

        def e2e_step():
          loss = runner.step()
          runner.prefetch(*h_args)
          loss_host.copy_(loss, non_blocking=False)
        runner.prefetch(*d_args)
        prime = lambda: None
        drain = lambda: runner.step()
      else:
        dev_step = lambda: runner.step(*d_in)

        def e2e_step():
          # a training loop prefetches: the H2D copy of the NEXT step's inputs (pinned host memory -> staging buffers,
          # copy stream) is issued right after this step is enqueued and overlaps it; every timed step contains one
          # full H2D of a batch and the synchronous D2H read of its loss
          loss = runner.step()
          runner.prefetch(*h_in)
          loss_host.copy_(loss, non_blocking=False)
        prime = lambda: runner.prefetch(*h_in)
        drain = lambda: runner.step()
    for _ in range(warmup):
      dev_step()
    sampler.start()
    n0 = _lib.lib().crn_launch_count()
    ms = timed(dev_step, steps)
    # kernels of this library per step: counted at graph capture (replays do not pass through the C-ABI counter)
    launches = runner.graph_launches * steps + (_lib.lib().crn_launch_count() - n0)
    prime()
    for _ in range(2):
      e2e_step()
    ms_e2e = timed(e2e_step, steps)
    clocks = sampler.stop()
    drain()
    units = world * b * VOX
    # per-kernel roofline pass: the same step enqueued eagerly with CUDA events around every convolution launch and
    # every HBM-bound kernel (a replayed graph has no per-launch events); same inputs, same process, right after the
    # timed region; weight gradients on the main stream so that every kernel is timed alone
    engine.PROFILE = []
    engine.WGRAD_SIDE_STREAM = False
    prof_step = (lambda: runner.step(*d_in)) if wl != "h7" else dev_step
    if wl == "m9":
      gt_now = d_args[3]()
      prof_step = lambda: runner.step(*d_in[:3], gt_now)
    timed(prof_step, steps)
    prof, engine.PROFILE = engine.PROFILE, None
    engine.WGRAD_SIDE_STREAM = True
    if int(engine.get_engine(model).tc_status) != 0:
      raise RuntimeError("tcgen05 conv kernel reported a barrier timeout: results are invalid")
    if wl == "m9" and rank == 0:
      # the GT pipeline alone (not overlapped), for the record
      ms_gt = timed(lambda: d_args[3](), steps)
      extra["gt_pipeline_ms"] = ms_gt / steps
    if world == 1 and wl in ("h5", "m7", "h7") and base_mode == "3xtf32" and not args.no_modes:
      # opt-in single-pass TF32 arithmetic (engine.set_precision("tf32"), one MMA per product instead of three):
      # reported beside the headline, never as the headline -- the parity bar (logits <= 1e-3) needs "3xtf32"
      engine.set_precision("tf32")
      try:
        for _ in range(warmup):
          dev_step()
        ms_fast = timed(dev_step, steps)
      finally:
        engine.set_precision("3xtf32")
      extra["precision_modes"] = {"tf32": {
          "ms_per_step": ms_fast / steps, "value": units * steps / (ms_fast * 1e-3), "unit": "voxels/s",
          "note": "non-default: hi x hi product only; eval logits within 1e-2 of the oracle instead of 1e-3 "
                  "(tests/test_gpu_configs.py, profiles/r02_precision_study.md)"}}
    roof = build_roofline(prof, steps, pk, pk_src, args.layers) if rank == 0 else None

  if rank == 0:
    value = units * steps / (ms * 1e-3)
    e2e_value = units * steps / (ms_e2e * 1e-3)
    line = {"metric": METRIC[wl], "value": value, "unit": "voxels/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if base_mode == "3xtf32" else "tf32 (diagnostic, non-default mode)",
            "data": "synthetic",
            "config": {"workload": desc.format(b=b), "name": wl, "global_batch": world * b, "precision": base_mode,
                       "parallelism": f"dp{world}", "l2": "256 MiB buffer written between timed iterations",
                       "cuda_graph": bool(getattr(runner, "graph_launches", 0)) if wl != "fill" else False,
                       "scenes_per_sec": value / VOX if wl != "fill" else None},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "voxels/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / steps,
                    "note": "public API with pinned HOST inputs: every timed step issues the full H2D of a batch "
                            "(copy stream, overlapping the running step like a prefetching data loader) and "
                            "synchronously reads the step's result back"},
            "gpu_launches": int(launches), "roofline": roof}
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
      line["cpu_baseline"] = cpu_baseline(wl)
    print(json.dumps(line))
  if world > 1:
    # the captured step holds NCCL kernels: destroy_process_group() blocks while such graphs are alive (observed on
    # torch 2.11 / NCCL 2.28), so every rank synchronises and leaves without tearing the communicator down
    dist.barrier()
    t.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def build_roofline(prof, steps, pk, pk_src, layers_path):
  """roofline object of the dominant conv family (tensor) + an `hbm` table of the HBM-bound kernels of the step."""
  fam, hbm = {}, {}
  for kind, name, work, e0, e1 in prof:
    if kind == "hbm":
      f = hbm.setdefault(name, [0.0, 0.0, 0])
      f[0] += e0.elapsed_time(e1); f[1] += work; f[2] += 1
      continue
    # one family per kernel function (source file in csrc/)
    k = {"wgrad": "wgrad_kernels(ffma)", "wgrad_tc": "wgrad_tc_kernel(tcgen05, per-tap: conv_wgrad_tc.cu)",
         "wgrad_line": "wgrad_line/tline/xline_kernel(tcgen05, tap-stacked: conv_wgrad_line.cu)",
         "fwd_gt": "gemm_tc_kernel(tcgen05: conv_gemm_tc.cu)", "dgrad_gt": "gemm_tc_kernel(tcgen05: conv_gemm_tc.cu)",
         "fwd_tc": "conv_tc5_kernel(tcgen05: conv_tc5.cu)", "dgrad_tc": "conv_tc5_kernel(tcgen05: conv_tc5.cu)",
         "fwd_tcs": "conv_tc5s_kernel(tcgen05, z-taps stacked: conv_tc5s.cu)",
         "dgrad_tcs": "conv_tc5s_kernel(tcgen05, z-taps stacked: conv_tc5s.cu)"}.get(kind, "conv_fwd_dgrad_kernels(ffma)")
    f = fam.setdefault(k, [0.0, 0.0, 0])
    f[0] += e0.elapsed_time(e1); f[1] += 2.0 * work; f[2] += 1
  if layers_path:
    per = {}
    for kind, name, work, e0, e1 in prof:
      if kind == "hbm":
        continue
      q = per.setdefault((name, kind), [0.0, 0.0])
      q[0] += e0.elapsed_time(e1) / steps; q[1] += 2.0 * work / steps
    rows_ = sorted(((v[0], k[0], k[1], v[1] / (v[0] * 1e-3) / 1e12) for k, v in per.items()), reverse=True)
    with open(layers_path, "w") as f:
      for ms_, n_, k_, tf_ in rows_:
        f.write(f"{ms_:9.3f} ms  {tf_:7.2f} TFLOP/s  {k_:6s} {n_}\n")
      for n_, v in sorted(hbm.items(), key=lambda kv: -kv[1][0]):
        f.write(f"{v[0] / steps:9.3f} ms  {v[1] / (v[0] * 1e-3) / 1e9:7.1f} GB/s     hbm    {n_} ({v[2] // steps} launches)\n")
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
  hbm_peak = pk["hbm_gbs"]
  all_flops = sum(v[1] for v in fam.values())
  return {"bound": "tensor", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
          "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
          "peak_source": f"{pk_src} bf16 sustained",
          "launches": dom[1][2], "kernel_ms_per_step": dom[1][0] / steps,
          "all_conv_ms_per_step": conv_ms / steps,
          "all_conv_tflops": all_flops / (conv_ms * 1e-3) / 1e12, "all_conv_frac": all_flops / (conv_ms * 1e-3) / 1e12 / peak,
          "conv_gflop_per_step": all_flops / steps / 1e9,
          # continuity with round 1, whose dominant family was the k5 / transposed-k7 forward+dgrad kernels
          # (conv_tc5.cu + conv_tc5s.cu together; 6.84 ms at 75.1 TFLOP/s then)
          "tracked_conv_tc5_family": (lambda a, b: {"ms_per_step": (a[0] + b[0]) / steps,
                                                   "tflops": (a[1] + b[1]) / ((a[0] + b[0]) * 1e-3) / 1e12,
                                                   "frac": (a[1] + b[1]) / ((a[0] + b[0]) * 1e-3) / 1e12 / peak})(
              fam.get("conv_tc5_kernel(tcgen05: conv_tc5.cu)", [1e-9, 0.0, 0]),
              fam.get("conv_tc5s_kernel(tcgen05, z-taps stacked: conv_tc5s.cu)", [1e-9, 0.0, 0])),
          "families": {k: {"ms_per_step": v[0] / steps, "tflops": v[1] / (v[0] * 1e-3) / 1e12}
                       for k, v in fam.items()},
          "hbm": {k: {"ms_per_step": v[0] / steps, "launches_per_step": v[2] // steps,
                      "achieved_gbs": v[1] / (v[0] * 1e-3) / 1e9, "frac": v[1] / (v[0] * 1e-3) / 1e9 / hbm_peak}
                  for k, v in sorted(hbm.items())},
          "hbm_peak_gbs": hbm_peak,
          "note": "algorithmic FLOPs = 2*MACs of each conv launch / CUDA-event time of that launch, from an eager "
                  "pass of the same step right after the timed region (the timed region replays a CUDA graph); "
                  "tcgen05 kernels run 3xTF32 (3 MMAs per product) for fp32-class accuracy; `hbm` = algorithmic bytes "
                  "of each HBM-bound kernel / its CUDA-event time in the same pass"}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--workload", default="h5", choices=sorted(WORKLOADS))
  ap.add_argument("--batch", type=int, default=0, help="scenes (grids for fill) per GPU; 0 = the workload's default")
  ap.add_argument("--impl", default="native", choices=["native", "reference", "reference-gpu"])
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--precision", default="3xtf32", choices=["3xtf32", "tf32"],
                  help="diagnostic: run the WHOLE bench in the opt-in single-pass TF32 mode (labelled in dtype/config)")
  ap.add_argument("--no-modes", action="store_true", help="skip the extra single-pass-TF32 measurement (precision_modes)")
  ap.add_argument("--layers", default=None, help="write a per-layer conv / HBM-kernel timing table to this file")
  args = ap.parse_args()
  if os.environ.get("CRN_FAULT_TIMEOUT"):        # debugging aid: dump all Python stacks and exit if the run stalls
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ["CRN_FAULT_TIMEOUT"]), exit=True)
  if args.impl == "reference":
    run_reference(args)
  elif args.impl == "reference-gpu":
    run_reference_gpu(args)
  else:
    run_native(args)


if __name__ == "__main__":
  main()
