"""Pins oracle/corenet_oracle.py against the REAL reference and writes tests/golden/*.npz.

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py
It imports the unmodified reference modules (with a stub for the absent
`dataclasses_jsonschema` package), runs them on seeded inputs, asserts that the
oracle restatement reproduces them, and stores compact fixtures (checksums +
sampled values of every pinned tensor) that travel to the GPU box.
"""
import dataclasses
import os
import sys
import types

import numpy as np
import torch as t

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/src"


def import_reference():
  m = types.ModuleType("dataclasses_jsonschema")

  class JsonSchemaMixin:
    def to_dict(self):
      return dataclasses.asdict(self)
  m.JsonSchemaMixin = JsonSchemaMixin
  sys.modules["dataclasses_jsonschema"] = m
  sys.path.insert(0, REF)
  from corenet import configuration as rc
  from corenet.model import core_net as rnet
  from corenet.model import losses as rloss
  return rc, rnet, rloss


def summarize(x: t.Tensor, n_samples=64, seed=1234):
  """Compact fingerprint of a tensor: shape, sums, and values at seeded indices."""
  x = x.detach().to(t.float64).reshape(-1)
  g = t.Generator().manual_seed(seed + x.numel() % 9973)
  idx = t.randint(0, x.numel(), (min(n_samples, x.numel()),), generator=g)
  return dict(sum=x.sum().item(), abssum=x.abs().sum().item(), max=x.abs().max().item(),
              idx=idx.numpy(), val=x[idx].numpy())


def synthetic_gt(batch, classes, seed=3, res=128):
  """Union of 1-3 random boxes / spheres per scene, labels 1..classes-1."""
  g = t.Generator().manual_seed(seed)
  zz, yy, xx = t.meshgrid([t.arange(res)] * 3, indexing="ij")
  out = t.zeros(batch, res, res, res, dtype=t.int64)
  for b in range(batch):
    for _ in range(int(t.randint(1, 4, (1,), generator=g))):
      c = t.randint(res // 4, 3 * res // 4, (3,), generator=g)
      r = t.randint(res // 10, res // 4, (3,), generator=g)
      lab = int(t.randint(1, classes, (1,), generator=g))
      if int(t.randint(0, 2, (1,), generator=g)):
        m = ((zz - c[0]).abs() <= r[0]) & ((yy - c[1]).abs() <= r[1]) & ((xx - c[2]).abs() <= r[2])
      else:
        m = ((zz - c[0]) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2) <= int(r[0]) ** 2
      out[b][m] = lab
  return out


def case_inputs(case: str):
  """Seeded inputs shared by make_golden.py and the tests."""
  from oracle import corenet_oracle as O
  g0 = t.Generator().manual_seed(0)
  if case == "A":       # dataset camera, centre offsets, fresh BRN state, B=1
    b = 1
    image = t.randint(0, 256, (b, 3, 256, 256), dtype=t.uint8, generator=g0)
    return dict(image=image, v2s=O.default_v2s(b), offsets=t.full((b, 3), 0.5), classes=2, perturb=False)
  if case == "B":       # random offsets, one shifted camera (voxels outside / behind), active r/d clamps
    b = 2
    image = t.randint(0, 256, (b, 3, 256, 256), dtype=t.uint8, generator=g0)
    offsets = t.rand(b, 3, generator=t.Generator().manual_seed(1))
    v2s = O.default_v2s(b).clone()
    shift = O.translate([0.35, -0.2, -1.1])       # pushes ~1/3 of the cube outside / behind the camera
    v2s[1] = O.dataset_camera() @ shift @ O.scale([128.0] * 3).inverse()
    return dict(image=image, v2s=v2s, offsets=offsets, classes=2, perturb=True)
  raise KeyError(case)


def perturb_brn(sd, seed=2):
  """num_batches_tracked=50000 and perturbed running stats so r,d clamps are active."""
  g = t.Generator().manual_seed(seed)
  for k in sd:
    if k.endswith("num_batches_tracked"):
      sd[k] = t.tensor(50000, dtype=t.int64)
    elif k.endswith("running_mean"):
      sd[k] = sd[k] + 0.3 * t.randn(sd[k].shape, generator=g)
    elif k.endswith("running_var"):
      sd[k] = sd[k] * (0.25 + 1.5 * t.rand(sd[k].shape, generator=g))
  return sd


def main():
  from oracle import corenet_oracle as O
  from corenet_b200 import configuration as C
  from corenet_b200.model.core_net import CoreNet
  rc, rnet, rloss = import_reference()
  out_dir = os.path.join(ROOT, "tests", "golden")
  os.makedirs(out_dir, exist_ok=True)

  cfg_ref = rc.CoreNetConfig(decoder=rc.DecoderConfig(
      resolution=(128, 128, 128), num_output_channels=2, last_upscale_factor=2, latent_channels=64,
      skip_fraction=0.75))
  t.manual_seed(0)
  ref = rnet.CoreNet(cfg_ref)
  t.manual_seed(0)
  mine = CoreNet(C.default_config(2))
  sd_ref, sd_mine = ref.state_dict(), mine.state_dict()
  assert list(sd_ref.keys()) == list(sd_mine.keys()), "state_dict keys/order differ"
  for k in sd_ref:
    assert sd_ref[k].shape == sd_mine[k].shape and sd_ref[k].dtype == sd_mine[k].dtype, k
    assert t.equal(sd_ref[k], sd_mine[k]), f"init differs: {k}"
  print(f"state_dict: {len(sd_ref)} keys identical (names, shapes, dtypes, seeded values)")
  golden = {"param_names": np.array(list(sd_ref.keys()))}
  golden["param_abssum"] = np.array([sd_ref[k].double().abs().sum().item() for k in sd_ref])

  for case in ("A", "B"):
    inp = case_inputs(case)
    sd = {k: v.clone() for k, v in sd_ref.items()}
    if inp["perturb"]:
      sd = perturb_brn(sd)
    gt = synthetic_gt(inp["image"].shape[0], inp["classes"])
    for mode in ("train", "eval"):
      ref.load_state_dict(sd)
      ref.train(mode == "train")
      for p in ref.parameters():
        p.grad = None
      logits_ref = ref(inp["image"], inp["v2s"], inp["offsets"])
      loss_ref = rloss.iou_fgbg(gt, logits_ref)
      loss_ref.backward()
      # ---- the oracle restatement on the same state
      st = {k: v.clone().requires_grad_(v.dtype == t.float32 and not k.split(".")[-1].startswith(("running", "num")))
            for k, v in sd.items()}
      nb, taps = {}, {}
      logits_o = O.corenet_forward(st, inp["image"], inp["v2s"], inp["offsets"], mode == "train", nb, taps)
      loss_o = O.iou_fgbg(gt, logits_o)
      loss_o.backward()
      dl = (logits_o - logits_ref).abs().max().item()
      assert dl <= 1e-5 * logits_ref.abs().max().item(), (case, mode, dl)
      assert abs(loss_o.item() - loss_ref.item()) < 1e-6
      worst = 0.0
      for n_, p in ref.named_parameters():
        go = st[n_].grad
        d = (go - p.grad).abs().max().item()
        worst = max(worst, d / (p.grad.abs().max().item() + 1e-12))
      assert worst < 1e-3, (case, mode, worst)
      if mode == "train":
        for k, v in ref.state_dict().items():
          if k in nb:
            assert t.allclose(v.to(t.float64), nb[k].to(t.float64), rtol=1e-6, atol=1e-7), k
      print(f"case {case}/{mode}: oracle == reference (logits max|d|={dl:.2e}, worst rel grad diff={worst:.2e}, "
            f"loss={loss_ref.item():.6f})")
      pre = f"{case}.{mode}."
      golden[pre + "loss"] = np.array(loss_ref.item())
      for name, ten in [("logits", logits_ref)] + [("tap." + k, v) for k, v in taps.items()]:
        for kk, vv in summarize(ten).items():
          golden[pre + name + "." + kk] = np.asarray(vv)
      for n_, p in ref.named_parameters():
        for kk, vv in summarize(p.grad, 16).items():
          golden[pre + "grad." + n_ + "." + kk] = np.asarray(vv)
      if mode == "train":
        for k, v in ref.state_dict().items():
          if "running" in k:
            golden[pre + "buf." + k + ".abssum"] = np.array(v.double().abs().sum().item())
  np.savez_compressed(os.path.join(out_dir, "corenet_reference.npz"), **golden)
  print("wrote", os.path.join(out_dir, "corenet_reference.npz"))


if __name__ == "__main__":
  main()
