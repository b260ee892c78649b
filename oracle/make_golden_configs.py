"""Golden fixtures for BASELINE.json's own configurations, generated from the REAL reference.

Run in the build container only (needs /root/reference):
    python oracle/make_golden_configs.py
Writes tests/golden/corenet_reference_configs.npz:

  C  h5  per-GPU batch: B=4, C=2, train mode, iou_fgbg      -> logits / decoder taps / loss / gradients / running stats
  D  h7  B=8, C=2, eval mode                                 -> logits / taps / confusion matrix of argmax vs the GT
  E  m7/m9  B=2, C=15, train mode, xent_times_iou_agnostic   -> logits / taps / loss / gradients

Logits, taps and fp32 gradients come from the unmodified reference modules (imported from /root/reference/src);
the fp64 gradients ("exact" answer, used to measure how accurate a gradient is when the reference's own fp32
gradients are several % off at random init) come from oracle/corenet_oracle.py, which make_golden.py pins bit-exact
to the reference.  Values are stored at seeded indices; the tests replay them on the GPU box without the reference.
"""
import os
import sys

import numpy as np
import torch as t

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_golden as MG  # noqa: E402

LOSS = {"C": "iou_fgbg", "D": None, "E": "xent_times_iou_agnostic"}


def config_inputs(case: str):
  """Seeded inputs of the three configurations (shared with the tests)."""
  from oracle import corenet_oracle as O
  b, classes, training = {"C": (4, 2, True), "D": (8, 2, False), "E": (2, 15, True)}[case]
  g = t.Generator().manual_seed({"C": 10, "D": 11, "E": 12}[case])
  image = t.randint(0, 256, (b, 3, 256, 256), dtype=t.uint8, generator=g)
  offsets = t.rand(b, 3, generator=g)
  v2s = O.default_v2s(b).clone()
  if case == "D":        # one scene with a shifted camera: voxels outside the image / behind the camera
    v2s[3] = O.dataset_camera() @ O.translate([0.35, -0.2, -1.1]) @ O.scale([128.0] * 3).inverse()
  gt = MG.synthetic_gt(b, classes, seed=20 + ord(case))
  return dict(image=image, v2s=v2s, offsets=offsets, classes=classes, training=training, gt=gt)


def grad_samples(g: t.Tensor, n=128, seed=77):
  g = g.detach().reshape(-1)
  gen = t.Generator().manual_seed(seed + g.numel() % 9973)
  return t.randint(0, g.numel(), (min(n, g.numel()),), generator=gen)


def main():
  from oracle import corenet_oracle as O
  rc, rnet, rloss = MG.import_reference()
  out = {}
  for case in ("C", "D", "E"):
    inp = config_inputs(case)
    cfg = rc.CoreNetConfig(decoder=rc.DecoderConfig(
        resolution=(128, 128, 128), num_output_channels=inp["classes"], last_upscale_factor=2, latent_channels=64,
        skip_fraction=0.75))
    t.manual_seed(0)
    ref = rnet.CoreNet(cfg)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    ref.train(inp["training"])
    logits = ref(inp["image"], inp["v2s"], inp["offsets"])
    pre = case + "."
    for kk, vv in MG.summarize(logits).items():
      out[pre + "logits." + kk] = np.asarray(vv)
    # decoder / encoder taps through the (pinned) oracle on the same state
    st = {k: v.clone() for k, v in sd.items()}
    nb, taps = {}, {}
    with t.no_grad():
      lo = O.corenet_forward(st, inp["image"], inp["v2s"], inp["offsets"], inp["training"], nb, taps)
    assert (lo - logits).abs().max().item() <= 1e-5 * logits.abs().max().item(), case
    for k, v in taps.items():
      for kk, vv in MG.summarize(v).items():
        out[pre + "tap." + k + "." + kk] = np.asarray(vv)
    if LOSS[case] is None:
      cm = O.confusion_matrix(logits.argmax(1), inp["gt"], inp["classes"])
      out[pre + "cm"] = cm.numpy()
      out[pre + "miou"] = np.array(O.mean_iou(cm))
      print(f"case {case}: eval B={logits.shape[0]}  mIoU {O.mean_iou(cm):.6f}")
      continue
    loss = getattr(rloss, LOSS[case])(inp["gt"], logits)
    loss.backward()
    out[pre + "loss"] = np.array(loss.item())
    # fp64 oracle gradients
    st64 = {k: (v.clone().double() if v.dtype == t.float32 else v.clone()) for k, v in sd.items()}
    for k, v in st64.items():
      if v.dtype == t.float64 and "running" not in k:
        v.requires_grad_(True)
    l64 = getattr(O, LOSS[case])(inp["gt"], O.corenet_forward(
        st64, inp["image"], inp["v2s"].double(), inp["offsets"].double(), True, {}, None, dtype=t.float64))
    l64.backward()
    out[pre + "loss64"] = np.array(l64.item())
    for n_, p in ref.named_parameters():
      idx = grad_samples(p.grad)
      g64 = st64[n_].grad.reshape(-1)
      out[pre + "grad." + n_ + ".idx"] = idx.numpy()
      out[pre + "grad." + n_ + ".ref32"] = p.grad.reshape(-1)[idx].double().numpy()
      out[pre + "grad." + n_ + ".f64"] = g64[idx].numpy()
      out[pre + "grad." + n_ + ".norm64"] = np.array(g64.norm().item())
    for k, v in ref.state_dict().items():
      if "running" in k:
        out[pre + "buf." + k + ".abssum"] = np.array(v.double().abs().sum().item())
        out[pre + "buf." + k + ".max"] = np.array(v.double().abs().max().item())
        idx = grad_samples(v, 16)
        out[pre + "buf." + k + ".idx"] = idx.numpy()
        out[pre + "buf." + k + ".val"] = v.reshape(-1)[idx].double().numpy()
    print(f"case {case}: train B={logits.shape[0]} C={inp['classes']} loss {loss.item():.6f} (fp64 {l64.item():.6f})")
  path = os.path.join(ROOT, "tests", "golden", "corenet_reference_configs.npz")
  np.savez_compressed(path, **out)
  print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
  main()
