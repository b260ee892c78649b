"""Measures how much rounding the convolution operands moves the CoReNet logits (DESIGN.md "Precision").

Runs the oracle (bit-exact restatement of the reference, torch CPU fp32) with every conv / linear operand pair
(activation, weight) rounded the way a tensor-core mode would see it, products accumulated in fp32:

  fp32     reference arithmetic
  tf32     a -> rna_tf32(a)                                        1 MMA / product   (what cuDNN's allow_tf32 does)
  bf16x2   a = hi + lo (bf16 each): hi*hi + hi*lo + lo*hi          3 MMAs / product
  3xtf32   a = hi + lo (tf32 hi, fp32 remainder): same 3 products  3 MMAs / product  (this repo's tcgen05 kernels)

in three states: train mode at the seeded init (batch statistics), eval mode at init (running stats 0/1), and eval
mode with running statistics set to the batch statistics of the input ("sane" statistics, a trained net's regime).
Prints a markdown table: max|logits - fp32| / max|fp32| and the fraction of voxels whose argmax flips.

    python scripts/precision_study.py > profiles/r02_precision_study.md
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch as t  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import corenet_oracle as O  # noqa: E402
from oracle import make_golden as MG  # noqa: E402


def rna_tf32(x):
  """round-to-nearest (ties away) to 10 explicit mantissa bits, like cvt.rna.tf32.f32."""
  i = x.contiguous().view(t.int32)
  return ((i + 0x1000) & ~0x1FFF).view(t.float32)


def split(x, mode):
  if mode == "fp32":
    return [(x, 1)]
  if mode == "tf32":
    return [(rna_tf32(x), 1)]
  hi = x.to(t.bfloat16).to(t.float32) if mode == "bf16x2" else rna_tf32(x)
  lo = x - hi
  if mode == "bf16x2":
    lo = lo.to(t.bfloat16).to(t.float32)
  return [(hi, 1), (lo, 0)]


def wrap(fn, mode):
  def f(x, w, b=None, *a, **k):
    xs, ws = split(x, mode), split(w, mode)
    out = None
    for xv, xh in xs:
      for wv, wh in ws:
        if not (xh or wh):
          continue                       # lo * lo is dropped
        y = fn(xv, wv, None, *a, **k)
        out = y if out is None else out + y
    if b is not None:
      out = out + b.reshape([1, -1] + [1] * (out.dim() - 2))
    return out
  return f


def run(state, inp, training, mode, nb=None):
  orig = {n: getattr(F, n) for n in ("conv2d", "conv3d", "conv_transpose3d", "linear")}
  try:
    for n, fn in orig.items():
      setattr(F, n, wrap(fn, mode))
    with t.no_grad():
      return O.corenet_forward(dict(state), inp["image"], inp["v2s"], inp["offsets"], training, nb)
  finally:
    for n, fn in orig.items():
      setattr(F, n, fn)


def main():
  from corenet_b200 import configuration
  from corenet_b200.model.core_net import CoreNet
  t.manual_seed(0)
  sd = {k: v.clone() for k, v in CoreNet(configuration.default_config(2)).state_dict().items()}
  inp = MG.case_inputs("B")
  # "sane" running statistics: the batch statistics of this input (undo the momentum-0.01 update)
  nb = {}
  run(sd, inp, True, "fp32", nb)
  sane = dict(sd)
  for k, v in nb.items():
    if k.endswith("running_mean") or k.endswith("running_var"):
      sane[k] = (v - 0.99 * sd[k]) / 0.01
  states = (("train mode @ seeded init (batch statistics)", sd, True),
            ("eval mode @ init (running stats 0 / 1)", sd, False),
            ("eval mode, running stats = batch statistics of the input", sane, False))
  print("# Operand-rounding study (scripts/precision_study.py; oracle = reference arithmetic, CPU fp32, case B inputs)\n")
  print("| state | mode | MMAs / product | max abs diff of logits / max abs logits (fp32) | argmax flips |")
  print("|---|---|---|---|---|")
  for name, state, training in states:
    ref = run(state, inp, training, "fp32")
    for mode, n in (("tf32", 1), ("bf16x2", 3), ("3xtf32", 3)):
      got = run(state, inp, training, mode)
      err = ((got - ref).abs().max() / ref.abs().max()).item()
      flips = (got.argmax(1) != ref.argmax(1)).float().mean().item()
      print(f"| {name} | {mode} | {n} | {err:.3e} | {flips:.3%} |", flush=True)


if __name__ == "__main__":
  main()
