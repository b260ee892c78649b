"""One eager training step (h5 shapes, B=4) + one evaluation batch + one GT pipeline call with cudaProfilerStart/Stop
around a chosen set of launches, to be run under

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/<name> \\
      python scripts/ncu_capture.py [--classes 2]

so that exactly one launch of every kernel family named in VERDICT r1 item 2 is captured (conv_tc5s fwd/dgrad,
conv_tc5 (transposed k7), gemm_tc, wgrad_tc, wgrad_line, the four skip_fwd scales, skip_bwd_sorted, BatchRenorm
stats/apply/bwd, loss_sums/loss_bwd, Adam, fill pack/flood/unpack, voxelize).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch as t  # noqa: E402

import bench  # noqa: E402
from corenet_b200 import configuration, engine  # noqa: E402
from corenet_b200.model.core_net import CoreNet  # noqa: E402
from corenet_b200.trainer import Trainer  # noqa: E402

CONV = {("fwd_tcs", "decoder.stage_6.c1"), ("dgrad_tcs", "decoder.stage_6.c1"), ("wgrad_line", "decoder.stage_6.c1"),
        ("fwd_tc", "decoder.stage_5.t1"), ("dgrad_tcs", "decoder.stage_5.t1"), ("fwd_tcs", "decoder.stage_6.t1"),
        ("fwd_gt", "decoder.stage_4.c1"), ("fwd_gt", "encoder.stage3.b.op_b.conv"),
        ("dgrad_gt", "encoder.stage4.b.op_c.conv"), ("wgrad_tc", "encoder.stage4.b.op_b.conv"),
        ("wgrad_line", "decoder.stage_4.c1"), ("dgrad", "decoder.stage_6.t1"), ("fwd", "encoder.stage1.conv")}
CALLS_ONCE = {"crn_loss_sums", "crn_loss_bwd", "crn_adam_step_guarded", "crn_unpack_wgrads",
              "crn_softmax_planar", "crn_argmax_confusion_labeled"}
CALLS_ALL = {"crn_skip_sample_fwd", "crn_skip_sample_bwd_sorted", "crn_skip_build_lists"}
BRN_ONCE = {"crn_brn_stats", "crn_brn_apply", "crn_brn_bwd_reduce", "crn_brn_bwd_dx"}
seen = set()
state = {"armed": False, "brn_rows": None}


def pick(kind, name):
  if not state["armed"]:
    return False
  if kind == "call":
    if name in CALLS_ALL:
      return True
    if name in CALLS_ONCE and name not in seen:
      seen.add(name)
      return True
    return False
  if (kind, name) in CONV and (kind, name) not in seen:
    seen.add((kind, name))
    return True
  return False


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--classes", type=int, default=2)
  ap.add_argument("--batch", type=int, default=4)
  ap.add_argument("--top", action="store_true", help="only the dominant tcgen05 launches of the training step")
  args = ap.parse_args()
  if args.top:
    global CONV, CALLS_ONCE, CALLS_ALL, BRN_ONCE
    CONV = {("fwd_tcs", "decoder.stage_6.c1"), ("dgrad_tcs", "decoder.stage_6.c1"), ("wgrad_line", "decoder.stage_6.c1"),
            ("fwd_tcs", "decoder.stage_5.c1"), ("wgrad_line", "decoder.stage_5.c1"), ("dgrad_tcs", "decoder.stage_5.t1"),
            ("wgrad_line", "decoder.stage_5.t1"), ("fwd_gt", "encoder.stage3.b.op_b.conv"),
            ("wgrad_tc", "encoder.stage4.b.op_b.conv")}
    CALLS_ONCE, CALLS_ALL, BRN_ONCE = set(), set(), set()
  dev = t.device("cuda", 0)
  t.manual_seed(0)
  model = CoreNet(configuration.default_config(args.classes)).to(dev).train()
  loss = "iou_fgbg" if args.classes == 2 else "xent_times_iou_agnostic"
  tr = Trainer(model, loss=loss, use_graph=False)
  d_in = [x.to(dev) for x in bench.synthetic_batch(args.batch, 0, args.classes)]
  engine.WGRAD_SIDE_STREAM = False
  for _ in range(2):
    tr.step(*d_in)
  t.cuda.synchronize()
  # BatchRenorm: the stage_6 instances (largest) -- bracket by row count
  big = 4 * 64 ** 3
  orig_call = engine._lib.call

  def brn_pick(kind, name):
    return pick(kind, name)
  engine.NCU_PICK = brn_pick
  real = engine._lib.call

  def call_hook(fn, *a):
    if state["armed"] and fn in BRN_ONCE and fn not in seen:
      rows = a[1] if fn in ("crn_brn_stats", "crn_brn_apply") else (a[8] if fn == "crn_brn_bwd_reduce" else a[6])
      if rows == args.batch * 64 ** 3:
        seen.add(fn)
        t.cuda.cudart().cudaProfilerStart()
        try:
          return real(fn, *a)
        finally:
          t.cuda.cudart().cudaProfilerStop()
    return real(fn, *a)
  engine._lib.call = call_hook
  state["armed"] = True
  tr.step(*d_in)
  t.cuda.synchronize()
  if args.top:
    print("captured:", sorted(seen), "not seen:", [c for c in CONV if c not in seen])
    return
  # evaluation kernels (softmax, argmax/confusion)
  from corenet_b200.evaluator import Evaluator
  model.eval()
  ev = Evaluator(model, use_graph=False)
  ev.add_batch(*d_in)
  t.cuda.synchronize()
  state["armed"] = False
  engine._lib.call = real
  engine.NCU_PICK = None
  # GT pipeline: voxelise + fill + merge of 12 meshes (one m9 batch)
  from corenet_b200.data import batched_example as be
  tri, ntri, labels = bench.synthetic_meshes(args.batch, 3, 100, 15)
  offs = t.full((args.batch, 3), 0.5)
  run = lambda: be.voxelize(tri.to(dev), ntri, offs, (128, 128, 128), be.VoxelContentSemanticLabel(labels),
                            image_resolution_multiplier=8, conservative_rasterization=False)
  run()
  t.cuda.synchronize()
  t.cuda.cudart().cudaProfilerStart()
  run()
  t.cuda.synchronize()
  t.cuda.cudart().cudaProfilerStop()
  missing = [c for c in CONV if c not in seen] + [c for c in CALLS_ONCE | BRN_ONCE if c not in seen]
  print("captured:", len(seen), "launch families; not seen:", missing)


if __name__ == "__main__":
  main()
