"""Pins the loss restatements with the reference's known answers (src/corenet/test/losses_test.py:25-88)."""
import math

import numpy as np
import torch as t

from oracle import corenet_oracle as O

LOGITS = t.as_tensor([
    [[[[0.8278376, 0.44923675, 0.9302666, 0.6919297], [0.38287663, 0.37834585, 0.051413298, 0.7789054]],
      [[0.71893823, 0.94472325, 0.35577738, 0.0018994808], [0.41523135, 0.7561617, 0.0044674873, 0.38063014]],
      [[0.3408773, 0.22092032, 0.0767951, 0.17644858], [0.3457942, 0.27810383, 0.74627364, 0.43618906]]],
     [[[0.70214736, 0.54277015, 0.4549327, 0.79017854], [0.4176488, 0.22357666, 0.43264854, 0.29656994]],
      [[0.15031266, 0.8952414, 0.011986375, 0.26919663], [0.084516525, 0.043944597, 0.6917249, 0.5230026]],
      [[0.42145348, 0.28770554, 0.50909555, 0.48172605], [0.97358274, 0.8910786, 0.5946312, 0.51896834]]]],
    [[[[0.9724301, 0.41606557, 0.1918621, 0.1327486], [0.6457069, 0.76746213, 0.022811055, 0.8097471]],
      [[0.44591904, 0.51651776, 0.89206624, 0.98763657], [0.75536454, 0.20767283, 0.01293385, 0.57412446]],
      [[0.551981, 0.2299962, 0.40206707, 0.7424828], [0.16304898, 0.26685357, 0.10787654, 0.48786318]]],
     [[[0.97532773, 0.52998006, 0.5693196, 0.28751576], [0.22973418, 0.5575429, 0.5877949, 0.461349]],
      [[0.320073, 0.69799054, 0.41638315, 0.13438594], [0.015848756, 0.45914185, 0.40993977, 0.031940937]],
      [[0.13979805, 0.24647367, 0.8555057, 0.40757453], [0.70918477, 0.9841, 0.93651617, 0.42834997]]]]
], dtype=t.float32).permute([0, 4, 1, 2, 3])
GT = t.as_tensor([[[[2, 0], [2, 2], [0, 3]], [[0, 2], [0, 3], [3, 2]]],
                  [[[2, 2], [3, 1], [2, 2]], [[1, 2], [1, 2], [1, 3]]]], dtype=t.int64)
WEIGHTS = t.as_tensor([
    [[[0.19875002, 0.77583194], [0.5079423, 0.10823226], [0.84881544, 0.38121593]],
     [[0.32796824, 0.6824727], [0.9398581, 0.45499086], [0.4005183, 0.025895357]]],
    [[[0.77079856, 0.5860559], [0.15548718, 0.40526056], [0.21678174, 0.81268084]],
     [[0.77574897, 0.27733755], [0.1688559, 0.69102776], [0.5144435, 0.42727184]]]], dtype=t.float32)


def test_iou_agnostic():
  np.testing.assert_allclose(O.iou_agnostic(GT, LOGITS), 0.8060565, rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(O.iou_agnostic(GT, LOGITS, WEIGHTS), 0.8174121, rtol=1e-5, atol=1e-6)


def test_iou_fgbg():
  np.testing.assert_allclose(O.iou_fgbg(GT, LOGITS), 0.3579613, rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(O.iou_fgbg(GT, LOGITS, WEIGHTS), 0.4265449, rtol=1e-5, atol=1e-6)


def test_xent():
  np.testing.assert_allclose(O.xent(GT, LOGITS), 1.4547757, rtol=1e-5, atol=1e-6)
  np.testing.assert_allclose(O.xent(GT, LOGITS, WEIGHTS), 0.7043564, rtol=1e-5, atol=1e-6)


def test_confusion_and_miou():
  # src/corenet/test/voxel_metrics_test.py:25-45 style check: cm[gt, pred]
  pred = t.tensor([0, 1, 1, 2, 2, 2])
  gt = t.tensor([0, 1, 2, 2, 2, 0])
  cm = O.confusion_matrix(pred, gt, 3)
  assert cm.tolist() == [[1, 0, 1], [0, 1, 0], [0, 1, 2]]
  np.testing.assert_allclose(O.mean_iou(cm), (1 / 2 + 2 / 4) / 2)


# src/corenet/test/voxel_metrics_test.py:25-80: the reference's own confusion-matrix / IoU known answers
VM_GT = t.tensor([[[3, 2, 2, 4], [4, 3, 2, 2], [3, 1, 3, 0]], [[3, 0, 1, 3], [2, 3, 1, 1], [2, 3, 0, 4]]], dtype=t.int32)
VM_PRED = t.tensor([[[0, 2, 3, 1], [1, 1, 1, 3], [4, 0, 2, 3]], [[1, 0, 1, 4], [2, 4, 4, 0], [4, 2, 4, 2]]], dtype=t.int32)
VM_CM = [[1, 0, 0, 1, 1], [2, 1, 0, 0, 1], [0, 1, 2, 2, 1], [1, 2, 2, 0, 3], [0, 2, 1, 0, 0]]
VM_IOU = [0.16666667, 0.11111111, 0.22222222, math.nan, math.nan]     # the test file says 0., the shipped code NaN (F10)
VM_PRECISION = [0.25, 0.16666667, 0.4, math.nan, math.nan]
VM_RECALL = [0.33333333, 0.25, 0.33333333, math.nan, math.nan]


def test_reference_confusion_matrix_known_answer():
  cm = O.confusion_matrix(VM_PRED, VM_GT, 5)
  assert cm.tolist() == VM_CM
  np.testing.assert_allclose(O.iou_per_class(cm).numpy(), VM_IOU, rtol=1e-6)
  # evaluation_results.py:262-266 is a pandas mean: the NaN classes drop out
  np.testing.assert_allclose(O.mean_iou(cm), np.mean(VM_IOU[1:3]), rtol=1e-6)


def test_evaluator_metrics_follow_the_reference():
  """Evaluator.tfpn / metrics / iou_per_class / mean_iou (host logic over the device confusion matrix) against
  voxel_metrics_test.py:47-80 and, when a copy of the reference is present, against its own
  compute_tfpn / compute_tfpn_fg / compute_voxel_metrics and the pandas mean of evaluation_results.py:188-266."""
  from corenet_b200.evaluator import Evaluator
  ev = Evaluator.__new__(Evaluator)
  ev.confusion_matrix = t.tensor(VM_CM, dtype=t.int64)
  tp, tn, fp, fn = ev.tfpn()
  assert tp.tolist() == [1, 1, 2, 0, 0] and tn.tolist() == [18, 15, 15, 13, 15]
  assert fp.tolist() == [3, 5, 3, 3, 6] and fn.tolist() == [2, 3, 4, 8, 3]
  mm = ev.metrics()
  np.testing.assert_allclose(mm["iou"][:-1].numpy(), VM_IOU, rtol=1e-6)
  np.testing.assert_allclose(mm["precision"][:-1].numpy(), VM_PRECISION, rtol=1e-6)
  np.testing.assert_allclose(mm["recall"][:-1].numpy(), VM_RECALL, rtol=1e-6)
  np.testing.assert_allclose(ev.mean_iou(), np.mean(VM_IOU[1:3]), rtol=1e-6)
  ev.confusion_matrix = t.tensor([[5, 0, 0], [0, 3, 0], [0, 0, 0]], dtype=t.int64)     # class 2 absent everywhere
  np.testing.assert_allclose(ev.mean_iou(), 1.0)
  ev.confusion_matrix = t.tensor([[5, 1], [2, 0]], dtype=t.int64)                       # no true positive at all
  assert math.isnan(ev.mean_iou()) and math.isnan(O.mean_iou(ev.confusion_matrix))
  from baseline import ref_import
  if ref_import.import_reference() is None:
    return
  import dataclasses
  import pandas
  from corenet import voxel_metrics as vm
  g = t.Generator().manual_seed(0)
  for k in (2, 5, 15):
    cm = t.randint(0, 50, (k, k), generator=g)
    cm[t.rand(k, generator=g) < 0.3, :] = 0           # classes without ground truth
    cm.fill_diagonal_(0) if k == 5 else None          # ... and a matrix without any true positive
    ev.confusion_matrix = cm.to(t.int64)
    ref = vm.compute_voxel_metrics(vm.compute_tfpn(cm))
    ref_fg = vm.compute_voxel_metrics(vm.compute_tfpn_fg(cm))
    mine = ev.metrics()
    for name in ("iou", "precision", "recall"):
      np.testing.assert_allclose(mine[name][:-1].numpy(), getattr(ref, name).numpy(), rtol=1e-12, equal_nan=True)
      np.testing.assert_allclose(mine[name][-1].numpy(), getattr(ref_fg, name).numpy(), rtol=1e-12, equal_nan=True)
    df = pandas.DataFrame(dataclasses.asdict(ref.cpu().numpy()), index=[f"c{i}" for i in range(k)]).T
    want = float(df.iloc[:, 1:].T.mean().iou)          # get_mean_iou without the trailing __global__ column
    np.testing.assert_allclose(ev.mean_iou(), want, rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(O.mean_iou(cm), want, rtol=1e-12, equal_nan=True)
