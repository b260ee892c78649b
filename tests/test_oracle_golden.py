"""Replays the committed reference fixture (tests/golden/corenet_reference.npz, written by oracle/make_golden.py
from the REAL reference) against the oracle restatement, without the reference being present."""
import numpy as np
import torch as t

from oracle import corenet_oracle as O
from oracle import make_golden as MG


def test_oracle_reproduces_reference_fixture(golden):
  from corenet_b200 import configuration as C
  from corenet_b200.model.core_net import CoreNet
  t.manual_seed(0)
  m = CoreNet(C.default_config(2))
  sd = {k: v.clone() for k, v in m.state_dict().items()}
  inp = MG.case_inputs("A")
  gt = MG.synthetic_gt(1, 2)
  for mode in ("eval", "train"):
    st = {k: v.clone().requires_grad_(v.dtype == t.float32 and "running" not in k) for k, v in sd.items()}
    logits = O.corenet_forward(st, inp["image"], inp["v2s"], inp["offsets"], mode == "train", {})
    loss = O.iou_fgbg(gt, logits)
    loss.backward()
    pre = f"A.{mode}."
    got = logits.detach().double().reshape(-1)[golden[pre + "logits.idx"]].numpy()
    np.testing.assert_allclose(got, golden[pre + "logits.val"], rtol=0, atol=1e-6 * float(golden[pre + "logits.max"]))
    assert abs(loss.item() - float(golden[pre + "loss"])) < 1e-6
    for name in ("decoder.stage_6.t1.weight", "encoder.stage2.a.op_a.conv.weight", "decoder.rt_skip_3.compress_channels.weight"):
      g = st[name].grad.double().reshape(-1)[golden[f"{pre}grad.{name}.idx"]].numpy()
      np.testing.assert_allclose(g, golden[f"{pre}grad.{name}.val"], rtol=0,
                                 atol=1e-5 * float(golden[f"{pre}grad.{name}.max"]) + 1e-12)
