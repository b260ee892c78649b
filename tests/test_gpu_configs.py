"""Parity on BASELINE.json's own configurations, against fixtures generated from the REAL reference
(oracle/make_golden_configs.py -> tests/golden/corenet_reference_configs.npz):

  C  h5 per-GPU batch   B=4, C=2,  train, iou_fgbg
  D  h7                 B=8, C=2,  eval, softmax + argmax + confusion matrix / mean IoU
  E  m7 / m9            B=2, C=15, train, xent_times_iou_agnostic

Kernel routing depends on the batch size and the channel counts (engine._conv_dispatch), so these are different
code paths from the B=1/B=2 cases of test_gpu_model.py.  Tolerances: forward 1e-3 of the tensor maximum (logits and
every decoder / encoder tap), mean IoU 0.1 pt, gradients "as accurate as the reference's own fp32 gradients" measured
against the fp64 answer (train mode at random init is chaotic, see test_gpu_model.py).
"""
import os

import numpy as np
import pytest
import torch as t

pytestmark = pytest.mark.gpu

from oracle import make_golden_configs as MGC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FWD_TOL = 1e-3


@pytest.fixture(scope="module")
def cfg_golden():
  return np.load(os.path.join(ROOT, "tests", "golden", "corenet_reference_configs.npz"), allow_pickle=False)


def build_model(classes):
  from corenet_b200 import configuration as C
  from corenet_b200.model.core_net import CoreNet
  t.manual_seed(0)
  return CoreNet(C.default_config(classes))


def check_samples(golden, key, ten, tol=FWD_TOL):
  idx, val, mx = golden[key + ".idx"], golden[key + ".val"], float(golden[key + ".max"])
  got = ten.detach().double().cpu().reshape(-1)[idx].numpy()
  err = np.abs(got - val).max() / mx
  assert err <= tol, f"{key}: {err:.3e}"
  return err


def plan_tap(plan, key):
  """Reference-layout (NCDHW / NCHW) copy of an engine buffer named like the oracle's taps."""
  def grid(buf, c):
    g = buf.spatial[0]
    return buf.v.view(plan.B, g, g, g, buf.cs)[..., :c].permute(0, 4, 1, 2, 3).contiguous()
  for sd in plan.stages:
    k = sd["stage"]
    if key == f"cat_{k}":
      return grid(sd["cat"], sd["cin"])
    if key == f"stage_{k}.c1":
      return grid(sd["c"], sd["mid"])
    if key == f"stage_{k}" and k < 6:
      return grid(sd["next"], sd["t_out"])
  raise KeyError(key)


def check_forward(golden, case, model, logits, need_grad):
  from corenet_b200 import engine
  eng = engine.get_engine(model)
  assert int(eng.tc_status) == 0, "tcgen05 conv kernel reported a barrier timeout"
  errs = {"logits": check_samples(golden, f"{case}.logits", logits)}
  plan = eng.plans[(logits.shape[0], need_grad)][0]
  for k in ("cat_3", "cat_4", "cat_5", "cat_6", "stage_2", "stage_3", "stage_4", "stage_5", "stage_3.c1",
            "stage_4.c1", "stage_5.c1", "stage_6.c1"):
    errs[k] = check_samples(golden, f"{case}.tap.{k}", plan_tap(plan, k))
  feats = plan.features_nchw()
  for name, f in zip(("stage1_64x128x128", "stage2_256x64x64", "stage3_512x32x32", "stage4_1024x16x16",
                      "stage5_2048x8x8", "global_average_2048"), feats):
    errs[name] = check_samples(golden, f"{case}.tap.{name}", f)
  worst = max(errs, key=errs.get)
  print(f"\n[{case}] forward: logits {errs['logits']:.2e}, worst tap {worst} {errs[worst]:.2e}")


def check_gradients(golden, case, model):
  """Sampled relative-L2 error per tensor against the fp64 gradients, next to the same error of the reference's own
  fp32 gradients: the CUDA path must be as accurate as the reference (<= 4x its error + 3e-3 for >= 90 % of the
  tensors, comparable median) -- the acceptance rule of test_gpu_model.test_forward_backward_parity."""
  e_mine, e_ref, names = [], [], []
  gscale = max(float(golden[f"{case}.grad.{n}.norm64"]) for n, _ in model.named_parameters())
  for n, p in model.named_parameters():
    assert p.grad is not None, n
    idx = golden[f"{case}.grad.{n}.idx"]
    g64, g32 = golden[f"{case}.grad.{n}.f64"], golden[f"{case}.grad.{n}.ref32"]
    mine = p.grad.detach().double().cpu().reshape(-1)[idx].numpy()
    den = np.linalg.norm(g64)
    if float(golden[f"{case}.grad.{n}.norm64"]) <= 1e-12 * gscale or den == 0.0:
      assert np.abs(mine).max() <= 1e-5 * gscale, n
      continue
    e_mine.append(np.linalg.norm(mine - g64) / den)
    e_ref.append(np.linalg.norm(g32 - g64) / den)
    names.append(n)
  e_mine, e_ref = np.array(e_mine), np.array(e_ref)
  worst = int(np.argmax(e_mine))
  print(f"[{case}] grad rel-L2 (sampled) vs fp64: CUDA median {np.median(e_mine):.2e} max {e_mine.max():.2e} "
        f"({names[worst]}); reference fp32 median {np.median(e_ref):.2e} max {e_ref.max():.2e}")
  ok = e_mine <= 4 * e_ref + 3e-3
  assert ok.mean() >= 0.9, f"only {ok.mean():.2%} of the gradient tensors are as accurate as the reference's"
  assert np.median(e_mine) <= 3 * np.median(e_ref) + 1e-3


@pytest.mark.parametrize("case", ["C", "E"])
def test_train_configs_match_reference(cfg_golden, case):
  from corenet_b200.model import losses
  dev = t.device("cuda", 0)
  inp = MGC.config_inputs(case)
  m = build_model(inp["classes"]).to(dev).train()
  logits = m(inp["image"].to(dev), inp["v2s"].to(dev), inp["offsets"].to(dev))
  loss = getattr(losses, MGC.LOSS[case])(inp["gt"].to(dev), logits)
  loss.backward()
  check_forward(cfg_golden, case, m, logits, True)
  ref_loss = float(cfg_golden[f"{case}.loss"])
  print(f"[{case}] loss {loss.item():.7f} vs reference {ref_loss:.7f}")
  assert abs(loss.item() - ref_loss) <= 1e-4 * max(1.0, abs(ref_loss))
  check_gradients(cfg_golden, case, m)
  bufs = dict(m.named_buffers())
  for k, v in bufs.items():
    if k.endswith("num_batches_tracked"):
      assert int(v) == 1
    else:
      idx, val = cfg_golden[f"{case}.buf.{k}.idx"], cfg_golden[f"{case}.buf.{k}.val"]
      got = v.detach().double().cpu().reshape(-1)[idx].numpy()
      assert np.abs(got - val).max() <= 3e-4 * float(cfg_golden[f"{case}.buf.{k}.max"]), k


def test_h7_eval_batch8_confusion_and_miou(cfg_golden):
  """B=8 eval forward + softmax + argmax + confusion through the Evaluator (eager, eager, captured graph, replay):
  the counts of every pass must equal the reference's up to near-tie voxels, mean IoU within 0.1 pt."""
  from corenet_b200.evaluator import Evaluator
  dev = t.device("cuda", 0)
  inp = MGC.config_inputs("D")
  m = build_model(2).to(dev).eval()
  ev = Evaluator(m)
  args = [inp["image"].to(dev), inp["v2s"].to(dev), inp["offsets"].to(dev), inp["gt"].to(dev)]
  cm_ref = t.from_numpy(cfg_golden["D.cm"])
  from oracle import corenet_oracle as O
  for i in range(4):
    if ev.confusion_matrix is not None:
      ev.confusion_matrix.zero_()
    pmf = ev.add_batch(*args)
    cm = ev.confusion_matrix.cpu()
    assert int(cm.sum()) == inp["gt"].numel()
    assert (cm - cm_ref).abs().sum().item() <= 1e-5 * inp["gt"].numel(), (i, cm, cm_ref)
    assert abs(ev.mean_iou() - float(cfg_golden["D.miou"])) <= 1e-3
    assert abs(ev.mean_iou() - O.mean_iou(cm)) < 1e-12
  assert ev.graph_launches > 100, "the fourth batch must have been a graph replay"
  ev.check_status()
  eng = ev.eng
  plan = eng.plans[(8, False)][0]
  check_forward(cfg_golden, "D", m, plan.logits, False)
  assert t.allclose(pmf, plan.logits.softmax(1), atol=1e-6)
  assert t.allclose(pmf.sum(1), t.ones_like(pmf[:, 0]), atol=1e-5)
  # prefetch path with pinned host inputs
  host = [x.pin_memory() for x in (inp["image"], inp["v2s"], inp["offsets"], inp["gt"])]
  ev.confusion_matrix.zero_()
  ev.prefetch(*host)
  ev.add_batch()
  assert (ev.confusion_matrix.cpu() - cm_ref).abs().sum().item() <= 1e-5 * inp["gt"].numel()


def test_fgbg_labeled_confusion_matches_reference_rule():
  """FG_BG evaluation scales 0/1 labels by the scene's dataset class (evaluation_results.py:40-51)."""
  from corenet_b200 import _lib
  dev = t.device("cuda", 0)
  g = t.Generator().manual_seed(5)
  b, c, k, s = 3, 2, 6, 4096
  logits = t.randn(b, c, 16, 16, 16, generator=g)
  gt = t.randint(0, 2, (b, 16, 16, 16), generator=g, dtype=t.int32)
  labels = t.tensor([1, 5, 3], dtype=t.int32)
  pred = logits.argmax(1).to(t.int64) * labels[:, None, None, None]
  gts = gt.to(t.int64) * labels[:, None, None, None]
  exp = t.bincount((gts * k + pred).reshape(-1), minlength=k * k).reshape(k, k)
  cm = t.zeros(k, k, dtype=t.int64, device=dev)
  d_logits, d_gt, d_labels = logits.to(dev), gt.to(dev), labels.to(dev)
  _lib.call("crn_argmax_confusion_labeled", d_logits.data_ptr(), d_gt.data_ptr(), 0, b, c, s,
            d_labels.data_ptr(), k, cm.data_ptr(), _lib.stream_ptr())
  assert t.equal(cm.cpu(), exp)


def test_module_path_graph_replay_matches_eager():
  """CoreNet.forward + autograd backward: calls 1-2 enqueue eagerly, call 3 captures CUDA graphs, call 4 replays.
  In eval mode (well conditioned) all four must agree; weights changed in between (optimizer.step through torch)
  must be picked up by the replay (re-pack outside the graph)."""
  from corenet_b200 import engine
  from corenet_b200.model import losses
  from oracle import make_golden as MG
  dev = t.device("cuda", 0)
  inp = MG.case_inputs("B")
  gt = MG.synthetic_gt(2, 2).to(dev)
  m = build_model(2).to(dev).eval()
  args = [inp["image"].to(dev), inp["v2s"].to(dev), inp["offsets"].to(dev)]
  outs = []
  for i in range(4):
    m.zero_grad()
    logits = m(*args)
    losses.iou_fgbg(gt, logits).backward()
    outs.append((logits.detach().clone(), t.cat([p.grad.reshape(-1) for p in m.parameters()])))
  plan = engine.get_engine(m).plans[(2, True)][0]
  assert any(k[0] == "fwd" and v["graph"] is not None for k, v in plan._graphs.items())
  assert any(k[0] == "bwd" and v["graph"] is not None for k, v in plan._graphs.items())
  for lo, g in outs[1:]:
    assert (lo - outs[0][0]).abs().max().item() <= 1e-5 * outs[0][0].abs().max().item()
    assert ((g - outs[0][1]).norm() / outs[0][1].norm()).item() <= 1e-4
  # change the weights through torch: the replayed graph must see them
  with t.no_grad():
    for p in m.parameters():
      p.mul_(1.01)
  lo2 = m(*args).detach()
  t.cuda.synchronize()
  engine.USE_GRAPHS = False
  try:
    lo2_eager = m(*args).detach()
  finally:
    engine.USE_GRAPHS = True
  assert (lo2 - outs[0][0]).abs().max().item() > 1e-4 * outs[0][0].abs().max().item()
  assert (lo2 - lo2_eager).abs().max().item() <= 1e-5 * lo2_eager.abs().max().item()
  # a forward with grad that is never back-propagated must release its plan (no new plan per call)
  for _ in range(3):
    m(*args)
  assert len(engine.get_engine(m).plans[(2, True)]) <= 2


def test_eval_forward_between_trainer_steps_sees_new_weights():
  """the fused Adam kernel updates the weights through raw pointers (also inside graph replays); an eval
  forward between training steps must re-pack them."""
  from corenet_b200 import engine
  from corenet_b200.trainer import Trainer
  from oracle import make_golden as MG
  dev = t.device("cuda", 0)
  inp = MG.case_inputs("A")
  gt = MG.synthetic_gt(1, 2).to(dev)
  m = build_model(2).to(dev).eval()
  tr = Trainer(m, lr=1e-2, eps=1e-4)
  args = [inp["image"].to(dev), inp["v2s"].to(dev), inp["offsets"].to(dev)]
  for i in range(5):
    tr.step(*args, gt)
    with t.no_grad():
      a = m(*args).clone()
    eng = engine.get_engine(m)
    eng._ver_sig = None            # force a fresh pack: the reference answer for the current weights
    with t.no_grad():
      b = m(*args).clone()
    # (atomics-order noise between two forwards is ~1e-6; one stale Adam step at lr = 1e-2 moves the logits by percents)
    assert (a - b).abs().max().item() <= 1e-5 * b.abs().max().item(), f"stale weights after step {i + 1}"
  assert tr.graph_launches > 100
  tr.check_status(wait=True)


def test_trainer_rejects_bad_label_dtype_and_guards_failed_steps():
  from corenet_b200.trainer import Trainer
  from oracle import make_golden as MG
  dev = t.device("cuda", 0)
  inp = MG.case_inputs("A")
  m = build_model(2).to(dev).train()
  tr = Trainer(m, use_graph=False)
  args = [inp["image"].to(dev), inp["v2s"].to(dev), inp["offsets"].to(dev)]
  with pytest.raises(AssertionError):
    tr.step(*args, MG.synthetic_gt(1, 2).to(dev).to(t.uint8))
  gt = MG.synthetic_gt(1, 2).to(dev)
  tr.step(*args, gt)
  before = tr.flat.clone()
  tr.eng.tc_status.fill_(1)          # simulate an mbarrier timeout reported by a tcgen05 kernel
  loss = tr.step(*args, gt)
  assert t.isnan(loss).all(), "a failed step must surface as a NaN loss"
  assert t.equal(tr.flat, before) and int(tr.step_dev) == 1, "the guarded Adam kernel must skip a failed step"
  with pytest.raises(RuntimeError):
    tr.check_status(wait=True)
  tr.eng.tc_status.zero_()
  tr.step(*args, gt)
  assert int(tr.step_dev) == 2 and not t.equal(tr.flat, before)


@pytest.mark.parametrize("c,mode", [(15, 1), (15, 0), (6, 1), (2, 0), (3, 1)])
def test_loss_kernels_rows_layout_equals_planar(c, mode):
  """The channels-last rows forms (crn_*_l with rows_cp > 0: what the padded logits layer writes for C > 4) against
  the planar forms on the same values: sums / loss / gradient / softmax / confusion must agree."""
  from corenet_b200 import _lib
  dev = t.device("cuda", 0)
  g = t.Generator().manual_seed(c * 7 + mode)
  b, s = 2, 24 * 24 * 24
  cp = (c + 3) // 4 * 4
  logits = t.randn(b, c, s, generator=g) * 3
  gt = t.randint(0, c, (b, s), generator=g, dtype=t.int32).to(dev)
  planar = logits.to(dev).contiguous()
  rows = t.full((b * s, cp), 7.0, device=dev)                    # pad channels hold garbage: must be ignored
  rows[:, :c] = logits.permute(0, 2, 1).reshape(b * s, c).to(dev)
  st = _lib.stream_ptr()
  out = {}
  for name, x, rc in (("planar", planar, 0), ("rows", rows, cp)):
    sums = t.empty(4 * b, dtype=t.float64, device=dev)
    loss = t.empty(1, device=dev)
    coef = t.empty(2 * b + 1, device=dev)
    _lib.call("crn_loss_sums_l", x.data_ptr(), rc, gt.data_ptr(), 0, b, c, s, mode, sums.data_ptr(), st)
    _lib.call("crn_loss_finalize", sums.data_ptr(), b, c, s, mode, loss.data_ptr(), coef.data_ptr(), st)
    d = t.full_like(x, 5.0)
    _lib.call("crn_loss_bwd_l", x.data_ptr(), rc, gt.data_ptr(), 0, b, c, s, mode, coef.data_ptr(), None,
              d.data_ptr(), rc, st)
    pmf = t.empty(b, c, s, device=dev)
    _lib.call("crn_softmax_l", x.data_ptr(), rc, b, c, s, pmf.data_ptr(), st)
    cm = t.zeros(c, c, dtype=t.int64, device=dev)
    _lib.call("crn_argmax_confusion_l", x.data_ptr(), rc, gt.data_ptr(), 0, b, c, s, None, c, cm.data_ptr(), st)
    out[name] = (sums.cpu(), loss.cpu(), d.cpu(), pmf.cpu(), cm.cpu())
  p, r = out["planar"], out["rows"]
  assert t.allclose(p[0][:3 * b], r[0][:3 * b], rtol=1e-12) or t.allclose(p[0], r[0], rtol=1e-9, atol=1e-9)
  assert t.allclose(p[1], r[1], rtol=1e-6)
  d_rows = r[2].reshape(b, s, cp)
  assert t.allclose(p[2], d_rows[..., :c].permute(0, 2, 1), rtol=1e-5, atol=1e-9)
  assert (d_rows[..., c:] == 0).all()                              # pad channels of the gradient are written as 0
  assert t.equal(p[3], r[3]) and t.equal(p[4], r[4])
  # reference arithmetic (torch) on the planar values
  from oracle import corenet_oracle as O
  lg = logits.reshape(b, c, 24, 24, 24).clone().requires_grad_(True)
  fn = O.iou_fgbg if mode == 0 else O.xent_times_iou_agnostic
  lo = fn(gt.cpu().long().reshape(b, 24, 24, 24), lg)
  lo.backward()
  assert abs(float(p[1]) - float(lo)) <= 1e-5 * max(1.0, abs(float(lo)))
  assert ((p[2].reshape(lg.shape) - lg.grad).abs().max() / lg.grad.abs().max()).item() <= 1e-4
  # rows -> planar conversion
  back = t.empty(b, c, s, device=dev)
  _lib.call("crn_rows_to_planar", rows.data_ptr(), b, c, s, cp, back.data_ptr(), st)
  assert t.equal(back, planar)


def test_trainer_rows_path_matches_module_path_c15():
  """C = 15: the Trainer keeps logits and their gradient as channels-last rows (padded tcgen05 logits layer, rows loss
  kernels); the module path converts to / from the reference's planar layout.  Same step, same gradients."""
  from corenet_b200.model import losses
  from corenet_b200.trainer import Trainer
  from corenet_b200 import engine
  dev = t.device("cuda", 0)
  inp = MGC.config_inputs("E")
  args = [inp["image"].to(dev), inp["v2s"].to(dev), inp["offsets"].to(dev)]
  gt = inp["gt"].to(dev)
  m = build_model(15).to(dev).eval()
  logits = m(*args)
  loss = losses.xent_times_iou_agnostic(gt, logits)
  loss.backward()
  g_mod = t.cat([p.grad.reshape(-1) for p in m.parameters()])
  plan = engine.get_engine(m).plans[(2, True)][0]
  assert plan.rows_cp == 16, "C = 15 must run the padded rows path"
  m2 = build_model(15).to(dev).eval()
  tr = Trainer(m2, loss="xent_times_iou_agnostic", use_graph=False, lr=0.0)
  l2 = tr.step(*args, gt)
  err = ((tr.grad - g_mod).double().norm() / g_mod.double().norm()).item()
  print(f"\nC=15 Trainer(rows) vs module(planar): loss {float(l2):.7f} vs {float(loss):.7f}, grad rel-L2 {err:.2e}")
  assert abs(float(l2) - float(loss)) <= 1e-5 * abs(float(loss))
  assert err <= 3e-4
  tr.check_status(wait=True)


def test_trainer_prefetch_with_device_resident_gt_pipeline():
  """m9-style loop: the ground truth of step k+1 (rasterise + fill + label merge on the GPU) is produced by a callable
  on the copy stream while step k runs; the losses must equal those of the same steps fed with precomputed grids."""
  from corenet_b200.data import batched_example as be
  from corenet_b200.trainer import Trainer
  from tests.conftest import cube_mesh
  dev = t.device("cuda", 0)
  inp = MGC.config_inputs("E")
  b = inp["image"].shape[0]
  meshes = [cube_mesh(0.99) / 3.0 * 0.5 + 0.1, cube_mesh(0.99) / 3.0 * 0.35 + 0.5, cube_mesh(0.99) / 3.0 * 0.3 + 0.3]
  verts = t.from_numpy(np.concatenate(meshes)).pin_memory()
  ntri = [t.tensor([12], dtype=t.int32), t.tensor([12, 12], dtype=t.int32)]
  labels = [[3], [7, 12]]
  offs = inp["offsets"]

  def gt_fn():
    return be.voxelize(verts, ntri, offs, (128, 128, 128), be.VoxelContentSemanticLabel(labels),
                       image_resolution_multiplier=8, conservative_rasterization=False)[1]
  grid = gt_fn()
  assert grid.dtype == t.int32 and tuple(grid.shape) == (b, 128, 128, 128) and int(grid.max()) == 12
  host = [x.pin_memory() for x in (inp["image"], inp["v2s"], inp["offsets"])]
  runs = []
  for use_fn in (False, True):
    tr = Trainer(build_model(15).to(dev).eval(), loss="xent_times_iou_agnostic")
    tr.prefetch(*host, gt_fn if use_fn else grid)
    ls = []
    for _ in range(4):
      loss = tr.step()
      tr.prefetch(*host, gt_fn if use_fn else grid)
      ls.append(float(loss))
    tr.step()
    tr.check_status(wait=True)
    runs.append(ls)
  print("\nprecomputed GT:", runs[0], "\nGT callable   :", runs[1])
  assert np.isfinite(runs[0]).all() and abs(runs[0][0] - runs[1][0]) <= 1e-5 * abs(runs[0][0])
  assert max(abs(x - y) for x, y in zip(*runs)) <= 2e-3 * abs(runs[0][0])


def test_evaluator_fgbg_labels_and_semantic_rows():
  """Evaluator end to end: FG_BG task (C=2, predictions / GT scaled by the scene's class id into a K x K matrix,
  evaluation_results.py:40-51) and SEMANTIC task (C=15: the rows-layout logits path) against torch on the logits the
  module itself returns."""
  from corenet_b200.evaluator import Evaluator
  dev = t.device("cuda", 0)
  inp = MGC.config_inputs("E")
  args = [inp["image"].to(dev), inp["v2s"].to(dev), inp["offsets"].to(dev)]
  # SEMANTIC
  m = build_model(15).to(dev).eval()
  gt = inp["gt"].to(dev).to(t.int32)
  ev = Evaluator(m)
  for _ in range(4):
    pmf = ev.add_batch(*args, gt)
  with t.no_grad():
    logits = m(*args)
  # two forwards of the same weights differ by split-K atomics order (~1e-5 of the logit range, |logit| ~ 5): 1e-4 on
  # the probabilities; a layout / channel mix-up would be O(1)
  assert t.allclose(pmf, logits.softmax(1), atol=1e-4)
  pred = logits.argmax(1)
  exp = t.bincount((gt.long() * 15 + pred).reshape(-1), minlength=225).reshape(15, 15)
  assert (ev.confusion_matrix.cpu() - 4 * exp.cpu()).abs().sum().item() <= 4e-5 * gt.numel()
  assert ev.graph_launches > 100
  # FG_BG with dataset classes
  m2 = build_model(2).to(dev).eval()
  gt2 = (inp["gt"] > 0).to(t.int32).to(dev)
  labels = t.tensor([4, 9], dtype=t.int32, device=dev)
  ev2 = Evaluator(m2, num_classes=14)
  ev2.add_batch(*args, gt2, labels)
  with t.no_grad():
    pred2 = m2(*args).argmax(1).long() * labels.long()[:, None, None, None]
  exp2 = t.bincount(((gt2.long() * labels.long()[:, None, None, None]) * 14 + pred2).reshape(-1), minlength=196).reshape(14, 14)
  assert (ev2.confusion_matrix.cpu() - exp2.cpu()).abs().sum().item() <= 1e-5 * gt2.numel()
  with pytest.raises(ValueError):
    ev2.add_batch(*args, gt2)          # FG_BG needs the scene labels


def test_precision_mode_tf32_is_an_opt_in_with_its_own_bound(cfg_golden):
  """engine.set_precision("tf32") issues only the hi x hi product in every tcgen05 kernel (precision
  modes as measured options).  Eval-mode logits then sit at the 1e-3 level of the oracle (profiles/r02_precision_study.md:
  1.3e-3 on case B) instead of 1e-6: bound 1e-2, and the mode must actually change the arithmetic.  A Trainer keeps
  working across a switch (its graphs are keyed on the mode) and the default comes back bit-equal."""
  from corenet_b200 import engine
  from corenet_b200.trainer import Trainer
  from oracle import corenet_oracle as O
  from oracle import make_golden as MG
  dev = t.device("cuda", 0)
  inp = MG.case_inputs("A")
  gt = MG.synthetic_gt(1, 2)
  m = build_model(2)
  sd = {k: v.clone() for k, v in m.state_dict().items()}
  ref = O.corenet_forward(sd, inp["image"], inp["v2s"], inp["offsets"], False, {}, {})
  m = m.to(dev).eval()
  args = [inp["image"].to(dev), inp["v2s"].to(dev), inp["offsets"].to(dev)]
  rel = lambda a: ((a.cpu().double() - ref.double()).abs().max() / ref.double().abs().max()).item()
  try:
    with t.no_grad():
      full = m(*args).clone()
      engine.set_precision("tf32")
      fast = m(*args).clone()
    e_full, e_fast = rel(full), rel(fast)
    print(f"\nlogits rel err vs oracle: 3xtf32 {e_full:.2e}, tf32 {e_fast:.2e}")
    assert e_full <= FWD_TOL
    assert 5 * e_full < e_fast <= 1e-2, "single-pass TF32 must be measurably coarser, and still sane"
    tr = Trainer(m, lr=1e-4, eps=1e-4)
    l_fast = [float(tr.step(*args, gt.to(dev))) for _ in range(4)]
    engine.set_precision("3xtf32")
    l_full = [float(tr.step(*args, gt.to(dev))) for _ in range(4)]
    tr.check_status(wait=True)
    assert all(np.isfinite(l_fast + l_full))
    assert len(tr._graphs) == 2, "one captured step per precision mode"
    with pytest.raises(ValueError):
      engine.set_precision("fp8")
  finally:
    engine.set_precision("3xtf32")
  m2 = build_model(2).to(dev).eval()
  with t.no_grad():
    again = m2(*args)
  # (two runs differ by split-K atomics order at the 1e-6 level; single-pass TF32 would be 1e-3)
  assert (again - full).abs().max().item() <= 2e-5 * full.abs().max().item(), "the default arithmetic is back"
