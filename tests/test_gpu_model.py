"""End-to-end parity of the CUDA CoreNet path against the oracle and the committed reference fixtures.

Tolerances: forward max|a-b|/max|b| <= 1e-3 per tensor; gradients <= 1e-2 (relative L2).
The net amplifies rounding ~1000x at init in train mode (DESIGN.md "Precision"), so fp32
re-association alone shows up at the 1e-4 level in the logits.  Gradients are compared in the relative
L2 norm against the fp64 oracle: with ~1e8 ReLU units a handful sit within rounding distance of 0, and ONE
flipped unit (measured: 1 of 131072 at encoder.stage5.b) moves single gradient entries by percents while
leaving L2 at 1e-3.
"""
import numpy as np
import pytest
import torch as t

pytestmark = pytest.mark.gpu

from oracle import corenet_oracle as O
from oracle import make_golden as MG

FWD_TOL = 1e-3
GRAD_TOL = 1e-2


def rel_err(a, b):
  a, b = a.detach().double().cpu(), b.detach().double().cpu()
  return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def build_model(classes=2):
  from corenet_b200 import configuration as C
  from corenet_b200.model.core_net import CoreNet
  t.manual_seed(0)
  return CoreNet(C.default_config(classes))


def oracle_run(sd, inp, gt, training, loss_name="iou_fgbg", dtype=t.float32):
  """fp32: the oracle (= the reference's arithmetic).  fp64: the exact answer, used to measure how much
  of a gradient mismatch is rounding noise that the reference's own fp32 path has too."""
  st = {k: (v.clone().to(dtype) if v.dtype == t.float32 else v.clone()) for k, v in sd.items()}
  for k, v in st.items():
    if v.dtype == dtype and "running" not in k:
      v.requires_grad_(True)
  nb, taps = {}, {}
  logits = O.corenet_forward(st, inp["image"], inp["v2s"].to(dtype), inp["offsets"].to(dtype), training, nb, taps,
                             dtype=dtype)
  loss = getattr(O, loss_name)(gt, logits)
  loss.backward()
  return logits.detach(), loss.item(), {k: v.grad for k, v in st.items() if v.grad is not None}, nb, taps


def check_golden(golden, prefix, name, ten, tol):
  idx = golden[f"{prefix}{name}.idx"]
  val = golden[f"{prefix}{name}.val"]
  mx = float(golden[f"{prefix}{name}.max"])
  got = ten.detach().double().cpu().reshape(-1)[idx].numpy()
  assert np.abs(got - val).max() <= tol * mx, f"{name}: {np.abs(got - val).max() / mx:.3e}"


def test_seeded_weights_match_reference_fixture(golden):
  """Same names, order and seeded values as the reference's CoreNet (fixture written by make_golden.py)."""
  m = build_model()
  sd = m.state_dict()
  assert list(sd.keys()) == [str(k) for k in golden["param_names"]]
  got = np.array([v.double().abs().sum().item() for v in sd.values()])
  np.testing.assert_allclose(got, golden["param_abssum"], rtol=1e-12)


@pytest.mark.parametrize("case,mode", [("A", "train"), ("A", "eval"), ("B", "train"), ("B", "eval")])
def test_forward_backward_parity(golden, case, mode):
  from corenet_b200.model import losses
  dev = t.device("cuda", 0)
  inp = MG.case_inputs(case)
  m = build_model(inp["classes"])
  sd = {k: v.clone() for k, v in m.state_dict().items()}
  if inp["perturb"]:
    sd = MG.perturb_brn(sd)
    m.load_state_dict(sd)
  gt = MG.synthetic_gt(inp["image"].shape[0], inp["classes"])
  training = mode == "train"
  lo, loss_o, grads_o, nb, taps = oracle_run(sd, inp, gt, training)
  m = m.to(dev)
  m.train(training)
  logits = m(inp["image"].to(dev), inp["v2s"].to(dev), inp["offsets"].to(dev))
  assert logits.shape == lo.shape and logits.is_contiguous()
  loss = losses.iou_fgbg(gt.to(dev), logits)
  loss.backward()
  e = rel_err(logits, lo)
  from corenet_b200 import engine as engine_lib
  assert int(engine_lib.get_engine(m).tc_status) == 0, "tcgen05 conv kernel reported a barrier timeout"
  print(f"\n[{case}/{mode}] logits rel err {e:.3e}  loss {loss.item():.7f} vs {loss_o:.7f}")
  assert e <= FWD_TOL
  assert abs(loss.item() - loss_o) <= 1e-4
  # committed fixture of the REAL reference run (values at seeded indices)
  pre = f"{case}.{mode}."
  check_golden(golden, pre, "logits", logits, FWD_TOL)
  assert abs(loss.item() - float(golden[pre + "loss"])) <= 1e-4
  # gradients.  Metric: relative L2 per tensor, measured against the fp64 ("exact") oracle, next to the
  # same error of the fp32 oracle (= the reference's own arithmetic).  Eval mode is well conditioned:
  # absolute bound GRAD_TOL.  Train mode at random init is chaotic (the REFERENCE's fp32 gradients are
  # themselves several % away from the exact ones, see DESIGN.md "Precision"), so there the CUDA path must be
  # as accurate as the reference: error <= 4x the reference's own error (+1e-3) for >= 90% of the tensors
  # and a comparable median.
  _, _, g64, _, _ = oracle_run(sd, inp, gt, training, dtype=t.float64)
  gscale = max(g.abs().max().item() for g in g64.values())
  e_mine, e_ref, names = [], [], []
  for n, p in m.named_parameters():
    assert p.grad is not None, n
    g_true = g64[n]
    if g_true.abs().max().item() <= 1e-12 * gscale:      # structurally zero gradient: absolute check
      assert p.grad.abs().max().item() <= 1e-5 * gscale, n
      continue
    den = g_true.norm().item()
    e_mine.append((p.grad.cpu().double() - g_true).norm().item() / den)
    e_ref.append((grads_o[n].double() - g_true).norm().item() / den)
    names.append(n)
  e_mine, e_ref = np.array(e_mine), np.array(e_ref)
  worst = int(np.argmax(e_mine))
  print(f"[{case}/{mode}] grad rel-L2 vs fp64: CUDA median {np.median(e_mine):.2e} max {e_mine.max():.2e} "
        f"({names[worst]}); fp32 oracle median {np.median(e_ref):.2e} max {e_ref.max():.2e}")
  if not training and not inp["perturb"]:
    assert e_mine.max() <= GRAD_TOL, (names[worst], e_mine.max())
  elif not training:
    # case B in eval mode: perturbed running statistics blow the activations up to ~1e6, so isolated
    # ReLU sign flips (fp32 noise) dominate single tensors; bound the distribution instead of the max
    assert np.median(e_mine) <= GRAD_TOL and e_mine.max() <= 5 * GRAD_TOL, (names[worst], e_mine.max())
  else:
    ok = e_mine <= 4 * e_ref + 3e-3
    assert ok.mean() >= 0.9, f"only {ok.mean():.2%} of the gradient tensors are as accurate as the reference's"
    assert np.median(e_mine) <= 3 * np.median(e_ref) + 1e-3
  # running statistics (train mode mutates the buffers exactly like the reference)
  if training:
    bufs = dict(m.named_buffers())
    for k, v in nb.items():
      if k.endswith("num_batches_tracked"):
        assert int(bufs[k]) == int(v)
      else:
        # running means of the deep decoder layers are ~1e-3 with heavy cancellation: they sit at 0.3 .. 1.0e-4 of
        # their max depending on the atomics order of the run -> 3e-4 (the forward bar of the path is 1e-3)
        assert rel_err(bufs[k], v) <= 3e-4, k


def test_encoder_features_and_stage_outputs():
  """Every encoder feature map against the oracle (eval mode: well conditioned -> tight tolerance)."""
  dev = t.device("cuda", 0)
  inp = MG.case_inputs("A")
  m = build_model()
  sd = m.state_dict()
  x = O.preprocess_image_caffe(inp["image"])
  f = O.resnet50_features({k: v for k, v in sd.items()}, x, False)
  m = m.to(dev).eval()
  got = m.encode_features(inp["image"].to(dev))
  for name, a, b in zip(f._fields, got, f):
    assert rel_err(a, b) <= 2e-4, name


def test_semantic_head_and_loss():
  """C=15 (multi-object configs m7/m9): forward + xent_times_iou_agnostic backward."""
  from corenet_b200.model import losses
  dev = t.device("cuda", 0)
  inp = MG.case_inputs("A")
  m = build_model(15)
  sd = {k: v.clone() for k, v in m.state_dict().items()}
  gt = MG.synthetic_gt(1, 15)
  lo, loss_o, grads_o, _, _ = oracle_run(sd, inp, gt, False, "xent_times_iou_agnostic")
  m = m.to(dev).eval()
  logits = m(inp["image"].to(dev), inp["v2s"].to(dev), inp["offsets"].to(dev))
  loss = losses.xent_times_iou_agnostic(gt.to(dev), logits)
  loss.backward()
  assert rel_err(logits, lo) <= FWD_TOL
  assert abs(loss.item() - loss_o) <= 1e-4 * abs(loss_o)
  g = dict(m.named_parameters())["decoder.stage_6.t1.weight"].grad
  assert rel_err(g, grads_o["decoder.stage_6.t1.weight"]) <= GRAD_TOL


def test_trainer_semantic_cuda_graph():
  """m7/m9-style step (C = 15, xent_times_iou_agnostic, B = 2) through the captured graph: the loss must fall and every
  tcgen05 kernel must report a clean status."""
  from corenet_b200.trainer import Trainer
  dev = t.device("cuda", 0)
  inp = MG.case_inputs("A")
  gt = MG.synthetic_gt(1, 15)
  m = build_model(15).to(dev).train()
  tr = Trainer(m, lr=4e-4, eps=1e-4, loss="xent_times_iou_agnostic")
  args = [t.cat([inp["image"]] * 2).to(dev), t.cat([inp["v2s"]] * 2).to(dev), t.cat([inp["offsets"]] * 2).to(dev),
          t.cat([gt] * 2).to(dev)]
  losses_ = [tr.step(*args).item() for _ in range(5)]
  assert tr.graph_launches > 100 and int(tr.eng.tc_status) == 0
  assert np.isfinite(losses_).all() and min(losses_[1:]) < losses_[0] + 5e-3, losses_      # finite, no divergence


def test_no_cpu_fallback():
  m = build_model()
  inp = MG.case_inputs("A")
  with pytest.raises(RuntimeError):
    m(inp["image"], inp["v2s"], inp["offsets"])


def test_trainer_step_matches_reference_adam():
  """Trainer (flat buffers, fused Adam) vs the oracle + torch.optim.Adam for two steps in eval-mode BN
  (well conditioned): parameters after the steps must agree."""
  from corenet_b200.trainer import Trainer
  dev = t.device("cuda", 0)
  inp = MG.case_inputs("A")
  gt = MG.synthetic_gt(1, 2)
  m = build_model()
  sd = {k: v.clone() for k, v in m.state_dict().items()}
  # oracle side
  params = {k: v.detach().clone().requires_grad_(True) for k, v in m.named_parameters()}
  state = dict(sd); state.update(params)
  opt = t.optim.Adam(list(params.values()), lr=4e-4, eps=1e-4)
  losses_o = []
  for _ in range(2):
    opt.zero_grad()
    loss = O.iou_fgbg(gt, O.corenet_forward(state, inp["image"], inp["v2s"], inp["offsets"], False))
    loss.backward(); opt.step(); losses_o.append(loss.item())
  # CUDA side
  m = m.to(dev).eval()
  tr = Trainer(m, lr=4e-4, eps=1e-4, loss="iou_fgbg")
  args = [inp["image"].to(dev), inp["v2s"].to(dev), inp["offsets"].to(dev), gt.to(dev)]
  losses_c = [tr.step(*args).item() for _ in range(2)]
  assert abs(losses_c[0] - losses_o[0]) < 1e-5 and abs(losses_c[1] - losses_o[1]) < 1e-4
  # Adam's first steps move every weight by ~lr: compare the UPDATE, not the weight
  errs = []
  for n, p in m.named_parameters():
    du_c = (p.detach().cpu() - sd[n]).double()
    du_o = (params[n].detach() - sd[n]).double()
    if du_o.abs().max() < 1e-9:
      continue
    errs.append((((du_c - du_o).norm() / du_o.norm()).item(), n))
  errs.sort(reverse=True)
  print("worst update errors:", errs[:6])
  # tiny bias vectors whose gradients are sums over few voxels inherit isolated ReLU sign flips (see module
  # docstring); bound the distribution: median tight, worst loose.  Adam's first updates are lr*sign(g): ONE flipped
  # near-zero gradient in a 40-element vector is already a relative-L2 error of sqrt(4/40) = 0.32, and which
  # element flips depends on the atomics order of the run, hence 0.5 for the worst vector.
  med = sorted(e for e, _ in errs)[len(errs) // 2]
  assert med < 1e-2 and errs[0][0] < 0.5, (med, errs[:3])


def test_trainer_cuda_graph_matches_eager():
  """The captured-and-replayed step must reproduce the eagerly enqueued step.  Adam's first updates are sign-like, so
  runs diverge from atomics-order noise alone; the test therefore compares ONE step from an identical snapshot:
  two warm-up steps, snapshot (weights, moments, BN buffers, step counter), step 3 eagerly, restore, step 3 as a
  captured graph: loss, gradient buffer and updated weights must agree; then two more replays must keep training."""
  from corenet_b200.trainer import Trainer
  dev = t.device("cuda", 0)
  inp = MG.case_inputs("A")
  gt = MG.synthetic_gt(1, 2)
  for train_mode in (False, True):
    m = build_model().to(dev)
    m = m.train() if train_mode else m.eval()
    tr = Trainer(m, lr=4e-4, eps=1e-4, loss="iou_fgbg", use_graph=True)
    args = [inp["image"].to(dev), inp["v2s"].to(dev), inp["offsets"].to(dev), gt.to(dev)]
    l12 = [tr.step(*args).item() for _ in range(2)]
    assert tr.graph_launches == 0
    bufs = dict(m.named_buffers())
    snap = (tr.flat.clone(), tr.m.clone(), tr.v.clone(), tr.step_dev.clone(), {k: v.clone() for k, v in bufs.items()})

    def restore():
      tr.flat.copy_(snap[0]); tr.m.copy_(snap[1]); tr.v.copy_(snap[2]); tr.step_dev.copy_(snap[3])
      for k, v in snap[4].items():
        bufs[k].copy_(v)
      tr.eng._ver_sig = None

    tr.use_graph = False
    loss_e = tr.step(*args).item()
    grad_e, flat_e = tr.grad.clone(), tr.flat.clone()
    nbt_e = int(bufs["decoder.stage_6.b2.num_batches_tracked"])
    restore()
    tr.use_graph = True
    loss_g = tr.step(*args).item()                      # capture + first replay
    assert tr.graph_launches > 100 and int(tr.eng.tc_status) == 0
    grad_g, flat_g = tr.grad.clone(), tr.flat.clone()
    assert int(bufs["decoder.stage_6.b2.num_batches_tracked"]) == nbt_e == (3 if train_mode else 0)
    assert int(tr.step_dev) == 3
    gerr = ((grad_e - grad_g).double().norm() / grad_e.double().norm()).item()
    uerr = ((flat_e - flat_g).double().norm() / (flat_e - snap[0]).double().norm()).item()
    print(f"{'train' if train_mode else 'eval'}: loss eager {loss_e:.7f} graph {loss_g:.7f}  grad rel-L2 {gerr:.2e}  "
          f"update rel-L2 {uerr:.2e}")
    # The two runs of the SAME step differ only by the order of the split-K / flush atomics (~1e-7 of a layer's
    # output); train-mode BN at init amplifies that ~1000x (module docstring), so the single-step loss agrees to ~1e-4
    # there (observed 3e-5 .. 2.3e-4 over repeated runs) and to ~1e-7 in eval mode.
    assert abs(loss_e - loss_g) < (2e-3 if train_mode else 5e-6)
    assert gerr < (1e-1 if train_mode else 1e-4)
    more = [tr.step(*args).item() for _ in range(2)]    # replays
    # five Adam steps from a random init are noisy: require "no divergence" (the IoU loss would jump towards 1), the
    # exact-step comparison above is the real check
    assert all(np.isfinite(more)) and int(tr.step_dev) == 5 and min(more) < l12[0] + 5e-3
  # host inputs (pinned) go straight into the graph's static buffers
  host = [inp["image"].pin_memory(), inp["v2s"].pin_memory(), inp["offsets"].pin_memory(), gt.pin_memory()]
  loss_h = tr.step(*host)
  assert t.isfinite(loss_h).all()
  # prefetch path: H2D on the copy stream into staging buffers, consumed by an argument-less step()
  tr.prefetch(*host)
  l1 = tr.step().item()
  tr.prefetch(*host)
  l2 = tr.step().item()
  assert np.isfinite([l1, l2]).all() and int(tr.step_dev) == 8


def test_mean_iou_parity_and_inference_plug():
  """h7-style eval (B=2, C=2): forward -> softmax -> argmax -> confusion matrix -> mean IoU through the CUDA
  kernels vs the oracle (voxel_metrics.py:33-58, evaluation_results.py:262-266): |delta mIoU| <= 0.1 pt."""
  from corenet_b200 import ops
  from corenet_b200.super_resolution import super_resolution_from_model
  dev = t.device("cuda", 0)
  inp = MG.case_inputs("B")
  m = build_model()
  sd = {k: v.clone() for k, v in m.state_dict().items()}
  gt = MG.synthetic_gt(2, 2)
  lo = O.corenet_forward(dict(sd), inp["image"], inp["v2s"], inp["offsets"], False)
  cm_o = O.confusion_matrix(lo.argmax(1), gt, 2)
  m = m.to(dev).eval()
  with t.no_grad():
    logits = m(inp["image"].to(dev), inp["v2s"].to(dev), inp["offsets"].to(dev))
  cm = ops.argmax_confusion(logits, gt.to(dev)).cpu()
  assert int(cm.sum()) == gt.numel()
  assert abs(O.mean_iou(cm) - O.mean_iou(cm_o)) <= 1e-3           # 0.1 pt
  assert (cm - cm_o).abs().sum().item() <= 1e-5 * gt.numel()      # only near-tie voxels may differ
  # inference plug point at the native resolution = softmax of the model output
  sr = super_resolution_from_model(m)
  cam = t.eye(4, device=dev)[None].expand(2, 4, 4)
  pmf = sr(inp["image"].to(dev), inp["v2s"].to(dev), cam, inp["offsets"].to(dev), (128, 128, 128))
  assert rel_err(pmf, lo.softmax(1)) <= 1e-3


def test_device_resident_gt_pipeline():
  """batched_example.voxelize on the GPU (rasterise -> fill -> label merge) vs the oracles, two scenes with
  1 and 2 meshes (cubes), semantic labels."""
  from corenet_b200.data import batched_example as be
  from oracle import fill_voxels_oracle as FO
  from oracle import voxelize_oracle as VO
  from tests.conftest import cube_mesh
  res = (12, 12, 12)
  c1 = cube_mesh(0.99) / 3.0 * 0.5 + 0.1          # cubes inside the unit cube
  c2 = cube_mesh(0.99) / 3.0 * 0.4 + 0.45
  c3 = cube_mesh(0.99) / 3.0 * 0.3 + 0.05
  verts = t.from_numpy(np.concatenate([c1, c2, c3]))
  ntri = [t.tensor([12], dtype=t.int32), t.tensor([12, 12], dtype=t.int32)]
  offs = t.tensor([[0.5, 0.5, 0.5], [0.25, 0.5, 0.75]])
  labels = [[3], [2, 5]]
  v2x, grid = be.voxelize(verts, ntri, offs, res, be.VoxelContentSemanticLabel(labels), image_resolution_multiplier=8)
  assert grid.dtype == t.int32 and tuple(grid.shape) == (2, 12, 12, 12) and grid.is_cuda
  # oracle chain
  w2x = [O.translate(offs[b] - 0.5) @ O.scale([12.0] * 3) for b in range(2)]
  mesh_v2x = np.stack([w2x[0].numpy(), w2x[1].numpy(), w2x[1].numpy()])
  occ = VO.voxelize_mesh_oracle(verts.numpy(), [12, 12, 12], res, mesh_v2x, image_resolution_multiplier=8)
  occ = FO.fill_inside_voxels_oracle(occ)
  exp = np.stack([3 * occ[0], np.maximum(2 * occ[1], 5 * occ[2])]).astype(np.int32)
  np.testing.assert_array_equal(grid.cpu().numpy(), exp)
  assert exp[0].sum() > 0 and (exp[1] == 5).sum() > 0
