"""Diagnostic: per-parameter gradient error of the CUDA path vs the oracle (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
from oracle import corenet_oracle as O, make_golden as MG
from corenet_b200 import configuration as C
from corenet_b200.model.core_net import CoreNet
from corenet_b200.model import losses

case, mode = sys.argv[1], sys.argv[2]
dev = t.device("cuda", 0)
t.manual_seed(0)
m = CoreNet(C.default_config(2))
sd = {k: v.clone() for k, v in m.state_dict().items()}
inp = MG.case_inputs(case)
if inp["perturb"]:
  sd = MG.perturb_brn(sd)
  m.load_state_dict(sd)
gt = MG.synthetic_gt(inp["image"].shape[0], 2)
st = {k: v.clone().requires_grad_(v.dtype == t.float32 and "running" not in k) for k, v in sd.items()}
taps = {}
lo = O.corenet_forward(st, inp["image"], inp["v2s"], inp["offsets"], mode == "train", {}, taps)
for v in taps.values():
  if v.requires_grad:
    v.retain_grad()
O.iou_fgbg(gt, lo).backward()
m = m.to(dev).train(mode == "train")
logits = m(inp["image"].to(dev), inp["v2s"].to(dev), inp["offsets"].to(dev))
losses.iou_fgbg(gt.to(dev), logits).backward()
rows = []
for n, p in m.named_parameters():
  go = st[n].grad
  sc = go.abs().max().item()
  e = (p.grad.cpu() - go).abs().max().item() / max(sc, 1e-30)
  rows.append((e, n, sc))
print("logits rel", ((logits.cpu() - lo.detach()).abs().max() / lo.detach().abs().max()).item())
for e, n, sc in rows:
  print(f"{e:.3e} {sc:.3e} {n}")
# activation gradients available from the plan
from corenet_b200 import engine
plan = engine.get_engine(m).plans[(inp["image"].shape[0], True)][0]
def cmp_act(name, buf, ref):
  b = ref.shape[0]
  sp = ref.shape[2:]
  got = buf.g.view(b, *sp, buf.cs)[..., :buf.C]
  perm = [0, len(sp) + 1] + list(range(1, len(sp) + 1))
  got = got.permute(perm).cpu()
  r = ref.grad
  print(f"ACT {name}: {((got - r).abs().max() / r.abs().max()).item():.3e}")
def fwd_cmp(name, buf, ref):
  b = ref.shape[0]
  sp = ref.shape[2:]
  got = buf.v.view(b, *sp, buf.cs)[..., :buf.C]
  perm = [0, len(sp) + 1] + list(range(1, len(sp) + 1))
  got = got.permute(perm).cpu()
  r = ref.detach()
  flips = ((got > 0) != (r > 0)).sum().item()
  print(f"FWD {name}: rel {((got - r).abs().max() / r.abs().max()).item():.3e} max {r.abs().max().item():.3e} "
        f"mask flips {flips} of {r.numel()}")
for blk in plan.blocks:
  cmp_act(blk["p"] + "out", blk["out"], taps[blk["p"] + "out"])
  for nm in ("a_y", "b_y", "out"):
    fwd_cmp(blk["p"] + nm, blk[nm], taps[blk["p"] + nm])
for sd_ in plan.stages:
  nm = f"stage_{sd_['stage']}.c1"
  cmp_act(nm, sd_["c"], taps[nm])
