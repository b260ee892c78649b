"""ConvTranspose3d k7 s2 forward on the tcgen05 kernel vs torch fp64 (GPU box only)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
import torch.nn.functional as F
from corenet_b200 import _lib, ops
dev = t.device("cuda", 0)
_lib.lib().crn_set_flags(int(os.environ.get("CRN_FLAGS", "0")))
r4 = lambda c: (c + 3) // 4 * 4

def run(n, cin, cout, d, h, w, planar=False, time_it=False):
  g = t.Generator().manual_seed(cin * 7 + cout)
  wt = t.randn(cin, cout, 7, 7, 7, generator=g) * 0.05
  bias = t.randn(cout, generator=g)
  x = t.randn(n, cin, d, h, w, generator=g)
  ref = F.conv_transpose3d(x.double(), wt.double(), bias.double(), stride=2, padding=3, output_padding=1)
  xin = t.zeros(n * d * h * w, r4(cin), device=dev)
  xin[:, :cin] = x.permute(0, 2, 3, 4, 1).reshape(-1, cin).to(dev)
  S = 8 * d * h * w
  ycs = r4(cout) + 4                      # a wider row: the layer writes into a concat buffer
  out = t.full((n * cout * S,) if planar else (n * S, ycs), float("nan"), device=dev)
  wtc = t.zeros(_lib.lib().crn_tct_packed_floats(cin, cout, 0), device=dev)
  st = _lib.stream_ptr()
  _lib.call("crn_tct_pack", wt.to(dev).contiguous().data_ptr(), cin, cout, 0, wtc.data_ptr(), st)
  desc = ops.make_desc(n, cin, cout, (d, h, w), (2 * d, 2 * h, 2 * w), (7, 7, 7), 2, 3, True, r4(cin), ycs)
  desc.y_planar = int(planar)
  status = t.zeros(1, dtype=t.int32, device=dev)
  b = bias.to(dev)
  call = lambda: _lib.call("crn_convt7_tc", C.byref(desc), xin.data_ptr(), wtc.data_ptr(), b.data_ptr(), out.data_ptr(), status.data_ptr(), st)
  call()
  t.cuda.synchronize()
  if planar:
    got = out.reshape(n, cout, 2 * d, 2 * h, 2 * w).cpu().double()
  else:
    got = out[:, :cout].reshape(n, 2 * d, 2 * h, 2 * w, cout).permute(0, 4, 1, 2, 3).cpu().double()
  err = ((got - ref).abs().max() / ref.abs().max()).item()
  msg = f"convT n={n} {cin}->{cout} grid {d}x{h}x{w} planar={planar}: status={int(status)} rel err {err:.3e}"
  if time_it:
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
      call()
    e1.record(); t.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    macs = n * d * h * w * 343 * cin * cout
    msg += f"  {ms:.3f} ms  {2 * macs / ms / 1e9:.1f} TFLOP/s(useful)"
  print(msg, flush=True)

if __name__ == "__main__":
  which = sys.argv[1] if len(sys.argv) > 1 else "small"
  if which == "small":
    run(1, 8, 2, 8, 16, 8)
    run(1, 8, 2, 8, 16, 8, planar=True)
    run(2, 16, 2, 8, 32, 16, planar=True)
    run(1, 12, 3, 8, 16, 8, planar=True)
    run(1, 32, 16, 8, 16, 16)
    run(1, 16, 8, 8, 16, 8)
    run(1, 16, 4, 8, 16, 8)
  else:
    run(4, 16, 2, 64, 64, 64, planar=True, time_it=True)
    run(4, 32, 16, 32, 32, 32, time_it=True)


def run_dgrad(n, cin, cout, d, h, w, time_it=False):
  g = t.Generator().manual_seed(cin * 7 + cout + 1)
  wt = t.randn(cin, cout, 7, 7, 7, generator=g) * 0.05
  dy = t.randn(n, cout, 2 * d, 2 * h, 2 * w, generator=g)
  # dgrad of convT = strided conv of dy with the same weight
  ref = F.conv3d(dy.double(), wt.double(), None, stride=2, padding=3)
  ycs = r4(cout) + 4
  dyin = t.zeros(n * 8 * d * h * w, ycs, device=dev)
  dyin[:, :cout] = dy.permute(0, 2, 3, 4, 1).reshape(-1, cout).to(dev)
  out = t.full((n * d * h * w, r4(cin)), float("nan"), device=dev)
  wtc = t.zeros(_lib.lib().crn_tct_packed_floats(cin, cout, 1), device=dev)
  st = _lib.stream_ptr()
  _lib.call("crn_tct_pack", wt.to(dev).contiguous().data_ptr(), cin, cout, 1, wtc.data_ptr(), st)
  desc = ops.make_desc(n, cin, cout, (d, h, w), (2 * d, 2 * h, 2 * w), (7, 7, 7), 2, 3, True, r4(cin), ycs)
  status = t.zeros(1, dtype=t.int32, device=dev)
  call = lambda: _lib.call("crn_convt7_tc_dgrad", C.byref(desc), dyin.data_ptr(), wtc.data_ptr(), out.data_ptr(), status.data_ptr(), st)
  call()
  t.cuda.synchronize()
  got = out[:, :cin].reshape(n, d, h, w, cin).permute(0, 4, 1, 2, 3).cpu().double()
  err = ((got - ref).abs().max() / ref.abs().max()).item()
  msg = f"convT dgrad n={n} {cin}->{cout} grid {d}x{h}x{w}: status={int(status)} rel err {err:.3e}"
  if time_it:
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
      call()
    e1.record(); t.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    macs = n * d * h * w * 343 * cin * cout
    msg += f"  {ms:.3f} ms  {2 * macs / ms / 1e9:.1f} TFLOP/s(useful)"
  print(msg, flush=True)


if __name__ == "__main__":
  if which == "small":
    run_dgrad(1, 8, 4, 8, 16, 8)
    run_dgrad(1, 32, 16, 8, 16, 16)
    run_dgrad(2, 20, 8, 8, 16, 8)
    run_dgrad(1, 64, 32, 8, 16, 8)
  else:
    run_dgrad(4, 32, 16, 32, 32, 32, time_it=True)
    run_dgrad(4, 64, 32, 16, 16, 16, time_it=True)
