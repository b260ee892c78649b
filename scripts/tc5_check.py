import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
import torch.nn.functional as F
from corenet_b200 import _lib, ops
dev = t.device("cuda", 0)
_lib.lib().crn_set_flags(int(os.environ.get("CRN_FLAGS", "0")))
r4 = lambda c: (c + 3) // 4 * 4

def run(n, cin, cout, d, h, w, kind, time_it=False):
  g = t.Generator().manual_seed(cin * 7 + cout)
  wt = (t.randn(cout, cin, 5, 5, 5, generator=g) * 0.05)
  bias = t.randn(cout, generator=g)
  if kind == 0:
    x = t.randn(n, cin, d, h, w, generator=g)
    ref = F.conv3d(x.double(), wt.double(), bias.double(), padding=2)
    K, N = cin, cout
  else:
    x = t.randn(n, cout, d, h, w, generator=g)      # dy
    ref = F.conv_transpose3d(x.double(), wt.double(), None, padding=2)   # dgrad of a stride-1 conv
    K, N = cout, cin
  xin = t.zeros(n * d * h * w, r4(K), device=dev)
  xin[:, :K] = x.permute(0, 2, 3, 4, 1).reshape(-1, K).to(dev)
  out = t.full((n * d * h * w, r4(N)), float("nan"), device=dev)
  nfl = _lib.lib().crn_tc5_packed_floats(K, N)
  wtc = t.zeros(nfl, device=dev)
  st = _lib.stream_ptr()
  _lib.call("crn_tc5_pack", wt.to(dev).contiguous().data_ptr(), cout, cin, kind, wtc.data_ptr(), st)
  desc = ops.make_desc(n, cin, cout, (d, h, w), (d, h, w), (5, 5, 5), 1, 2, False, r4(cin), r4(cout))
  status = t.zeros(1, dtype=t.int32, device=dev)
  b = bias.to(dev)
  _lib.call("crn_conv5_tc", C.byref(desc), kind, xin.data_ptr(), wtc.data_ptr(), b.data_ptr(), out.data_ptr(), status.data_ptr(), st)
  t.cuda.synchronize()
  got = out[:, :N].reshape(n, d, h, w, N).permute(0, 4, 1, 2, 3).cpu().double()
  err = ((got - ref).abs().max() / ref.abs().max()).item()
  msg = f"kind={kind} n={n} {cin}->{cout} grid {d}x{h}x{w}: status={int(status)} rel err {err:.3e}"
  if time_it:
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
      _lib.call("crn_conv5_tc", C.byref(desc), kind, xin.data_ptr(), wtc.data_ptr(), b.data_ptr(), out.data_ptr(), status.data_ptr(), st)
    e1.record(); t.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    macs = n * d * h * w * 125 * cin * cout
    msg += f"  {ms:.3f} ms  {2 * macs / ms / 1e9:.1f} TFLOP/s(useful)"
  print(msg, flush=True)

if __name__ == "__main__":
  which = sys.argv[1] if len(sys.argv) > 1 else "small"
  if which == "small":
    run(1, 8, 16, 8, 16, 8, 0)
    run(1, 8, 16, 8, 16, 8, 1)
    run(2, 28, 16, 16, 32, 16, 0)
    run(2, 28, 16, 16, 32, 16, 1)
    run(1, 56, 32, 8, 16, 16, 0)
    run(1, 56, 32, 8, 16, 16, 1)
  else:
    run(4, 28, 16, 64, 64, 64, 0, True)
    run(4, 28, 16, 64, 64, 64, 1, True)
    run(4, 56, 32, 32, 32, 32, 0, True)
    run(4, 56, 32, 32, 32, 32, 1, True)
    run(4, 112, 64, 16, 16, 16, 0, True)
