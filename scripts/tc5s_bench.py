"""Times crn_conv5_tcs (kz-stacked forward) against crn_conv5_tc on the stage_6.c1 shape (4 x 64^3, 28 -> 16)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
from corenet_b200 import _lib, ops
dev = t.device("cuda", 0)
n, cin, cout, g = 4, 28, 16, 64
w = t.randn(cout, cin, 5, 5, 5, device=dev) * 0.05
bias = t.randn(cout, device=dev)
x = t.randn(n * g ** 3, cin, device=dev)
y0 = t.zeros(n * g ** 3, cout, device=dev); y1 = t.zeros_like(y0)
d = ops.make_desc(n, cin, cout, (g, g, g), (g, g, g), (5, 5, 5), 1, 2, False, cin, cout)
st = _lib.stream_ptr(); status = t.zeros(1, dtype=t.int32, device=dev)
w0 = t.zeros(_lib.lib().crn_tc5_packed_floats(cin, cout), device=dev)
w1 = t.zeros(_lib.lib().crn_tc5s_packed_floats(cin), device=dev)
_lib.call("crn_tc5_pack", w.data_ptr(), cout, cin, 0, w0.data_ptr(), st)
_lib.call("crn_tc5s_pack", w.data_ptr(), cout, cin, w1.data_ptr(), st)
f0 = lambda: _lib.call("crn_conv5_tc", C.byref(d), 0, x.data_ptr(), w0.data_ptr(), bias.data_ptr(), y0.data_ptr(), status.data_ptr(), st)
f1 = lambda: _lib.call("crn_conv5_tcs", C.byref(d), x.data_ptr(), w1.data_ptr(), bias.data_ptr(), y1.data_ptr(), status.data_ptr(), st)
res = []
for f in (f0, f1):
  f(); t.cuda.synchronize()
  e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(3): f()
  e1.record(); t.cuda.synchronize()
  res.append(e0.elapsed_time(e1) / 3)
err = ((y0 - y1).abs().max() / y0.abs().max()).item()
macs = n * g ** 3 * 125 * cin * cout
print(f"6.c1 fwd: tc5 {res[0]:.3f} ms  kz-stacked {res[1]:.3f} ms ({2 * macs / res[1] / 1e9:.1f} TF/s)  err {err:.1e} status {int(status)}", flush=True)

def compare(name, n, cin, cout, g, kind):
  w = t.randn(cout, cin, 5, 5, 5, device=dev) * 0.05
  K, N = (cin, cout) if kind == 0 else (cout, cin)
  src = t.randn(n * g ** 3, K, device=dev)
  o0 = t.zeros(n * g ** 3, N, device=dev); o1 = t.zeros_like(o0)
  d = ops.make_desc(n, cin, cout, (g, g, g), (g, g, g), (5, 5, 5), 1, 2, False, cin, cout)
  w0 = t.zeros(_lib.lib().crn_tc5_packed_floats(K, N), device=dev)
  w1 = t.zeros(_lib.lib().crn_tc5s_packed_floats(K), device=dev)
  _lib.call("crn_tc5_pack", w.data_ptr(), cout, cin, kind, w0.data_ptr(), st)
  _lib.call("crn_tc5s_pack2", w.data_ptr(), cout, cin, kind, w1.data_ptr(), st)
  f0 = lambda: _lib.call("crn_conv5_tc", C.byref(d), kind, src.data_ptr(), w0.data_ptr(), None, o0.data_ptr(), status.data_ptr(), st)
  f1 = lambda: _lib.call("crn_conv5_tcs2", C.byref(d), kind, src.data_ptr(), w1.data_ptr(), None, o1.data_ptr(), status.data_ptr(), st)
  res = []
  for f in (f0, f1):
    f(); t.cuda.synchronize()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f()
    e1.record(); t.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 3)
  err = ((o0 - o1).abs().max() / o0.abs().max()).item()
  macs = n * g ** 3 * 125 * cin * cout
  print(f"{name}: tc5 {res[0]:.3f} ms  kz-stacked {res[1]:.3f} ms ({2 * macs / res[1] / 1e9:.1f} TF/s)  err {err:.1e} status {int(status)}", flush=True)

compare("6.c1 dgrad", 4, 28, 16, 64, 1)
compare("5.c1 fwd", 4, 56, 32, 32, 0)
