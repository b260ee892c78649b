"""Times crn_conv_wgrad_line against the FFMA wgrad on the stage_6.c1 / stage_5.c1 shapes (B=4)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
from corenet_b200 import _lib, ops
dev = t.device("cuda", 0)
for name, n, cin, cout, g in (("6.c1", 4, 28, 16, 64), ("5.c1", 4, 56, 32, 32)):
  x = t.randn(n * g ** 3, cin, device=dev); dy = t.randn(n * g ** 3, cout, device=dev)
  d = ops.make_desc(n, cin, cout, (g, g, g), (g, g, g), (5, 5, 5), 1, 2, False, cin, cout)
  st = _lib.stream_ptr(); status = t.zeros(1, dtype=t.int32, device=dev)
  dw0 = t.zeros(125, cin, cout, device=dev); dw1 = t.zeros_like(dw0)
  f0 = lambda: _lib.call("crn_conv_wgrad", C.byref(d), x.data_ptr(), dy.data_ptr(), dw0.data_ptr(), st)
  f1 = lambda: _lib.call("crn_conv_wgrad_line", C.byref(d), x.data_ptr(), dy.data_ptr(), dw1.data_ptr(), status.data_ptr(), st)
  res = []
  for f in (f0, f1):
    f(); t.cuda.synchronize()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f()
    e1.record(); t.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 3)
  err = ((dw0 - dw1).abs().max() / dw0.abs().max()).item()
  macs = n * g ** 3 * 125 * cin * cout
  print(f"{name}: ffma {res[0]:.3f} ms  line-tc {res[1]:.3f} ms ({2 * macs / res[1] / 1e9:.1f} TF/s)  err {err:.1e} status {int(status)}", flush=True)
