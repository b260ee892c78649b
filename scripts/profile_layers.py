"""Runs single decoder layers through the C-ABI on random data (for ncu / timing).
usage: python scripts/profile_layers.py [layer ...]   layers: 6c1 6t1 5c1 5t1 4c1 4t1"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
from corenet_b200 import _lib, ops

LAYERS = {  # name: (cin, cout, k, grid_in, transposed)
    "6c1": (28, 16, 5, 64, False), "6t1": (16, 2, 7, 64, True), "5c1": (56, 32, 5, 32, False),
    "5t1": (32, 16, 7, 32, True), "4c1": (112, 64, 5, 16, False), "4t1": (64, 32, 7, 16, True),
    "3c1": (224, 128, 5, 8, False), "3t1": (128, 64, 7, 8, True)}


def run(name, B=4, iters=3, flags=0):
  cin, cout, k, g, tr = LAYERS[name]
  dev = t.device("cuda", 0)
  _lib.lib().crn_set_flags(flags)
  go = 2 * g if tr else g
  w = t.randn((cin, cout, k, k, k) if tr else (cout, cin, k, k, k), device=dev) * 0.05
  wf, wd, taps, cinp, coutp = ops.pack_weight(w, tr)
  r4 = lambda c: (c + 3) // 4 * 4
  x = t.randn(B * g ** 3, r4(cin), device=dev)
  y = t.zeros(B * go ** 3, r4(cout), device=dev)
  dy = t.randn(B * go ** 3, r4(cout), device=dev)
  dx = t.zeros_like(x)
  dw = t.zeros_like(wf)
  bias = t.randn(cout, device=dev)
  d = ops.make_desc(B, cin, cout, (g, g, g), (go, go, go), (k, k, k), 2 if tr else 1, k // 2, tr, r4(cin), r4(cout))
  st = _lib.stream_ptr()
  macs = B * g ** 3 * k ** 3 * cin * cout
  res = {}
  for kind, fn in (("fwd", lambda: _lib.call("crn_conv_fwd", C.byref(d), x.data_ptr(), wf.data_ptr(), bias.data_ptr(), y.data_ptr(), 0, st)),
                   ("dgrad", lambda: _lib.call("crn_conv_dgrad", C.byref(d), dy.data_ptr(), wd.data_ptr(), dx.data_ptr(), 0, st)),
                   ("wgrad", lambda: _lib.call("crn_conv_wgrad", C.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), st))):
    fn()
    t.cuda.synchronize()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
      fn()
    e1.record()
    t.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    res[kind] = (ms, 2 * macs / ms / 1e9)
  return res


if __name__ == "__main__":
  names = [a for a in sys.argv[1:] if a in LAYERS] or list(LAYERS)
  flags = int(os.environ.get("CRN_FLAGS", "0"))
  for n in names:
    r = run(n, flags=flags)
    print(n, "flags", flags, "  ".join(f"{k} {v[0]:.3f} ms {v[1]:.1f} TFLOP/s" for k, v in r.items()))
