"""Where does wgrad_line_kernel (Conv3d k=5 weight gradient on tcgen05) spend its time?  Same accounting as
scripts/tc5s_waits.py (crn_set_flags bit 8, include/corenet_b200_diag.h) on the k5 layers of the B=4 training step.

  python scripts/wgrad_line_waits.py            # on a B200
"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch as t

from corenet_b200 import _lib, ops

dev = t.device("cuda", 0)
lib = _lib.lib()

SHAPES = [("stage_6.c1 28->16 @64^3", 4, 28, 16, 64), ("stage_5.c1 56->32 @32^3", 4, 56, 32, 32)]


def run(name, n, cin, cout, g):
  gen = t.Generator().manual_seed(1)
  x = t.randn(n * g ** 3, cin, generator=gen).to(dev)
  dy = t.randn(n * g ** 3, cout, generator=gen).to(dev)
  r4 = lambda c: (c + 3) // 4 * 4
  dw = t.zeros(125, r4(cin), r4(cout), device=dev)
  desc = ops.make_desc(n, cin, cout, (g, g, g), (g, g, g), (5, 5, 5), 1, 2, False, cin, cout)
  status = t.zeros(1, dtype=t.int32, device=dev)
  st = _lib.stream_ptr()
  call = lambda: _lib.call("crn_conv_wgrad_line", C.byref(desc), x.data_ptr(), dy.data_ptr(), dw.data_ptr(),
                           status.data_ptr(), st)
  lib.crn_set_flags(0)
  for _ in range(3):
    call()
  e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(5):
    call()
  e1.record()
  t.cuda.synchronize()
  us = e0.elapsed_time(e1) / 5 * 1e3
  lib.crn_set_flags(256)
  call()
  t.cuda.synchronize()
  lib.crn_set_flags(0)
  assert int(status) == 0
  buf = np.zeros(148 * 8, dtype=np.int64)
  assert lib.crn_wgrad_line_debug_read(buf.ctypes.data, buf.size) == 0
  d = np.median(buf.reshape(148, 8).astype(np.float64), axis=0)
  pct = lambda a, b: 100.0 * a / max(b, 1.0)
  print(f"{name} {us:8.1f} us | mma thread {d[0]:9.0f} cyc: wait full_x {pct(d[1], d[0]):5.1f}% full_y {pct(d[2], d[0]):5.1f}% "
        f"acc_empty {pct(d[3], d[0]):5.1f}% | producer {d[4]:9.0f} cyc: wait empty_x/y {pct(d[5], d[4]):5.1f}% | "
        f"epilogue {d[6]:9.0f} cyc: wait acc_full {pct(d[7], d[6]):5.1f}%")


for s in SHAPES:
  run(*s)
