"""Where does conv_tc5s_kernel spend its time?  Runs the three stacked Conv3d k=5 launches of the B=4 training step with
crn_set_flags bit 8 (per-CTA wait-time accounting, include/corenet_b200_diag.h) in both precision modes and prints, per
role, the share of its life spent waiting on each barrier (median over the 148 CTAs).

  python scripts/tc5s_waits.py            # on a B200
"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch as t

from corenet_b200 import _lib, ops

dev = t.device("cuda", 0)
lib = _lib.lib()

SHAPES = [("stage_6.c1 fwd   28->16 @64^3", 4, 28, 16, 64, 0),
          ("stage_6.c1 dgrad 16->28 @64^3", 4, 28, 16, 64, 1),
          ("stage_5.c1 fwd   56->32 @32^3", 4, 56, 32, 32, 0)]


def run(name, n, cin, cout, g, kind, single, skip=False):
  K, N = (cin, cout) if kind == 0 else (cout, cin)
  gen = t.Generator().manual_seed(1)
  wt = (t.randn(cout, cin, 5, 5, 5, generator=gen) * 0.05).to(dev)
  xin = t.randn(n * g ** 3, K, generator=gen).to(dev)
  out = t.empty(n * g ** 3, N, device=dev)
  wtc = t.zeros(lib.crn_tc5s_packed_floats(K), device=dev)
  st = _lib.stream_ptr()
  _lib.call("crn_tc5s_pack2", wt.data_ptr(), cout, cin, kind, wtc.data_ptr(), st)
  desc = ops.make_desc(n, cin, cout, (g, g, g), (g, g, g), (5, 5, 5), 1, 2, False, cin, cout)
  status = t.zeros(1, dtype=t.int32, device=dev)
  flags = 256 | (8192 if single else 0) | (512 if skip else 0)
  call = lambda: _lib.call("crn_conv5_tcs2", C.byref(desc), kind, xin.data_ptr(), wtc.data_ptr(), None, out.data_ptr(),
                           status.data_ptr(), st)
  lib.crn_set_flags(flags & ~256)
  for _ in range(3):
    call()
  e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(5):
    call()
  e1.record()
  t.cuda.synchronize()
  us = e0.elapsed_time(e1) / 5 * 1e3
  lib.crn_set_flags(flags)
  call()
  t.cuda.synchronize()
  lib.crn_set_flags(0)
  assert int(status) == 0
  buf = np.zeros(148 * 8, dtype=np.int64)
  assert lib.crn_tc5s_debug_read(buf.ctypes.data, buf.size) == 0
  d = np.median(buf.reshape(148, 8).astype(np.float64), axis=0)
  pct = lambda a, b: 100.0 * a / max(b, 1.0)
  print(f"{name}  {'noMMA ' if skip else 'tf32  ' if single else '3xtf32'} {us:8.1f} us | mma thread {d[0]:9.0f} cyc: wait acc_empty {pct(d[1], d[0]):5.1f}% "
        f"w_full {pct(d[2], d[0]):5.1f}% plane_full {pct(d[3], d[0]):5.1f}% | producer {d[4]:9.0f} cyc: wait plane_empty "
        f"{pct(d[5], d[4]):5.1f}% | epilogue {d[6]:9.0f} cyc: wait acc_full {pct(d[7], d[6]):5.1f}%")


for s in SHAPES:
  for single in (False, True):
    run(*s, single)
  run(*s, False, skip=True)
