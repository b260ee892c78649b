"""Times crn_conv_gemm_tc against the FFMA kernels on the encoder / coarse-decoder layer shapes (B=4)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
from corenet_b200 import _lib, ops
dev = t.device("cuda", 0)
SHAPES = [  # name, N, cin, cout, idims, odims, k, stride, pad
  ("s2.1x1_64_256", 4, 64, 256, (1, 64, 64), (1, 64, 64), (1, 1, 1), 1, 0),
  ("s2.3x3_64", 4, 64, 64, (1, 64, 64), (1, 64, 64), (1, 3, 3), 1, 1),
  ("s3.1x1_512_128", 4, 512, 128, (1, 32, 32), (1, 32, 32), (1, 1, 1), 1, 0),
  ("s3.3x3_128", 4, 128, 128, (1, 32, 32), (1, 32, 32), (1, 3, 3), 1, 1),
  ("s4.1x1_256_1024", 4, 256, 1024, (1, 16, 16), (1, 16, 16), (1, 1, 1), 1, 0),
  ("s4.3x3_256", 4, 256, 256, (1, 16, 16), (1, 16, 16), (1, 3, 3), 1, 1),
  ("s5.1x1_512_2048", 4, 512, 2048, (1, 8, 8), (1, 8, 8), (1, 1, 1), 1, 0),
  ("s5.3x3_512", 4, 512, 512, (1, 8, 8), (1, 8, 8), (1, 3, 3), 1, 1),
  ("s4.sc_512_1024_s2", 4, 512, 1024, (1, 32, 32), (1, 16, 16), (1, 1, 1), 2, 0),
  ("dec4.c1_112_64", 4, 112, 64, (16, 16, 16), (16, 16, 16), (5, 5, 5), 1, 2),
  ("dec3.c1_224_128", 4, 224, 128, (8, 8, 8), (8, 8, 8), (5, 5, 5), 1, 2),
]
def timeit(fn, iters=5):
  fn(); t.cuda.synchronize()
  e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters): fn()
  e1.record(); t.cuda.synchronize()
  return e0.elapsed_time(e1) / iters
ONLY = [a for a in sys.argv[1:]]
for name, n, cin, cout, idims, odims, k, stride, pad in SHAPES:
  if ONLY and name not in ONLY: continue
  w = t.randn((cout, cin) + k, device=dev) * 0.05
  wf, wd, taps, cinp, coutp = ops.pack_weight(w, False)
  wtc_f, wtc_d = ops.gemm_tc_pack([w, w], [0, 1])
  ri = n * idims[0] * idims[1] * idims[2]; ro = n * odims[0] * odims[1] * odims[2]
  x = t.randn(ri, cin, device=dev); dy = t.randn(ro, cout, device=dev)
  y0 = t.zeros(ro, cout, device=dev); y1 = t.zeros(ro, cout, device=dev)
  dx0 = t.zeros(ri, cin, device=dev); dx1 = t.zeros(ri, cin, device=dev)
  bias = t.randn(cout, device=dev)
  d = ops.make_desc(n, cin, cout, idims, odims, k, stride, pad, False, cin, cout)
  st = _lib.stream_ptr(); status = t.zeros(1, dtype=t.int32, device=dev)
  macs = ro * k[0] * k[1] * k[2] * cin * cout
  f0 = lambda: _lib.call("crn_conv_fwd", C.byref(d), x.data_ptr(), wf.data_ptr(), bias.data_ptr(), y0.data_ptr(), 0, st)
  f1 = lambda: _lib.call("crn_conv_gemm_tc", C.byref(d), 0, x.data_ptr(), wtc_f.data_ptr(), bias.data_ptr(), y1.data_ptr(), 0, status.data_ptr(), st)
  a, b = timeit(f0), timeit(f1)
  err = ((y0 - y1).abs().max() / y0.abs().max()).item()
  msg = f"{name:22s} fwd ffma {a*1e3:8.1f} us  tc {b*1e3:8.1f} us ({2*macs/b/1e9:7.1f} TF/s) err {err:.1e}"
  if stride == 1:
    g0 = lambda: _lib.call("crn_conv_dgrad", C.byref(d), dy.data_ptr(), wd.data_ptr(), dx0.data_ptr(), 0, st)
    g1 = lambda: _lib.call("crn_conv_gemm_tc", C.byref(d), 1, dy.data_ptr(), wtc_d.data_ptr(), None, dx1.data_ptr(), 0, status.data_ptr(), st)
    a, b = timeit(g0), timeit(g1)
    err = ((dx0 - dx1).abs().max() / dx0.abs().max()).item()
    msg += f" | dgrad ffma {a*1e3:8.1f} us  tc {b*1e3:8.1f} us ({2*macs/b/1e9:7.1f} TF/s) err {err:.1e}"
  dw0 = t.zeros(k[0] * k[1] * k[2], cin, cout, device=dev); dw1 = t.zeros_like(dw0)
  h0 = lambda: _lib.call("crn_conv_wgrad", C.byref(d), x.data_ptr(), dy.data_ptr(), dw0.data_ptr(), st)
  h1 = lambda: _lib.call("crn_conv_wgrad_tc", C.byref(d), x.data_ptr(), dy.data_ptr(), dw1.data_ptr(), status.data_ptr(), st)
  a, b = timeit(h0), timeit(h1)
  err = ((dw0 - dw1).abs().max() / dw0.abs().max()).item()
  msg += f" | wgrad ffma {a*1e3:8.1f} us  tc {b*1e3:8.1f} us ({2*macs/b/1e9:7.1f} TF/s) err {err:.1e}"
  print(msg, "status", int(status), flush=True)
