"""Per-CTA clock64 timeline of gemm_tc_kernel (debug flags, -DCRN_DIAG build) on representative encoder shapes at B=4."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch as t
from corenet_b200 import _lib, ops
dev = t.device("cuda", 0)
SHAPES = [("stage2.op_a 1x1 256->64 @64^2", 256, 64, 64, 1), ("stage2.op_b 3x3 64->64 @64^2", 64, 64, 64, 3),
          ("stage3.op_b 3x3 128->128 @32^2", 128, 128, 32, 3), ("stage3.op_c 1x1 128->512 @32^2", 128, 512, 32, 1),
          ("stage4.op_a 1x1 1024->256 @16^2", 1024, 256, 16, 1), ("stage4.op_b 3x3 256->256 @16^2", 256, 256, 16, 3),
          ("stage4.op_c 1x1 256->1024 @16^2", 256, 1024, 16, 1), ("stage5.op_b 3x3 512->512 @8^2", 512, 512, 8, 3),
          ("stage5.op_c 1x1 512->2048 @8^2", 512, 2048, 8, 1)]
n = 4
for name, cin, cout, hw, k in SHAPES:
  w = t.randn(cout, cin, k, k, device=dev) * 0.05
  wtc = ops.gemm_tc_pack([w], [0])[0]
  x = t.randn(n * hw * hw, cin, device=dev); y = t.zeros(n * hw * hw, cout, device=dev); bias = t.randn(cout, device=dev)
  d = ops.make_desc(n, cin, cout, (1, hw, hw), (1, hw, hw), (1, k, k), 1, k // 2, False, cin, cout)
  st = _lib.stream_ptr(); status = t.zeros(1, dtype=t.int32, device=dev)
  def run():
    _lib.call("crn_conv_gemm_tc", C.byref(d), 0, x.data_ptr(), wtc.data_ptr(), bias.data_ptr(), y.data_ptr(), 0, status.data_ptr(), st)
  _lib.lib().crn_set_flags(0)
  run(); t.cuda.synchronize()
  e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(20): run()
  e1.record(); t.cuda.synchronize()
  us = e0.elapsed_time(e1) * 50
  _lib.lib().crn_set_flags(1 << 8)
  run(); t.cuda.synchronize()
  buf = (C.c_longlong * 4096)()
  _lib.lib().crn_gemm_tc_debug_read(buf, 4096)
  a = np.array(buf[:512 * 8], dtype=np.int64).reshape(512, 8)
  a = a[a[:, 0] != 0][:256]
  rel = a - a[:, :1]
  med = np.median(rel, axis=0)
  span = (a[:, 7].max() - a[:, 0].min())
  flops = 2.0 * n * hw * hw * cin * cout * k * k
  print(f"{name:34s} {us:6.1f} us/launch ({flops / us / 1e6:5.1f} TF) CTAs {len(a):3d} | cycles since CTA start: setup {med[1]:.0f} first_stage "
        f"{med[2]:.0f} mma_issued {med[3]:.0f} epi_sums {med[4]:.0f} stores_issued {med[5]:.0f} after_sync {med[6]:.0f} dealloc {med[7]:.0f} | "
        f"first CTA start -> last CTA end {span} cycles", flush=True)
_lib.lib().crn_set_flags(0)
