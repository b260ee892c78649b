"""Per-CTA clock64 timeline of gemm_tc_kernel (debug flags) on one encoder-like shape."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch as t
from corenet_b200 import _lib, ops
dev = t.device("cuda", 0)
n, cin, cout, hw = 4, 64, 256, 64
w = t.randn(cout, cin, 1, 1, device=dev) * 0.05
wtc = ops.gemm_tc_pack([w], [0])[0]
x = t.randn(n * hw * hw, cin, device=dev); y = t.zeros(n * hw * hw, cout, device=dev); bias = t.randn(cout, device=dev)
d = ops.make_desc(n, cin, cout, (1, hw, hw), (1, hw, hw), (1, 1, 1), 1, 0, False, cin, cout)
st = _lib.stream_ptr(); status = t.zeros(1, dtype=t.int32, device=dev)
def run():
  _lib.call("crn_conv_gemm_tc", C.byref(d), 0, x.data_ptr(), wtc.data_ptr(), bias.data_ptr(), y.data_ptr(), 0, status.data_ptr(), st)
for dbg in (1, 1 | 2, 1 | 4, 1 | 8, 1 | 2 | 4 | 8):
  _lib.lib().crn_set_flags(dbg << 8)
  run(); t.cuda.synchronize()
  e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10): run()
  e1.record(); t.cuda.synchronize()
  buf = (C.c_longlong * 4096)()
  _lib.lib().crn_gemm_tc_debug_read(buf, 4096)
  a = np.array(buf[:256 * 8], dtype=np.int64).reshape(256, 8)
  rel = a - a[:, :1]
  med = np.median(rel, axis=0)
  print(f"dbg={dbg:2d}  kernel {e0.elapsed_time(e1) * 100:.1f} us | median cycles since CTA start: setup {med[1]:.0f} first_stage_ready {med[2]:.0f} "
        f"mma_done_issue {med[3]:.0f} epi_sums_done {med[4]:.0f} epi_stores_issued {med[5]:.0f} after_sync {med[6]:.0f} dealloc {med[7]:.0f}", flush=True)
_lib.lib().crn_set_flags(0)
