"""Times crn_brn_stats on typical encoder / decoder shapes (L2-hot and after an L2 flush)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
from corenet_b200 import _lib
dev = t.device("cuda", 0)
st = _lib.stream_ptr()
flush = t.empty(256 << 20, dtype=t.uint8, device=dev)
for rows, C in ((4096, 512), (16384, 256), (65536, 64), (1024, 1024), (262144, 56), (1048576, 28)):
  x = t.randn(rows, C, device=dev)
  acc = t.zeros(3 * C, dtype=t.float64, device=dev)
  call = lambda: _lib.call("crn_brn_stats", x.data_ptr(), rows, C, C, 0, 0, acc.data_ptr(), st)
  call(); t.cuda.synchronize()
  res = []
  for hot in (True, False):
    ts = []
    for _ in range(5):
      if not hot:
        flush.zero_()
      e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
      e0.record(); call(); e1.record(); t.cuda.synchronize()
      ts.append(e0.elapsed_time(e1) * 1e3)
    res.append(sorted(ts)[2])
  mb = rows * C * 4 / 1e6
  print(f"rows {rows:8d} C {C:5d}  {mb:7.1f} MB  hot {res[0]:7.1f} us ({mb / res[0] * 1e-3 * 1e3:6.2f} GB/ms)  cold {res[1]:7.1f} us", flush=True)
