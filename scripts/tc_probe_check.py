import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
from corenet_b200 import _lib
dev = t.device("cuda", 0)
t.manual_seed(0)
for N in (16, 32):
  for K in (8, 32, 64):
    A = t.randn(128, K, device=dev); B = t.randn(N, K, device=dev)
    ref = (A.double() @ B.double().t())
    for mode in (0, 1):
      D = t.full((128, N), float("nan"), device=dev)
      status = t.zeros(1, dtype=t.int32, device=dev)
      _lib.call("crn_tc_probe", A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, mode, status.data_ptr(), _lib.stream_ptr())
      t.cuda.synchronize()
      err = ((D.double() - ref).abs().max() / ref.abs().max()).item()
      print(f"N={N} K={K} mode={mode} status={int(status)} rel err {err:.3e}", flush=True)
