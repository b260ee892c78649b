import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as t
from corenet_b200 import _lib
dev = t.device("cuda", 0)
t.manual_seed(0)
for shift in (0, 1, 2, 3, 4):
  for N in (16, 32, 64, 96, 128, 256):
    for K in (8, 32):
      A = t.randn(K + 4, 128, device=dev); B = t.randn(K + 4, N, device=dev)
      ref = (A[shift:shift + K].double().t() @ B[:K].double())
      for mode in (0, 1):
        D = t.full((128, N), float("nan"), device=dev)
        status = t.zeros(1, dtype=t.int32, device=dev)
        _lib.call("crn_tc_probe_mn", A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, mode, shift, status.data_ptr(), _lib.stream_ptr())
        t.cuda.synchronize()
        err = ((D.double() - ref).abs().max() / ref.abs().max()).item()
        print(f"shift={shift} N={N} K={K} mode={mode} status={int(status)} rel err {err:.3e}", flush=True)
