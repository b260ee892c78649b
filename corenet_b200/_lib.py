"""ctypes binding of libcorenet_b200.so (the C-ABI declared in include/corenet_b200.h).

The product path has NO CPU fallback: if the library cannot be loaded, or a
tensor is not on a CUDA device, the ops raise.
"""
import ctypes as C
import os
import threading

import torch as t

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcorenet_b200.so")

_lib = None
_lock = threading.Lock()

i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class ConvDesc(C.Structure):
  """Mirror of crn_conv_desc."""
  _fields_ = [(n, i32) for n in (
      "N", "Cin", "Cout", "iD", "iH", "iW", "oD", "oH", "oW", "kD", "kH", "kW", "stride", "pad",
      "transposed", "x_cs", "x_co", "y_cs", "y_co", "CinP", "CoutP", "y_planar", "bias_n_stride")]


class PackItem(C.Structure):
  """Mirror of crn_pack_item."""
  _fields_ = [("src", vp), ("dst_fwd", vp), ("dst_dgrad", vp), ("Cin", i32), ("Cout", i32),
              ("taps", i32), ("CinP", i32), ("CoutP", i32), ("src_is_transposed", i32)]


class UnpackItem(C.Structure):
  """Mirror of crn_unpack_item."""
  _fields_ = [("src_packed", vp), ("dst", vp), ("Cin", i32), ("Cout", i32), ("taps", i32),
              ("CinP", i32), ("CoutP", i32), ("dst_is_transposed", i32)]


class F64CopyItem(C.Structure):
  """Mirror of crn_f64_copy_item."""
  _fields_ = [("src", vp), ("dst", vp)]


class GemmTcPackItem(C.Structure):
  """Mirror of crn_gemm_tc_pack_item."""
  _fields_ = [("src", vp), ("dst", vp), ("Cout", i32), ("Cin", i32), ("taps", i32), ("dgrad", i32)]


_P = C.POINTER
_SIGS = {
    "crn_version": ([], i32),
    "crn_build_arch": ([], C.c_char_p),
    "crn_last_error": ([], C.c_char_p),
    "crn_launch_count": ([], i64),
    "crn_set_flags": ([i32], None),
    "crn_pack_weights": ([vp, vp, i32, i64, vp], i32),
    "crn_unpack_wgrads": ([vp, vp, i32, i64, vp], i32),
    "crn_conv_fwd": ([_P(ConvDesc), vp, vp, vp, vp, i32, vp], i32),
    "crn_conv_dgrad": ([_P(ConvDesc), vp, vp, vp, i32, vp], i32),
    "crn_conv_wgrad": ([_P(ConvDesc), vp, vp, vp, vp], i32),
    "crn_brn_stats": ([vp, i64, i32, i32, i32, i32, vp, vp], i32),
    "crn_brn_finalize": ([vp, i64, i32, vp, vp, vp, vp, vp, f32, f32, i32, vp, vp], i32),
    "crn_brn_apply": ([vp, i64, i32, i32, i32, vp, vp, i32, i32, vp, i32, i32, vp, vp], i32),
    "crn_brn_bwd_reduce": ([vp, i32, i32, vp, vp, vp, i32, i32, i64, i32, vp, i32, i32, vp, vp, vp], i32),
    "crn_brn_bwd_dx": ([vp, i32, i32, vp, i32, i32, i64, i32, vp, vp, vp, i32, i32, vp, i32, i32, i32,
                        vp, vp, vp, vp], i32),
    "crn_brn_fused_supported": ([i64, i32], i32),
    "crn_brn_nbt_snapshot": ([vp, i32, vp, vp], i32),
    "crn_brn_fwd_fused": ([vp, i64, i32, i32, i32, i32, vp, vp, vp, vp, vp, f32, f32, i32, vp, i32, vp, i32, i32, vp, vp,
                           vp], i32),
    "crn_brn_bwd_fused": ([vp, i32, i32, vp, vp, vp, i32, i32, i64, i32, vp, i32, i32, i32, vp, vp, i32, i32, i32, vp, vp,
                           vp, vp], i32),
    "crn_preprocess_image": ([vp, i32, i32, i32, vp, vp], i32),
    "crn_maxpool_fwd": ([vp, i32, i32, i32, i32, vp, vp, vp], i32),
    "crn_maxpool_bwd": ([vp, vp, i32, i32, i32, i32, vp, vp], i32),
    "crn_spatial_mean_fwd": ([vp, i32, i32, i32, vp, vp], i32),
    "crn_spatial_mean_bwd": ([vp, i32, i32, i32, vp, i32, vp], i32),
    "crn_colsum": ([vp, i64, i32, i32, i32, vp, vp, vp], i32),
    "crn_colsum_planar": ([vp, i32, i32, i64, vp, vp, vp], i32),
    "crn_planar_to_rows": ([vp, i32, i32, i64, i32, vp, vp], i32),
    "crn_skip_sample_fwd": ([vp, i32, i32, i32, i32, i32, vp, vp, i32, i32, i32, vp, i32, i32, vp], i32),
    "crn_skip_sample_bwd": ([vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, i32, i32, vp, vp], i32),
    "crn_skip_lists_workspace_bytes": ([i32, i32, i32, i32, i32, i32], i64),
    "crn_skip_build_lists": ([i32, i32, i32, vp, vp, i32, i32, i32, vp, i64, vp, vp, vp], i32),
    "crn_skip_sample_bwd_sorted": ([vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp], i32),
    "crn_skip_indices": ([i32, i32, i32, vp, vp, i32, i32, i32, vp, vp], i32),
    "crn_loss_sums": ([vp, vp, i32, i32, i32, i64, i32, vp, vp], i32),
    "crn_loss_finalize": ([vp, i32, i32, i64, i32, vp, vp, vp], i32),
    "crn_loss_bwd": ([vp, vp, i32, i32, i32, i64, i32, vp, vp, vp, vp], i32),
    "crn_softmax_planar": ([vp, i32, i32, i64, vp, vp], i32),
    "crn_argmax_confusion": ([vp, vp, i32, i32, i32, i64, vp, vp], i32),
    "crn_argmax_confusion_labeled": ([vp, vp, i32, i32, i32, i64, vp, i32, vp, vp], i32),
    "crn_loss_sums_l": ([vp, i32, vp, i32, i32, i32, i64, i32, vp, vp], i32),
    "crn_loss_bwd_l": ([vp, i32, vp, i32, i32, i32, i64, i32, vp, vp, vp, i32, vp], i32),
    "crn_softmax_l": ([vp, i32, i32, i32, i64, vp, vp], i32),
    "crn_argmax_confusion_l": ([vp, i32, vp, i32, i32, i32, i64, vp, i32, vp, vp], i32),
    "crn_rows_to_planar": ([vp, i32, i32, i64, i32, vp, vp], i32),
    "crn_fill_workspace_bytes": ([i32, i32, i32, i32], i64),
    "crn_fill_inside": ([vp, vp, i32, i32, i32, i32, i32, i32, vp, vp], i32),
    "crn_voxelize_mesh": ([vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp], i32),
    "crn_merge_mesh_grids": ([vp, vp, vp, i32, i64, vp, vp], i32),
    "crn_tc_probe": ([vp, vp, vp, i32, i32, i32, vp, vp], i32),
    "crn_tc_probe_mn": ([vp, vp, vp, i32, i32, i32, i32, vp, vp], i32),
    "crn_tc5_packed_floats": ([i32, i32], i64),
    "crn_tc5_pack": ([vp, i32, i32, i32, vp, vp], i32),
    "crn_conv5_tc": ([_P(ConvDesc), i32, vp, vp, vp, vp, vp, vp], i32),
    "crn_tc5s_packed_floats": ([i32], i64),
    "crn_tc5s_pack": ([vp, i32, i32, vp, vp], i32),
    "crn_conv5_tcs": ([_P(ConvDesc), vp, vp, vp, vp, vp, vp], i32),
    "crn_tc5s_pack2": ([vp, i32, i32, i32, vp, vp], i32),
    "crn_conv5_tcs2": ([_P(ConvDesc), i32, vp, vp, vp, vp, vp, vp], i32),
    "crn_tct_packed_floats": ([i32, i32, i32], i64),
    "crn_tct_pack": ([vp, i32, i32, i32, vp, vp], i32),
    "crn_convt7_tc_dgrad": ([_P(ConvDesc), vp, vp, vp, vp, vp], i32),
    "crn_convt7_tc": ([_P(ConvDesc), vp, vp, vp, vp, vp, vp], i32),
    "crn_tctsf_packed_floats": ([i32], i64),
    "crn_tctsf_pack": ([vp, i32, i32, vp, vp], i32),
    "crn_convt7_tcs_fwd": ([_P(ConvDesc), vp, vp, vp, vp, vp, vp], i32),
    "crn_tcts_packed_floats": ([i32], i64),
    "crn_tcts_pack": ([vp, i32, i32, vp, vp], i32),
    "crn_convt7_tcs_dgrad": ([_P(ConvDesc), vp, vp, vp, vp, vp], i32),
    "crn_gemm_tc_packed_floats": ([i32, i32, i32], i64),
    "crn_gemm_tc_pack": ([vp, vp, i32, i64, vp], i32),
    "crn_conv_gemm_tc": ([_P(ConvDesc), i32, vp, vp, vp, vp, i32, vp, vp], i32),
    "crn_gemm_tc_debug_read": ([vp, i32], i32),
    "crn_tc5s_debug_read": ([vp, i32], i32),
    "crn_wgrad_line_debug_read": ([vp, i32], i32),
    "crn_conv_wgrad_tc": ([_P(ConvDesc), vp, vp, vp, vp, vp], i32),
    "crn_conv_wgrad_line_supported": ([_P(ConvDesc)], i32),
    "crn_conv_wgrad_line": ([_P(ConvDesc), vp, vp, vp, vp, vp], i32),
    "crn_conv_wgrad_xline_supported": ([_P(ConvDesc)], i32),
    "crn_conv_wgrad_xline": ([_P(ConvDesc), vp, vp, vp, vp, vp], i32),
    "crn_convt7_wgrad_line_supported": ([_P(ConvDesc)], i32),
    "crn_convt7_wgrad_line": ([_P(ConvDesc), vp, vp, vp, vp, vp], i32),
    "crn_gather_f64_to_f32": ([vp, vp, i32, i64, vp], i32),
    "crn_adam_step_dev": ([vp, vp, vp, vp, i64, f32, f32, f32, f32, vp, f32, vp], i32),
    "crn_adam_step_guarded": ([vp, vp, vp, vp, i64, f32, f32, f32, f32, vp, i32, f32, vp, vp], i32),
    "crn_status_poison": ([vp, vp, i64, i32, vp], i32),
    "crn_adam_step": ([vp, vp, vp, vp, i64, f32, f32, f32, f32, i32, f32, vp], i32),
}

EXPORTS = tuple(_SIGS)


def lib() -> C.CDLL:
  """Loads (once) and returns the C-ABI library; raises if it is missing."""
  global _lib
  if _lib is None:
    with _lock:
      if _lib is None:
        if not os.path.exists(LIB_PATH):
          raise RuntimeError(
              f"{LIB_PATH} is missing: build it with `python -m corenet_b200.build` "
              "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGS.items():
          fn = getattr(l, name)
          fn.argtypes = argtypes
          fn.restype = restype
        _lib = l
  return _lib


def check(status: int, what: str = ""):
  if status != 0:
    msg = lib().crn_last_error().decode()
    raise ValueError(f"corenet_b200 {what} failed (status {status}): {msg}")


def stream_ptr() -> int:
  """The current torch CUDA stream as a raw cudaStream_t."""
  return t.cuda.current_stream().cuda_stream


def ptr(x) -> int:
  """Device pointer of a CUDA tensor (None -> NULL)."""
  if x is None:
    return None
  if not x.is_cuda:
    raise ValueError("corenet_b200 ops need CUDA tensors (there is no CPU fallback)")
  return x.data_ptr()


def call(name: str, *args):
  check(getattr(lib(), name)(*args), name)
