/* corenet_b200 -- DIAGNOSTIC entry points (not part of the drop-in boundary).
 *
 * Hardware bring-up probes that pin the tcgen05 descriptor layouts the kernels rely on, and a debug read-back.  They
 * are compiled into the library only with -DCRN_DIAG (corenet_b200/build.py passes it unless CRN_NO_DIAG=1 is set);
 * the test-suite uses them (tests/test_gpu_ops.py tc_probe cases), the product path never calls them.
 */
#ifndef CORENET_B200_DIAG_H_
#define CORENET_B200_DIAG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Bring-up / self-test of the tcgen05 (5th-gen tensor core) path: D[128,N] = A[128,K] * B[N,K]^T with
 * tf32 operands and an fp32 TMEM accumulator; mode 0 = single-pass TF32, 1 = 3xTF32 split.
 * status (device int) is set to 1 if the mbarrier wait timed out. */
int crn_tc_probe(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t mode,
                 int32_t* status, void* stream);
/* Same for MN-major operands (reduction index slow in memory, the weight-gradient case):
 * D[128,N] = sum_k A[k+shift][m] * B[k][n], A [K+4][128], B [K+4][N]; shift moves the A descriptor by whole rows. */
int crn_tc_probe_mn(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t mode, int32_t shift,
                    int32_t* status, void* stream);

/* Debug: with crn_set_flags bit 8 the kernel stamps a per-CTA clock64 timeline; copies n (<= 4096) int64 to host. */
int crn_gemm_tc_debug_read(long long* host_dst, int32_t n);
/* Same switch for conv_tc5s_kernel: per CTA 8 int64 = cycles of [MMA thread total, its waits on acc_empty / w_full /
 * plane_full, producer total, its wait on plane_empty, epilogue total, its wait on acc_full]; n <= 148 * 8. */
int crn_tc5s_debug_read(long long* host_dst, int32_t n);
/* ... and for wgrad_line_kernel (Conv3d k=5 weight gradient): [MMA thread total, its waits on full_x / full_y /
 * acc_empty, producer total, its waits on empty_x + empty_y, epilogue total, its wait on acc_full]. */
int crn_wgrad_line_debug_read(long long* host_dst, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* CORENET_B200_DIAG_H_ */
