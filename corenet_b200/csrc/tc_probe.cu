// Bring-up probe for the tcgen05 path: D[128 x N] = A[128 x K] * B[N x K]^T, tf32 operands, fp32 TMEM
// accumulator.  mode 0: single-pass TF32 (operands truncated by the tensor core), mode 1: 3xTF32 split.
// Exercises exactly the pieces the conv kernel relies on: no-swizzle K-major smem descriptors, the
// instruction descriptor, TMEM alloc/ld, tcgen05.commit -> mbarrier.
#ifdef CRN_DIAG
#include "common.cuh"
#include "corenet_b200_diag.h"
#include "tc_common.cuh"

namespace {

template <int N>
__global__ void __launch_bounds__(128) tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                       float* __restrict__ D, int K, int mode, int* status) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int KC = K / 4;                                   // 16-byte K chunks
  const uint32_t a_lbo = 128 / 8 * 128, b_lbo = N / 8 * 128;   // bytes between K chunks
  float* a_hi = reinterpret_cast<float*>(smem_raw);
  float* a_lo = a_hi + 128 * K;
  float* b_hi = a_lo + 128 * K;
  float* b_lo = b_hi + N * K;

  if (warp == 0) tc::tmem_alloc(&tmem_slot, 32);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  // stage operands in the canonical no-swizzle K-major layout
  for (int i = tid; i < 128 * K; i += 128) {
    const int m = i / K, k = i % K;
    float hi = A[i], lo = 0.f;
    if (mode == 1) tc::split_tf32(A[i], hi, lo);
    const int off = (k / 4) * (a_lbo / 4) + (m / 8) * 32 + (m % 8) * 4 + (k % 4);
    a_hi[off] = hi; a_lo[off] = lo;
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    float hi = B[i], lo = 0.f;
    if (mode == 1) tc::split_tf32(B[i], hi, lo);
    const int off = (k / 4) * (b_lbo / 4) + (n / 8) * 32 + (n % 8) * 4 + (k % 4);
    b_hi[off] = hi; b_lo[off] = lo;
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t idesc = tc::make_idesc_tf32(128, N, 0, 0);
  if (tid == 0) {
    uint32_t acc = 0;
    const int passes = mode == 1 ? 3 : 1;
    for (int p = 0; p < passes; ++p) {
      const float* as = (p == 1) ? a_lo : a_hi;
      const float* bs = (p == 2) ? b_lo : b_hi;
      for (int j = 0; j < K / 8; ++j) {
        const uint64_t da = tc::make_desc(tc::smem_u32(as) + j * 2 * a_lbo, a_lbo, 128);
        const uint64_t db = tc::make_desc(tc::smem_u32(bs) + j * 2 * b_lbo, b_lbo, 128);
        tc::mma_tf32(tmem, da, db, idesc, acc);
        acc = 1;
      }
    }
    tc::commit(&bar);
  }
  const bool ok = tc::mbar_wait(&bar, 0);
  if (!ok && tid == 0) *status = 1;
  tc::fence_after_sync();
  float v[16];
  for (int c0 = 0; c0 < N; c0 += 16) {
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int c = 0; c < 16; ++c) D[(warp * 32 + (tid & 31)) * N + c0 + c] = v[c];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 32);
}
}  // namespace

// A [128][K], B [N][K], D [128][N] row-major device fp32; N in {16, 32}; K multiple of 8, <= 64.
extern "C" int crn_tc_probe(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t mode,
                            int32_t* status, void* stream) {
  CRN_REQUIRE(A && B && D && status && (N == 16 || N == 32) && K % 8 == 0 && K >= 8 && K <= 64, "crn_tc_probe: bad args");
  const size_t smem = sizeof(float) * 2 * (128 * K + N * K);
  cudaStream_t st = crn_stream(stream);
  if (N == 16) tc_probe_kernel<16><<<1, 128, smem, st>>>(A, B, D, K, mode, status);
  else tc_probe_kernel<32><<<1, 128, smem, st>>>(A, B, D, K, mode, status);
  CRN_LAUNCH_CHECK("tc_probe");
  return CRN_OK;
}

// ---------------------------------------------------------------------------------------------
// MN-major probe (the layout the weight-gradient kernels need: the reduction index -- rows/voxels -- is the
// slow index in memory, channels are contiguous):  D[128 x N] = sum_k A[k + shift][m] * B[k][n],
// A [K+4][128], B [K+4][N] row-major.  tf32 MN-major operands exist only in the "128B swizzle, 32B base" layout
// (descriptor layout type 1):  element (m, k) of a tile lives at
//   (m/32)*LBO + (k/4)*SBO + (k%4)*128 + ((((m%32)/8) ^ (k%4))*32 + (m%8)*4      [bytes]
// i.e. one reduction row = 128 B = 32 channels, 32-byte chunks XOR-swizzled by the row index; with SBO = 512 the
// rows of one 32-channel block are simply consecutive 128-byte lines, and `shift` moves the A start address by
// whole rows (the tap shift of the weight-gradient kernels) -- the swizzle is a function of the absolute address.
namespace {
template <int N>
__global__ void __launch_bounds__(128) tc_probe_mn_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ D, int K, int mode, int shift,
                                                          int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int R = K + 4;                                   // staged rows
  const uint32_t blk = (uint32_t)((R * 128 + 511) / 512 * 512);   // bytes per 32-channel block (LBO)
  constexpr int NB = (N + 31) / 32;
  uint8_t* a_hi = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms: 512 B aligned
  uint8_t* a_lo = a_hi + 4 * blk;
  uint8_t* b_hi = a_lo + 4 * blk;
  uint8_t* b_lo = b_hi + NB * blk;
  constexpr int TC = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  if (warp == 0) tc::tmem_alloc(&tmem_slot, TC);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  auto sw_off = [&](int m, int k) -> uint32_t {
    return (uint32_t)(m / 32) * blk + (uint32_t)k * 128 + (uint32_t)((((m % 32) / 8) ^ (k & 3)) * 32 + (m % 8) * 4);
  };
  for (int i = tid; i < 128 * R; i += 128) {
    const int k = i / 128, m = i % 128;
    float hi = A[i], lo = 0.f;
    if (mode == 1) tc::split_tf32(A[i], hi, lo);
    *reinterpret_cast<float*>(a_hi + sw_off(m, k)) = hi;
    *reinterpret_cast<float*>(a_lo + sw_off(m, k)) = lo;
  }
  for (int i = tid; i < N * R; i += 128) {
    const int k = i / N, n = i % N;
    float hi = B[i], lo = 0.f;
    if (mode == 1) tc::split_tf32(B[i], hi, lo);
    *reinterpret_cast<float*>(b_hi + sw_off(n, k)) = hi;
    *reinterpret_cast<float*>(b_lo + sw_off(n, k)) = lo;
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  constexpr uint32_t idesc = tc::make_idesc_tf32(128, N, 1, 1);
  if (tid == 0) {
    uint32_t acc = 0;
    const int passes = mode == 1 ? 3 : 1;
    for (int p = 0; p < passes; ++p) {
      const uint8_t* as = (p == 1) ? a_lo : a_hi;
      const uint8_t* bs = (p == 2) ? b_lo : b_hi;
      for (int j = 0; j < K / 8; ++j) {
        const uint64_t da = tc::make_desc_mn32(tc::smem_u32(as) + (j * 8 + shift) * 128, blk, 512);
        const uint64_t db = tc::make_desc_mn32(tc::smem_u32(bs) + j * 8 * 128, blk, 512);
        tc::mma_tf32(tmem, da, db, idesc, acc);
        acc = 1;
      }
    }
    tc::commit(&bar);
  }
  const bool ok = tc::mbar_wait(&bar, 0);
  if (!ok && tid == 0) *status = 1;
  tc::fence_after_sync();
  float v[16];
  for (int c0 = 0; c0 < N; c0 += 16) {
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int c = 0; c < 16; ++c) D[(warp * 32 + (tid & 31)) * N + c0 + c] = v[c];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, TC);
}
}  // namespace

// A [K+4][128], B [K+4][N], D [128][N] row-major device fp32; K multiple of 8, <= 32; 0 <= shift <= 4.
extern "C" int crn_tc_probe_mn(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t mode,
                               int32_t shift, int32_t* status, void* stream) {
  CRN_REQUIRE(A && B && D && status && K % 8 == 0 && K >= 8 && K <= 32 && shift >= 0 && shift <= 4,
              "crn_tc_probe_mn: bad args");
  const size_t blk = (size_t)((K + 4) * 128 + 511) / 512 * 512;
  const size_t smem = 2 * (4 + (N + 31) / 32) * blk + 1024;
  cudaStream_t st = crn_stream(stream);
#define PROBE_MN(NN)                                                                                          \
  case NN:                                                                                                    \
    cudaFuncSetAttribute(tc_probe_mn_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
    tc_probe_mn_kernel<NN><<<1, 128, smem, st>>>(A, B, D, K, mode, shift, status);                            \
    break;
  switch (N) {
    PROBE_MN(16) PROBE_MN(32) PROBE_MN(64) PROBE_MN(96) PROBE_MN(128) PROBE_MN(256)
    default: crn_set_error("crn_tc_probe_mn: unsupported N"); return CRN_ERR_BAD_ARG;
  }
#undef PROBE_MN
  CRN_LAUNCH_CHECK("tc_probe_mn");
  return CRN_OK;
}
#endif  // CRN_DIAG
