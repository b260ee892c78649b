// Fused losses / eval reductions over the logits.
// Replaces model/losses.py:19-61 (iou_agnostic), :64-114 (iou_fgbg), :117-141
// (xent), :144-160 (xent_times_iou_agnostic) and, for eval,
// evaluation_results.py:40-51 + voxel_metrics.py:33-58 (argmax -> confusion).
// One pass over the logits forward, one pass backward (HBM-bound).
//
// Two logits layouts (crn_logits_layout in the *_l entry points):
//   planar  [N, C, S]      the reference's NCDHW (S = D*H*W): element (n, c, i) at (n*C + c)*S + i
//   rows    [N*S, CP]      channels-last rows as the last transposed convolution's tcgen05 epilogue writes them
//                          (CP = C rounded up to 4, pad channels ignored on read and written as 0 on write)
// The class count is a template parameter for the configurations of the reference (2 = FG_BG, 15 = SEMANTIC with
// the 14 ShapeNet classes) so that the per-voxel softmax stays in registers; other counts use the generic kernel.
#include "common.cuh"

namespace {
constexpr int NT = 256;
constexpr int MAXC = 32;

// labels outside [0, C) would index logits out of bounds: clamped here (the reference's one_hot raises on them,
// model/losses.py:36; the host wrappers validate the dtype, range validation needs a device sync and is opt-in)
template <typename GT>
__device__ __forceinline__ int load_gt(const void* gt, int64_t i, int C) {
  const GT v = reinterpret_cast<const GT*>(gt)[i];
  return v < (GT)0 ? 0 : (v >= (GT)C ? C - 1 : (int)v);
}

struct Layout {
  int rows;      // 0 planar, 1 rows
  int CP;        // row pitch (floats) in rows mode
};

// CT > 0: compile-time class count (arrays fully unrolled -> registers); CT == 0: runtime C <= MAXC
template <int CT>
struct Cls {
  static constexpr int CAP = CT > 0 ? CT : MAXC;
  static __device__ __forceinline__ int n(int C) { return CT > 0 ? CT : C; }
};

template <int CT>
__device__ __forceinline__ void load_logits(const float* __restrict__ base, int C, int64_t S, int64_t i, Layout L,
                                            float (&s)[Cls<CT>::CAP]) {
  const int n = Cls<CT>::n(C);
  if (L.rows) {
    const float4* row = reinterpret_cast<const float4*>(base + i * L.CP);
#pragma unroll
    for (int q = 0; q < (Cls<CT>::CAP + 3) / 4; ++q) {
      if (q * 4 < n) {
        const float4 v = __ldg(row + q);
        if (q * 4 + 0 < Cls<CT>::CAP) s[q * 4 + 0] = v.x;
        if (q * 4 + 1 < Cls<CT>::CAP) s[q * 4 + 1] = v.y;
        if (q * 4 + 2 < Cls<CT>::CAP) s[q * 4 + 2] = v.z;
        if (q * 4 + 3 < Cls<CT>::CAP) s[q * 4 + 3] = v.w;
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < Cls<CT>::CAP; ++c)
      if (c < n) s[c] = __ldg(base + c * S + i);
  }
}

// per-voxel softmax in place, returns the log-sum-exp pieces
template <int CT>
__device__ __forceinline__ void softmax_c(float (&s)[Cls<CT>::CAP], int C, float& mx, float& sum) {
  const int n = Cls<CT>::n(C);
  mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < Cls<CT>::CAP; ++c)
    if (c < n) mx = fmaxf(mx, s[c]);
  sum = 0.f;
#pragma unroll
  for (int c = 0; c < Cls<CT>::CAP; ++c)
    if (c < n) { s[c] = expf(s[c] - mx); sum += s[c]; }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int c = 0; c < Cls<CT>::CAP; ++c)
    if (c < n) s[c] *= inv;
}

template <typename GT, int CT>
__global__ void __launch_bounds__(NT) loss_sums_kernel(const float* __restrict__ logits, const void* gt, int N,
                                                       int C, int64_t S, int mode, Layout L, double* sums) {
  const int n = blockIdx.y;
  const int nc = Cls<CT>::n(C);
  const float* base = logits + (int64_t)n * (L.rows ? S * L.CP : (int64_t)C * S);
  double dI = 0.0, dU = 0.0, dX = 0.0;
  float fI = 0.f, fU = 0.f, fX = 0.f;
  int cnt = 0;
  bool bad = false;       // fminf/fmaxf drop NaN operands; torch.min/max (losses.py:80-81) propagate them: so do we
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < S; i += (int64_t)gridDim.x * NT) {
    float s[Cls<CT>::CAP], mx, sum;
    load_logits<CT>(base, C, S, i, L, s);
    const int lab = load_gt<GT>(gt, (int64_t)n * S + i, C);
    float logit_L = 0.f;
    if (mode != 0) {
#pragma unroll
      for (int c = 0; c < Cls<CT>::CAP; ++c)
        if (c < nc && c == lab) logit_L = s[c];
    }
    softmax_c<CT>(s, C, mx, sum);
    bad |= sum != sum;
    if (mode == 0) {
      float pfg = 0.f;
#pragma unroll
      for (int c = 1; c < Cls<CT>::CAP; ++c)
        if (c < nc) pfg += s[c];
      const float g = lab > 0 ? 1.f : 0.f;
      fI += fminf(g, pfg);
      fU += fmaxf(g, pfg);
    } else {
      float rest = 0.f, sL = 0.f;
#pragma unroll
      for (int c = 1; c < Cls<CT>::CAP; ++c)
        if (c < nc) { rest += (c == lab) ? 0.f : s[c]; sL = (c == lab) ? s[c] : sL; }
      if (lab >= 1) { fI += (float)(nc - 1) * sL; fU += (float)(nc - 1); }
      fU += rest;
      // cross entropy = -log softmax_L  (computed from the unnormalised pieces)
      fX += (mx + logf(sum)) - logit_L;
    }
    if (++cnt == 64) { dI += fI; dU += fU; dX += fX; fI = fU = fX = 0.f; cnt = 0; }
  }
  dI += fI; dU += fU; dX += fX;
  if (bad) dI = __longlong_as_double(0x7ff8000000000000LL);
  dI = warp_sum(dI); dU = warp_sum(dU); dX = warp_sum(dX);
  __shared__ double red[3][NT / 32];
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = dI; red[1][threadIdx.x >> 5] = dU; red[2][threadIdx.x >> 5] = dX;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int k = 0; k < NT / 32; ++k) t += red[threadIdx.x][k];
    atomic_add_f64(sums + n * 4 + threadIdx.x, t);
  }
}

// loss scalar + backward coefficients from the per-scene sums.
//   coef[2n] = dL/dI_n, coef[2n+1] = dL/dU_n, coef[2N] = dL/d(sum of xent)
__global__ void loss_finalize_kernel(const double* __restrict__ sums, int N, int C, int64_t S, int mode,
                                     float* loss, float* coef) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float mean_iou = 0.f;
  for (int n = 0; n < N; ++n) {
    const float I = (float)sums[n * 4 + 0], U = (float)sums[n * 4 + 1];
    const float Ud = U == 0.f ? 1.f : U;
    mean_iou += I / Ud;
  }
  mean_iou /= (float)N;
  const float A = 1.f - mean_iou;
  // mode 0: iou_fgbg, 1: (1 + iou_agnostic)(1 + xent), 2: iou_agnostic alone (losses.py:19-61), 3: xent alone (:117-141)
  float X = 0.f, fA = 1.f, fX = 0.f;
  if (mode == 1 || mode == 3) {
    double xs = 0.0;
    for (int n = 0; n < N; ++n) xs += sums[n * 4 + 2];
    X = (float)(xs / ((double)N * (double)S));
  }
  if (mode == 1) {
    loss[0] = (1.f + A) * (1.f + X);
    fA = 1.f + X;          // d loss / dA
    fX = 1.f + A;          // d loss / dX
  } else if (mode == 3) {
    loss[0] = X;
    fA = 0.f;
    fX = 1.f;
  } else {
    loss[0] = A;
  }
  for (int n = 0; n < N; ++n) {
    const float I = (float)sums[n * 4 + 0], U = (float)sums[n * 4 + 1];
    const float Ud = U == 0.f ? 1.f : U;
    coef[2 * n] = -fA / ((float)N * Ud);
    coef[2 * n + 1] = U == 0.f ? 0.f : fA * I / ((float)N * Ud * Ud);
  }
  coef[2 * N] = fX / ((float)N * (float)S);
}

template <int CT>
__device__ __forceinline__ void store_logits(float* __restrict__ base, int C, int64_t S, int64_t i, Layout L,
                                             const float (&d)[Cls<CT>::CAP]) {
  const int n = Cls<CT>::n(C);
  if (L.rows) {
    float4* row = reinterpret_cast<float4*>(base + i * L.CP);
#pragma unroll
    for (int q = 0; q < (Cls<CT>::CAP + 3) / 4; ++q) {
      if (q * 4 < L.CP) {
        float4 v;
        v.x = (q * 4 + 0 < Cls<CT>::CAP && q * 4 + 0 < n) ? d[q * 4 + 0 < Cls<CT>::CAP ? q * 4 + 0 : 0] : 0.f;
        v.y = (q * 4 + 1 < Cls<CT>::CAP && q * 4 + 1 < n) ? d[q * 4 + 1 < Cls<CT>::CAP ? q * 4 + 1 : 0] : 0.f;
        v.z = (q * 4 + 2 < Cls<CT>::CAP && q * 4 + 2 < n) ? d[q * 4 + 2 < Cls<CT>::CAP ? q * 4 + 2 : 0] : 0.f;
        v.w = (q * 4 + 3 < Cls<CT>::CAP && q * 4 + 3 < n) ? d[q * 4 + 3 < Cls<CT>::CAP ? q * 4 + 3 : 0] : 0.f;
        row[q] = v;
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < Cls<CT>::CAP; ++c)
      if (c < n) base[c * S + i] = d[c];
  }
}

template <typename GT, int CT>
__global__ void __launch_bounds__(NT) loss_bwd_kernel(const float* __restrict__ logits, const void* gt, int N,
                                                      int C, int64_t S, int mode, Layout L, Layout Lo,
                                                      const float* __restrict__ coef,
                                                      const float* __restrict__ gscale,
                                                      float* __restrict__ dlogits) {
  const int n = blockIdx.y;
  const int nc = Cls<CT>::n(C);
  const float gs = gscale ? __ldg(gscale) : 1.f;
  const float cI = coef[2 * n] * gs, cU = coef[2 * n + 1] * gs, cX = coef[2 * N] * gs;
  const float* base = logits + (int64_t)n * (L.rows ? S * L.CP : (int64_t)C * S);
  float* dbase = dlogits + (int64_t)n * (Lo.rows ? S * Lo.CP : (int64_t)C * S);
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < S; i += (int64_t)gridDim.x * NT) {
    float s[Cls<CT>::CAP], mx, sum;
    load_logits<CT>(base, C, S, i, L, s);
    softmax_c<CT>(s, C, mx, sum);
    const int lab = load_gt<GT>(gt, (int64_t)n * S + i, C);
    float d[Cls<CT>::CAP];
    if (mode == 0) {
      float pfg = 0.f;
#pragma unroll
      for (int c = 1; c < Cls<CT>::CAP; ++c)
        if (c < nc) pfg += s[c];
      const float G = lab > 0 ? cI : cU;     // dL/dp_fg
#pragma unroll
      for (int k = 0; k < Cls<CT>::CAP; ++k)
        if (k < nc) d[k] = G * s[k] * ((k >= 1 ? 1.f : 0.f) - pfg);
    } else {
      // a_c = dL/ds_c
      float dot = 0.f;
#pragma unroll
      for (int c = 1; c < Cls<CT>::CAP; ++c)
        if (c < nc) dot += ((c == lab) ? cI * (float)(nc - 1) : cU) * s[c];
#pragma unroll
      for (int k = 0; k < Cls<CT>::CAP; ++k) {
        if (k < nc) {
          const float a = k == 0 ? 0.f : ((k == lab) ? cI * (float)(nc - 1) : cU);
          d[k] = s[k] * (a - dot) + cX * (s[k] - (k == lab ? 1.f : 0.f));
        }
      }
    }
    store_logits<CT>(dbase, C, S, i, Lo, d);
  }
}

template <int CT>
__global__ void __launch_bounds__(NT) softmax_kernel(const float* __restrict__ logits, int N, int C, int64_t S,
                                                     Layout L, float* __restrict__ pmf) {
  const int n = blockIdx.y;
  const float* base = logits + (int64_t)n * (L.rows ? S * L.CP : (int64_t)C * S);
  float* obase = pmf + (int64_t)n * C * S;
  const Layout planar{0, 0};
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < S; i += (int64_t)gridDim.x * NT) {
    float s[Cls<CT>::CAP], mx, sum;
    load_logits<CT>(base, C, S, i, L, s);
    softmax_c<CT>(s, C, mx, sum);
    store_logits<CT>(obase, C, S, i, planar, s);
  }
}

template <typename GT, int CT>
__global__ void __launch_bounds__(NT) argmax_confusion_kernel(const float* __restrict__ logits, const void* gt,
                                                              int N, int C, int64_t S, Layout L,
                                                              const int32_t* __restrict__ scene_label, int K,
                                                              unsigned long long* cm) {
  extern __shared__ unsigned int hist[];   // K*K
  for (int k = threadIdx.x; k < K * K; k += NT) hist[k] = 0u;
  __syncthreads();
  const int n = blockIdx.y;
  const int nc = Cls<CT>::n(C);
  // FG_BG evaluation (evaluation_results.py:40-51): labels 0/1 are scaled by the scene's dataset class
  const int mul = scene_label ? __ldg(scene_label + n) : 1;
  const float* base = logits + (int64_t)n * (L.rows ? S * L.CP : (int64_t)C * S);
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < S; i += (int64_t)gridDim.x * NT) {
    float s[Cls<CT>::CAP];
    load_logits<CT>(base, C, S, i, L, s);
    float best = s[0];
    int bi = 0;
#pragma unroll
    for (int c = 1; c < Cls<CT>::CAP; ++c)
      if (c < nc && s[c] > best) { best = s[c]; bi = c; }   // first max wins, like torch.argmax
    const int lab = load_gt<GT>(gt, (int64_t)n * S + i, C);
    atomicAdd(&hist[(lab * mul) * K + bi * mul], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K * K; k += NT)
    if (hist[k]) atomicAdd(cm + k, (unsigned long long)hist[k]);
}

inline dim3 grid_ns(int N, int64_t S) {
  int64_t b = crn_ceil_div(S, NT * 4);
  const int64_t cap = crn_ceil_div(8LL * kNumSMs, N);
  if (b > cap) b = cap;
  return dim3((unsigned)(b < 1 ? 1 : b), (unsigned)N, 1);
}

inline bool layout_ok(int32_t rows_cp, int32_t C) { return rows_cp == 0 || (rows_cp >= C && rows_cp % 4 == 0); }

// dispatch on (label type, class count)
#define CRN_LOSS_DISPATCH(KERNEL, GRID, SH, ST, ...)                                                          \
  do {                                                                                                         \
    if (gt_is_i64) {                                                                                           \
      if (C == 2) KERNEL<int64_t, 2><<<GRID, NT, SH, ST>>>(__VA_ARGS__);                                       \
      else if (C == 15) KERNEL<int64_t, 15><<<GRID, NT, SH, ST>>>(__VA_ARGS__);                                \
      else KERNEL<int64_t, 0><<<GRID, NT, SH, ST>>>(__VA_ARGS__);                                              \
    } else {                                                                                                   \
      if (C == 2) KERNEL<int32_t, 2><<<GRID, NT, SH, ST>>>(__VA_ARGS__);                                       \
      else if (C == 15) KERNEL<int32_t, 15><<<GRID, NT, SH, ST>>>(__VA_ARGS__);                                \
      else KERNEL<int32_t, 0><<<GRID, NT, SH, ST>>>(__VA_ARGS__);                                              \
    }                                                                                                          \
  } while (0)
}  // namespace

extern "C" int crn_loss_sums_l(const float* logits, int32_t rows_cp, const void* gt, int32_t gt_is_i64, int32_t N,
                               int32_t C, int64_t S, int32_t mode, double* sums, void* stream) {
  CRN_REQUIRE(logits && gt && sums && N > 0 && C >= 2 && C <= MAXC && S > 0, "crn_loss_sums: bad args (C<=32)");
  CRN_REQUIRE(layout_ok(rows_cp, C), "crn_loss_sums: row pitch must be a multiple of 4 and >= C");
  cudaStream_t st = crn_stream(stream);
  cudaMemsetAsync(sums, 0, sizeof(double) * 4 * N, st);
  const Layout L{rows_cp > 0, rows_cp};
  CRN_LOSS_DISPATCH(loss_sums_kernel, grid_ns(N, S), 0, st, logits, gt, N, C, S, mode, L, sums);
  CRN_LAUNCH_CHECK("loss_sums");
  return CRN_OK;
}

extern "C" int crn_loss_sums(const float* logits, const void* gt, int32_t gt_is_i64, int32_t N, int32_t C,
                             int64_t S, int32_t mode, double* sums, void* stream) {
  return crn_loss_sums_l(logits, 0, gt, gt_is_i64, N, C, S, mode, sums, stream);
}

extern "C" int crn_loss_finalize(const double* sums, int32_t N, int32_t C, int64_t S, int32_t mode,
                                 float* loss, float* coef, void* stream) {
  CRN_REQUIRE(sums && loss && coef && N > 0, "crn_loss_finalize: bad args");
  loss_finalize_kernel<<<1, 32, 0, crn_stream(stream)>>>(sums, N, C, S, mode, loss, coef);
  CRN_LAUNCH_CHECK("loss_finalize");
  return CRN_OK;
}

extern "C" int crn_loss_bwd_l(const float* logits, int32_t rows_cp, const void* gt, int32_t gt_is_i64, int32_t N,
                              int32_t C, int64_t S, int32_t mode, const float* coef, const float* gscale,
                              float* dlogits, int32_t d_rows_cp, void* stream) {
  CRN_REQUIRE(logits && gt && coef && dlogits && N > 0 && C >= 2 && C <= MAXC && S > 0,
              "crn_loss_bwd: bad args (C<=32)");
  CRN_REQUIRE(layout_ok(rows_cp, C) && layout_ok(d_rows_cp, C), "crn_loss_bwd: row pitch must be a multiple of 4, >= C");
  cudaStream_t st = crn_stream(stream);
  const Layout L{rows_cp > 0, rows_cp}, Lo{d_rows_cp > 0, d_rows_cp};
  CRN_LOSS_DISPATCH(loss_bwd_kernel, grid_ns(N, S), 0, st, logits, gt, N, C, S, mode, L, Lo, coef, gscale, dlogits);
  CRN_LAUNCH_CHECK("loss_bwd");
  return CRN_OK;
}

extern "C" int crn_loss_bwd(const float* logits, const void* gt, int32_t gt_is_i64, int32_t N, int32_t C,
                            int64_t S, int32_t mode, const float* coef, const float* gscale,
                            float* dlogits, void* stream) {
  return crn_loss_bwd_l(logits, 0, gt, gt_is_i64, N, C, S, mode, coef, gscale, dlogits, 0, stream);
}

extern "C" int crn_softmax_l(const float* logits, int32_t rows_cp, int32_t N, int32_t C, int64_t S, float* pmf,
                             void* stream) {
  CRN_REQUIRE(logits && pmf && N > 0 && C >= 1 && C <= MAXC && S > 0, "crn_softmax_planar: bad args");
  CRN_REQUIRE(layout_ok(rows_cp, C), "crn_softmax: row pitch must be a multiple of 4 and >= C");
  const Layout L{rows_cp > 0, rows_cp};
  cudaStream_t st = crn_stream(stream);
  if (C == 2) softmax_kernel<2><<<grid_ns(N, S), NT, 0, st>>>(logits, N, C, S, L, pmf);
  else if (C == 15) softmax_kernel<15><<<grid_ns(N, S), NT, 0, st>>>(logits, N, C, S, L, pmf);
  else softmax_kernel<0><<<grid_ns(N, S), NT, 0, st>>>(logits, N, C, S, L, pmf);
  CRN_LAUNCH_CHECK("softmax_planar");
  return CRN_OK;
}

extern "C" int crn_softmax_planar(const float* logits, int32_t N, int32_t C, int64_t S, float* pmf,
                                  void* stream) {
  return crn_softmax_l(logits, 0, N, C, S, pmf, stream);
}

extern "C" int crn_argmax_confusion_l(const float* logits, int32_t rows_cp, const void* gt, int32_t gt_is_i64,
                                      int32_t N, int32_t C, int64_t S, const int32_t* scene_label, int32_t K,
                                      int64_t* cm, void* stream) {
  CRN_REQUIRE(logits && gt && cm && N > 0 && C >= 1 && C <= MAXC && S > 0 && K >= C && K <= 64,
              "crn_argmax_confusion: bad args");
  CRN_REQUIRE(scene_label || K == C, "crn_argmax_confusion: K != C needs scene labels");
  CRN_REQUIRE(layout_ok(rows_cp, C), "crn_argmax_confusion: row pitch must be a multiple of 4 and >= C");
  cudaStream_t st = crn_stream(stream);
  const size_t sh = sizeof(unsigned int) * K * K;
  const Layout L{rows_cp > 0, rows_cp};
  unsigned long long* cmu = reinterpret_cast<unsigned long long*>(cm);
  CRN_LOSS_DISPATCH(argmax_confusion_kernel, grid_ns(N, S), sh, st, logits, gt, N, C, S, L, scene_label, K, cmu);
  CRN_LAUNCH_CHECK("argmax_confusion");
  return CRN_OK;
}

extern "C" int crn_argmax_confusion_labeled(const float* logits, const void* gt, int32_t gt_is_i64, int32_t N,
                                            int32_t C, int64_t S, const int32_t* scene_label, int32_t K,
                                            int64_t* cm, void* stream) {
  return crn_argmax_confusion_l(logits, 0, gt, gt_is_i64, N, C, S, scene_label, K, cm, stream);
}

extern "C" int crn_argmax_confusion(const float* logits, const void* gt, int32_t gt_is_i64, int32_t N,
                                    int32_t C, int64_t S, int64_t* cm, void* stream) {
  return crn_argmax_confusion_l(logits, 0, gt, gt_is_i64, N, C, S, nullptr, C, cm, stream);
}
