// Fused losses / eval reductions over planar logits [N, C, S] (S = D*H*W).
// Replaces model/losses.py:19-61 (iou_agnostic), :64-114 (iou_fgbg), :117-141
// (xent), :144-160 (xent_times_iou_agnostic) and, for eval,
// evaluation_results.py:40-51 + voxel_metrics.py:33-58 (argmax -> confusion).
// One pass over the logits forward, one pass backward (HBM-bound).
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

// per-voxel softmax into s[], returns log-sum-exp pieces
__device__ __forceinline__ void softmax_c(const float* __restrict__ p, int C, int64_t S, float* s,
                                          float& mx, float& sum) {
  mx = -INFINITY;
  for (int c = 0; c < C; ++c) { s[c] = __ldg(p + c * S); mx = fmaxf(mx, s[c]); }
  sum = 0.f;
  for (int c = 0; c < C; ++c) { s[c] = expf(s[c] - mx); sum += s[c]; }
  const float inv = 1.0f / sum;
  for (int c = 0; c < C; ++c) s[c] *= inv;
}

template <typename GT>
__global__ void __launch_bounds__(NT) loss_sums_kernel(const float* __restrict__ logits, const void* gt, int N,
                                                       int C, int64_t S, int mode, double* sums) {
  const int n = blockIdx.y;
  const float* base = logits + (int64_t)n * C * S;
  double dI = 0.0, dU = 0.0, dX = 0.0;
  float fI = 0.f, fU = 0.f, fX = 0.f;
  int cnt = 0;
  bool bad = false;       // fminf/fmaxf drop NaN operands; torch.min/max (losses.py:80-81) propagate them: so do we
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < S; i += (int64_t)gridDim.x * NT) {
    float s[MAXC], mx, sum;
    softmax_c(base + i, C, S, s, mx, sum);
    bad |= sum != sum;
    const int L = load_gt<GT>(gt, (int64_t)n * S + i, C);
    if (mode == 0) {
      float pfg = 0.f;
      for (int c = 1; c < C; ++c) pfg += s[c];
      const float g = L > 0 ? 1.f : 0.f;
      fI += fminf(g, pfg);
      fU += fmaxf(g, pfg);
    } else {
      float rest = 0.f;
      for (int c = 1; c < C; ++c) rest += (c == L) ? 0.f : s[c];
      if (L >= 1) { fI += (float)(C - 1) * s[L]; fU += (float)(C - 1); }
      fU += rest;
      // cross entropy = -log softmax_L  (computed from the unnormalised pieces)
      const float logit_L = __ldg(base + i + (int64_t)L * S);
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
  float X = 0.f, fA = 1.f, fX = 0.f;
  if (mode == 1) {
    double xs = 0.0;
    for (int n = 0; n < N; ++n) xs += sums[n * 4 + 2];
    X = (float)(xs / ((double)N * (double)S));
    loss[0] = (1.f + A) * (1.f + X);
    fA = 1.f + X;          // d loss / dA
    fX = 1.f + A;          // d loss / dX
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

template <typename GT>
__global__ void __launch_bounds__(NT) loss_bwd_kernel(const float* __restrict__ logits, const void* gt, int N,
                                                      int C, int64_t S, int mode,
                                                      const float* __restrict__ coef,
                                                      const float* __restrict__ gscale,
                                                      float* __restrict__ dlogits) {
  const int n = blockIdx.y;
  const float gs = gscale ? __ldg(gscale) : 1.f;
  const float cI = coef[2 * n] * gs, cU = coef[2 * n + 1] * gs, cX = coef[2 * N] * gs;
  const float* base = logits + (int64_t)n * C * S;
  float* dbase = dlogits + (int64_t)n * C * S;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < S; i += (int64_t)gridDim.x * NT) {
    float s[MAXC], mx, sum;
    softmax_c(base + i, C, S, s, mx, sum);
    const int L = load_gt<GT>(gt, (int64_t)n * S + i, C);
    if (mode == 0) {
      float pfg = 0.f;
      for (int c = 1; c < C; ++c) pfg += s[c];
      const float G = L > 0 ? cI : cU;     // dL/dp_fg
      for (int k = 0; k < C; ++k) dbase[i + (int64_t)k * S] = G * s[k] * ((k >= 1 ? 1.f : 0.f) - pfg);
    } else {
      // a_c = dL/ds_c
      float dot = 0.f;
      for (int c = 1; c < C; ++c) {
        const float a = (c == L) ? cI * (float)(C - 1) : cU;
        dot += a * s[c];
      }
      for (int k = 0; k < C; ++k) {
        const float a = k == 0 ? 0.f : ((k == L) ? cI * (float)(C - 1) : cU);
        dbase[i + (int64_t)k * S] = s[k] * (a - dot) + cX * (s[k] - (k == L ? 1.f : 0.f));
      }
    }
  }
}

__global__ void __launch_bounds__(NT) softmax_planar_kernel(const float* __restrict__ logits, int N, int C,
                                                            int64_t S, float* __restrict__ pmf) {
  const int n = blockIdx.y;
  const float* base = logits + (int64_t)n * C * S;
  float* obase = pmf + (int64_t)n * C * S;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < S; i += (int64_t)gridDim.x * NT) {
    float s[MAXC], mx, sum;
    softmax_c(base + i, C, S, s, mx, sum);
    for (int c = 0; c < C; ++c) obase[i + (int64_t)c * S] = s[c];
  }
}

template <typename GT>
__global__ void __launch_bounds__(NT) argmax_confusion_kernel(const float* __restrict__ logits, const void* gt,
                                                              int N, int C, int64_t S,
                                                              const int32_t* __restrict__ scene_label, int K,
                                                              unsigned long long* cm) {
  extern __shared__ unsigned int hist[];   // K*K
  for (int k = threadIdx.x; k < K * K; k += NT) hist[k] = 0u;
  __syncthreads();
  const int n = blockIdx.y;
  // FG_BG evaluation (evaluation_results.py:40-51): labels 0/1 are scaled by the scene's dataset class
  const int mul = scene_label ? __ldg(scene_label + n) : 1;
  const float* base = logits + (int64_t)n * C * S;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < S; i += (int64_t)gridDim.x * NT) {
    float best = __ldg(base + i);
    int bi = 0;
    for (int c = 1; c < C; ++c) {
      const float v = __ldg(base + i + (int64_t)c * S);
      if (v > best) { best = v; bi = c; }   // first max wins, like torch.argmax
    }
    const int L = load_gt<GT>(gt, (int64_t)n * S + i, C);
    atomicAdd(&hist[(L * mul) * K + bi * mul], 1u);
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
}  // namespace

extern "C" int crn_loss_sums(const float* logits, const void* gt, int32_t gt_is_i64, int32_t N, int32_t C,
                             int64_t S, int32_t mode, double* sums, void* stream) {
  CRN_REQUIRE(logits && gt && sums && N > 0 && C >= 2 && C <= MAXC && S > 0, "crn_loss_sums: bad args (C<=32)");
  cudaStream_t st = crn_stream(stream);
  cudaMemsetAsync(sums, 0, sizeof(double) * 4 * N, st);
  if (gt_is_i64) loss_sums_kernel<int64_t><<<grid_ns(N, S), NT, 0, st>>>(logits, gt, N, C, S, mode, sums);
  else loss_sums_kernel<int32_t><<<grid_ns(N, S), NT, 0, st>>>(logits, gt, N, C, S, mode, sums);
  CRN_LAUNCH_CHECK("loss_sums");
  return CRN_OK;
}

extern "C" int crn_loss_finalize(const double* sums, int32_t N, int32_t C, int64_t S, int32_t mode,
                                 float* loss, float* coef, void* stream) {
  CRN_REQUIRE(sums && loss && coef && N > 0, "crn_loss_finalize: bad args");
  loss_finalize_kernel<<<1, 32, 0, crn_stream(stream)>>>(sums, N, C, S, mode, loss, coef);
  CRN_LAUNCH_CHECK("loss_finalize");
  return CRN_OK;
}

extern "C" int crn_loss_bwd(const float* logits, const void* gt, int32_t gt_is_i64, int32_t N, int32_t C,
                            int64_t S, int32_t mode, const float* coef, const float* gscale,
                            float* dlogits, void* stream) {
  CRN_REQUIRE(logits && gt && coef && dlogits && N > 0 && C >= 2 && C <= MAXC && S > 0,
              "crn_loss_bwd: bad args (C<=32)");
  cudaStream_t st = crn_stream(stream);
  if (gt_is_i64)
    loss_bwd_kernel<int64_t><<<grid_ns(N, S), NT, 0, st>>>(logits, gt, N, C, S, mode, coef, gscale, dlogits);
  else
    loss_bwd_kernel<int32_t><<<grid_ns(N, S), NT, 0, st>>>(logits, gt, N, C, S, mode, coef, gscale, dlogits);
  CRN_LAUNCH_CHECK("loss_bwd");
  return CRN_OK;
}

extern "C" int crn_softmax_planar(const float* logits, int32_t N, int32_t C, int64_t S, float* pmf,
                                  void* stream) {
  CRN_REQUIRE(logits && pmf && N > 0 && C >= 1 && C <= MAXC && S > 0, "crn_softmax_planar: bad args");
  softmax_planar_kernel<<<grid_ns(N, S), NT, 0, crn_stream(stream)>>>(logits, N, C, S, pmf);
  CRN_LAUNCH_CHECK("softmax_planar");
  return CRN_OK;
}

extern "C" int crn_argmax_confusion_labeled(const float* logits, const void* gt, int32_t gt_is_i64, int32_t N,
                                            int32_t C, int64_t S, const int32_t* scene_label, int32_t K,
                                            int64_t* cm, void* stream) {
  CRN_REQUIRE(logits && gt && cm && N > 0 && C >= 1 && C <= MAXC && S > 0 && K >= C && K <= 64,
              "crn_argmax_confusion: bad args");
  CRN_REQUIRE(scene_label || K == C, "crn_argmax_confusion: K != C needs scene labels");
  cudaStream_t st = crn_stream(stream);
  const size_t sh = sizeof(unsigned int) * K * K;
  if (gt_is_i64)
    argmax_confusion_kernel<int64_t><<<grid_ns(N, S), NT, sh, st>>>(
        logits, gt, N, C, S, scene_label, K, reinterpret_cast<unsigned long long*>(cm));
  else
    argmax_confusion_kernel<int32_t><<<grid_ns(N, S), NT, sh, st>>>(
        logits, gt, N, C, S, scene_label, K, reinterpret_cast<unsigned long long*>(cm));
  CRN_LAUNCH_CHECK("argmax_confusion");
  return CRN_OK;
}

extern "C" int crn_argmax_confusion(const float* logits, const void* gt, int32_t gt_is_i64, int32_t N,
                                    int32_t C, int64_t S, int64_t* cm, void* stream) {
  return crn_argmax_confusion_labeled(logits, gt, gt_is_i64, N, C, S, nullptr, C, cm, stream);
}
