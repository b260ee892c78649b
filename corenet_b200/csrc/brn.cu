// Batch renormalisation kernels (HBM-bound, channels-last rows x C).
// Replaces model/batch_renorm.py:33-62 of the reference, plus the ReLU that
// precedes it in the decoder (reconstruction_decoder.py:51-60) and the
// residual add + ReLU that follows it in the encoder (resnet50.py:72-83,110-115).
//
// "Column-owner" mapping: a thread owns VEC consecutive channels and strides
// over rows, so every per-channel reduction stays in registers; partial sums
// are combined in double precision: shared memory inside a block, distributed shared memory
// inside a thread-block cluster of 8, then ONE atomicAdd per cluster and channel -- same-address
// double atomics serialise in L2 (~50 ns each), and with one per block they were a ~15 us floor
// under every pass (scripts/brn_bench.py).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NT = 256;

struct ColGrid {
  int txc;      // threads across channel groups (power of two <= 8)
  int cluster;  // blocks per cluster along x (8, or 1 for tiny grids)
  dim3 grid;
};

static ColGrid col_grid(int64_t rows, int cgroups) {
  ColGrid g;
  // narrow blocks (<= 8 channel groups = 128 B per row) put the parallelism on gridDim.y, where blocks do
  // not share accumulator addresses; gridDim.x (blocks that do share them) is grouped into clusters of 8
  int txc = 1;
  while (txc < cgroups && txc < 8) txc <<= 1;
  g.txc = txc;
  const int tyc = NT / txc;
  const int gy = (int)crn_ceil_div(cgroups, txc);
  int64_t gx = crn_ceil_div(rows, (int64_t)tyc * 4);      // >= 4 rows per thread
  int64_t target = crn_ceil_div(2LL * kNumSMs, gy);
  if (gx > target) gx = target;
  if (gx < 1) gx = 1;
  g.cluster = 1;
  if (gx >= 8) { gx = gx / 8 * 8; g.cluster = 8; }
  g.grid = dim3((unsigned)gx, (unsigned)gy, 1);
  return g;
}

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  static __device__ __forceinline__ void load(const float* p, float v[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float v[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec<1> {
  static __device__ __forceinline__ void load(const float* p, float v[1]) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void store(float* p, const float v[1]) { p[0] = v[0]; }
};

constexpr int UNR = 4;   // independent row loads in flight per thread in the column-owner passes

// plain (coherent) load: dx may have been written earlier in this kernel's lifetime by other kernels, but never
// by this one before the read -- kept off the read-only path only because dx is not const/restrict-read-only
template <int VEC>
__device__ __forceinline__ void load_plain(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = p[0];
  }
}

// Block-level combine of NV per-thread double partials per owned channel.
template <int VEC, int NV>
__device__ __forceinline__ void block_col_reduce(double (&part)[NV][VEC], int txc, int tx, int ty,
                                                 int c_first, int C, double* const (&dst)[NV]) {
  __shared__ double red[NT * 4];
  __shared__ double blk[NV][8 * 4];          // this block's per-channel sums (txc <= 8), read by cluster rank 0
  const int tyc = NT / txc;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < VEC; ++e) red[(ty * txc + tx) * VEC + e] = part[q][e];
    __syncthreads();
    if (ty == 0) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        double s = 0.0;
        for (int y = 0; y < tyc; ++y) s += red[(y * txc + tx) * VEC + e];
        blk[q][tx * VEC + e] = s;
      }
    }
  }
  cg::cluster_group cluster = cg::this_cluster();
  cluster.sync();                            // every block's blk[][] is written
  if (cluster.block_rank() == 0 && ty == 0) {
    const unsigned nb = cluster.num_blocks();
#pragma unroll
    for (int q = 0; q < NV; ++q) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const int c = c_first + e;
        if (c >= C || !dst[q]) continue;
        double s = 0.0;
        for (unsigned rk = 0; rk < nb; ++rk) s += *cluster.map_shared_rank(&blk[q][tx * VEC + e], rk);
        atomic_add_f64(dst[q] + c, s);
      }
    }
  }
  cluster.sync();                            // keep every block's shared memory alive until rank 0 has read it
}

// launch with a thread-block cluster of `cluster` blocks along x (grid.x is a multiple of it)
template <typename... KArgs, typename... Args>
cudaError_t launch_clustered(void (*kern)(KArgs...), dim3 grid, int cluster, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NT, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------- stats
template <int VEC>
__global__ void __launch_bounds__(NT) brn_stats_kernel(const float* __restrict__ x, int64_t rows,
                                                       int C, int x_cs, int x_co, int relu_in,
                                                       double* acc, int txc) {
  const int tx = threadIdx.x % txc, ty = threadIdx.x / txc, tyc = NT / txc;
  const int c_first = (blockIdx.y * txc + tx) * VEC;
  double part[2][VEC];
  float shift[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) { part[0][e] = 0.0; part[1][e] = 0.0; shift[e] = 0.f; }
  if (c_first < C) {
    Vec<VEC>::load(x + x_co + c_first, shift);   // row 0 as the shift
    if (relu_in) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) shift[e] = fmaxf(shift[e], 0.f);
    }
    float s1[VEC], s2[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { s1[e] = 0.f; s2[e] = 0.f; }
    int cnt = 0;
    const int64_t stride = (int64_t)gridDim.x * tyc;
    int64_t r = (int64_t)blockIdx.x * tyc + ty;
    auto accum = [&](const float (&v)[VEC]) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        float u = relu_in ? fmaxf(v[e], 0.f) : v[e];
        u -= shift[e];
        s1[e] += u;
        s2[e] = fmaf(u, u, s2[e]);
      }
    };
    // UNR independent row loads in flight per thread: these passes are latency-bound, not bandwidth-bound
    for (; r + (UNR - 1) * stride < rows; r += UNR * stride) {
      float v[UNR][VEC];
#pragma unroll
      for (int u = 0; u < UNR; ++u) Vec<VEC>::load(x + (r + u * stride) * x_cs + x_co + c_first, v[u]);
#pragma unroll
      for (int u = 0; u < UNR; ++u) accum(v[u]);
      if (++cnt == 256 / UNR) {   // flush fp32 partials into double to bound the error
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          part[0][e] += s1[e]; part[1][e] += s2[e]; s1[e] = 0.f; s2[e] = 0.f;
        }
        cnt = 0;
      }
    }
    for (; r < rows; r += stride) {
      float v[VEC];
      Vec<VEC>::load(x + r * x_cs + x_co + c_first, v);
      accum(v);
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) { part[0][e] += s1[e]; part[1][e] += s2[e]; }
    if (blockIdx.x == 0 && ty == 0) {
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        if (c_first + e < C) acc[2 * C + c_first + e] = (double)shift[e];
    }
  }
  double* const dst[2] = {acc, acc + C};
  block_col_reduce<VEC, 2>(part, txc, tx, ty, c_first, C, dst);
}

// ---------------------------------------------------------------- finalize
__global__ void brn_finalize_kernel(const double* __restrict__ acc, int64_t rows, int C,
                                    const float* __restrict__ weight, const float* __restrict__ bias,
                                    float* running_mean, float* running_var, int64_t* nbt, float eps,
                                    float momentum, int training, float* coef) {
  // ONE block: every thread reads the step counter, the block synchronises, then thread 0 increments it
  // (num_batches_tracked += 1, batch_renorm.py:58) -- no separate launch
  const long long nt_now = training ? *nbt : 0;
  __syncthreads();
  if (training && threadIdx.x == 0) *nbt = nt_now + 1;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
  const float w = weight[c], bz = bias[c];
  const float rm = running_mean[c], rv = running_var[c];
  const float rstd = sqrtf(rv + eps);
  float a, b, mean, invstd, r, d;
  if (training) {
    const long long nt = nt_now;
    float dmax = 5.0f * (float)(nt - 5000) / 20000.0f;
    dmax = fminf(fmaxf(dmax, 0.f), 5.f);
    float rmax = 2.0f * (float)(nt - 5000) / 35000.0f;
    rmax = 1.0f + fminf(fmaxf(rmax, 0.f), 2.f);
    const double R = (double)rows;
    const double m1 = acc[c] / R;
    double var_d = acc[C + c] / R - m1 * m1;
    if (var_d < 0.0) var_d = 0.0;
    mean = (float)(acc[2 * C + c] + m1);
    const float var = (float)var_d;
    const float std = sqrtf(var + eps);
    r = fminf(fmaxf(std / rstd, 1.0f / rmax), rmax);
    d = fminf(fmaxf((mean - rm) / rstd, -dmax), dmax);
    invstd = 1.0f / std;
    a = w * r / std;
    b = fmaf(w, d, bz);          // y = a * (x - mean) + b, mean subtracted first like the reference
    // running statistics (batch_renorm.py:54-57; "Bessel" uses the channel count)
    const float unbiased = var * (float)C / (float)(C - 1);
    running_var[c] = rv + momentum * (unbiased - rv);
    running_mean[c] = rm + momentum * (mean - rm);
  } else {
    mean = rm; invstd = 1.0f / rstd; r = 1.f; d = 0.f;
    a = w / rstd;
    b = bz;
  }
  coef[c] = a; coef[C + c] = b; coef[2 * C + c] = mean; coef[3 * C + c] = invstd;
  coef[4 * C + c] = r; coef[5 * C + c] = d;
  }
}

// ---------------------------------------------------------------- apply
template <int VEC>
__global__ void __launch_bounds__(NT) brn_apply_kernel(const float* __restrict__ x, int64_t rows, int C,
                                                       int x_cs, int x_co, const float* __restrict__ coef,
                                                       const float* __restrict__ res, int relu_in,
                                                       int relu_out, float* __restrict__ y, int y_cs,
                                                       int y_co, float* __restrict__ y_pre) {
  const int cg = C / VEC;
  const int64_t total = rows * cg;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int64_t r = i / cg;
    const int c = (int)(i - r * cg) * VEC;
    float v[VEC], a[VEC], b[VEC], mu[VEC], o[VEC];
    Vec<VEC>::load(x + r * x_cs + x_co + c, v);
    Vec<VEC>::load(coef + c, a);
    Vec<VEC>::load(coef + C + c, b);
    Vec<VEC>::load(coef + 2 * C + c, mu);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float u = relu_in ? fmaxf(v[e], 0.f) : v[e];
      o[e] = fmaf(a[e], u - mu[e], b[e]);
    }
    if (res) {
      float rr[VEC];
      Vec<VEC>::load(res + r * y_cs + y_co + c, rr);
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[e] += rr[e];
    }
    if (y_pre) Vec<VEC>::store(y_pre + r * y_cs + y_co + c, o);
    if (relu_out) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[e] = fmaxf(o[e], 0.f);
    }
    Vec<VEC>::store(y + r * y_cs + y_co + c, o);
  }
}

// ---------------------------------------------------------------- backward pass 1
template <int VEC>
__global__ void __launch_bounds__(NT) brn_bwd_reduce_kernel(
    const float* __restrict__ dy, int dy_cs, int dy_co, const float* __restrict__ y_act,
    const float* __restrict__ g_extra, const float* __restrict__ x, int x_cs, int x_co, int64_t rows,
    int C, const float* __restrict__ coef, int relu_in, int relu_out, float* __restrict__ g_out,
    double* acc, int txc) {
  const int tx = threadIdx.x % txc, ty = threadIdx.x / txc, tyc = NT / txc;
  const int c_first = (blockIdx.y * txc + tx) * VEC;
  double part[2][VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) { part[0][e] = 0.0; part[1][e] = 0.0; }
  if (c_first < C) {
    float mean[VEC], invstd[VEC];
    Vec<VEC>::load(coef + 2 * C + c_first, mean);
    Vec<VEC>::load(coef + 3 * C + c_first, invstd);
    float s1[VEC], s2[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { s1[e] = 0.f; s2[e] = 0.f; }
    int cnt = 0;
    const int64_t stride = (int64_t)gridDim.x * tyc;
    int64_t r = (int64_t)blockIdx.x * tyc + ty;
    auto process = [&](int64_t go, float (&g)[VEC], const float (&ya)[VEC], const float (&ge)[VEC],
                       const float (&v)[VEC]) {
      if (relu_out) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) g[e] = ya[e] > 0.f ? g[e] : 0.f;
      }
      if (g_extra) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) g[e] += ge[e];
      }
      if (g_out) Vec<VEC>::store(g_out + go, g);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float u = relu_in ? fmaxf(v[e], 0.f) : v[e];
        const float xh = (u - mean[e]) * invstd[e];
        s1[e] += g[e];
        s2[e] = fmaf(g[e], xh, s2[e]);
      }
    };
    for (; r + (UNR - 1) * stride < rows; r += UNR * stride) {
      float g[UNR][VEC], ya[UNR][VEC], ge[UNR][VEC], v[UNR][VEC];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t go = (r + u * stride) * dy_cs + dy_co + c_first;
        Vec<VEC>::load(dy + go, g[u]);
        if (relu_out) Vec<VEC>::load(y_act + go, ya[u]);
        if (g_extra) Vec<VEC>::load(g_extra + go, ge[u]);
        Vec<VEC>::load(x + (r + u * stride) * x_cs + x_co + c_first, v[u]);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) process((r + u * stride) * dy_cs + dy_co + c_first, g[u], ya[u], ge[u], v[u]);
      if (++cnt == 256 / UNR) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          part[0][e] += s1[e]; part[1][e] += s2[e]; s1[e] = 0.f; s2[e] = 0.f;
        }
        cnt = 0;
      }
    }
    for (; r < rows; r += stride) {
      float g[VEC], ya[VEC], ge[VEC], v[VEC];
      const int64_t go = r * dy_cs + dy_co + c_first;
      Vec<VEC>::load(dy + go, g);
      if (relu_out) Vec<VEC>::load(y_act + go, ya);
      if (g_extra) Vec<VEC>::load(g_extra + go, ge);
      Vec<VEC>::load(x + r * x_cs + x_co + c_first, v);
      process(go, g, ya, ge, v);
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) { part[0][e] += s1[e]; part[1][e] += s2[e]; }
  }
  double* const dst[2] = {acc, acc + C};
  block_col_reduce<VEC, 2>(part, txc, tx, ty, c_first, C, dst);
}

// ---------------------------------------------------------------- backward pass 2
template <int VEC>
__global__ void __launch_bounds__(NT) brn_bwd_dx_kernel(
    const float* __restrict__ g, int g_cs, int g_co, const float* __restrict__ x, int x_cs, int x_co,
    int64_t rows, int C, const float* __restrict__ coef, const double* __restrict__ acc, int relu_in,
    int training, float* __restrict__ dx, int dx_cs, int dx_co, int dx_accumulate,
    float* __restrict__ dweight, float* __restrict__ dbias, double* dxsum, int txc) {
  const int tx = threadIdx.x % txc, ty = threadIdx.x / txc, tyc = NT / txc;
  const int c_first = (blockIdx.y * txc + tx) * VEC;
  double part[1][VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) part[0][e] = 0.0;
  if (c_first < C) {
    float a[VEC], mean[VEC], invstd[VEC], m1[VEC], m2[VEC];
    Vec<VEC>::load(coef + c_first, a);
    Vec<VEC>::load(coef + 2 * C + c_first, mean);
    Vec<VEC>::load(coef + 3 * C + c_first, invstd);
    const double invR = 1.0 / (double)rows;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int c = c_first + e;
      const double S1 = c < C ? acc[c] : 0.0, S2 = c < C ? acc[C + c] : 0.0;
      m1[e] = training ? (float)(S1 * invR) : 0.f;
      m2[e] = training ? (float)(S2 * invR) : 0.f;
      if (blockIdx.x == 0 && ty == 0 && c < C) {
        const float r = coef[4 * C + c], d = coef[5 * C + c];
        if (dweight) dweight[c] = (float)(r * S2 + d * S1);
        if (dbias) dbias[c] = (float)S1;
      }
    }
    float sdx[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) sdx[e] = 0.f;
    int cnt = 0;
    const int64_t stride = (int64_t)gridDim.x * tyc;
    int64_t r = (int64_t)blockIdx.x * tyc + ty;
    auto process = [&](int64_t rr, const float (&gv)[VEC], const float (&v)[VEC], const float (&old)[VEC]) {
      float o[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float u = relu_in ? fmaxf(v[e], 0.f) : v[e];
        const float xh = (u - mean[e]) * invstd[e];
        float t = a[e] * (gv[e] - m1[e] - xh * m2[e]);
        if (relu_in && !(v[e] > 0.f)) t = 0.f;
        o[e] = t;
        sdx[e] += t;
      }
      if (dx_accumulate) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) o[e] += old[e];
      }
      Vec<VEC>::store(dx + rr * dx_cs + dx_co + c_first, o);
    };
    for (; r + (UNR - 1) * stride < rows; r += UNR * stride) {
      float gv[UNR][VEC], v[UNR][VEC], old[UNR][VEC];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t rr = r + u * stride;
        Vec<VEC>::load(g + rr * g_cs + g_co + c_first, gv[u]);
        Vec<VEC>::load(x + rr * x_cs + x_co + c_first, v[u]);
        if (dx_accumulate) load_plain<VEC>(dx + rr * dx_cs + dx_co + c_first, old[u]);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) process(r + u * stride, gv[u], v[u], old[u]);
      if (++cnt == 256 / UNR) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) { part[0][e] += sdx[e]; sdx[e] = 0.f; }
        cnt = 0;
      }
    }
    for (; r < rows; r += stride) {
      float gv[VEC], v[VEC], old[VEC];
      Vec<VEC>::load(g + r * g_cs + g_co + c_first, gv);
      Vec<VEC>::load(x + r * x_cs + x_co + c_first, v);
      if (dx_accumulate) load_plain<VEC>(dx + r * dx_cs + dx_co + c_first, old);
      process(r, gv, v, old);
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) part[0][e] += sdx[e];
  }
  if (dxsum) {   // uniform across the grid
    double* const dst[1] = {dxsum};
    block_col_reduce<VEC, 1>(part, txc, tx, ty, c_first, C, dst);
  }
}

// ---------------------------------------------------------------- column sums
template <int VEC>
__global__ void __launch_bounds__(NT) colsum_kernel(const float* __restrict__ x, int64_t rows, int C,
                                                    int x_cs, int x_co, double* acc, int txc) {
  const int tx = threadIdx.x % txc, ty = threadIdx.x / txc, tyc = NT / txc;
  const int c_first = (blockIdx.y * txc + tx) * VEC;
  double part[1][VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) part[0][e] = 0.0;
  if (c_first < C) {
    float s[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) s[e] = 0.f;
    int cnt = 0;
    for (int64_t r = (int64_t)blockIdx.x * tyc + ty; r < rows; r += (int64_t)gridDim.x * tyc) {
      float v[VEC];
      Vec<VEC>::load(x + r * x_cs + x_co + c_first, v);
#pragma unroll
      for (int e = 0; e < VEC; ++e) s[e] += v[e];
      if (++cnt == 256) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) { part[0][e] += s[e]; s[e] = 0.f; }
        cnt = 0;
      }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) part[0][e] += s[e];
  }
  double* const dst[1] = {acc};
  block_col_reduce<VEC, 1>(part, txc, tx, ty, c_first, C, dst);
}

__global__ void colsum_planar_kernel(const float* __restrict__ x, int N, int C, int64_t S, double* acc) {
  // grid: (chunks, C, N)
  const int c = blockIdx.y, n = blockIdx.z;
  const float* p = x + ((int64_t)n * C + c) * S;
  float s = 0.f;
  double ds = 0.0;
  int cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < S; i += (int64_t)gridDim.x * NT) {
    s += __ldg(p + i);
    if (++cnt == 256) { ds += s; s = 0.f; cnt = 0; }
  }
  ds += s;
  ds = warp_sum(ds);
  __shared__ double red[NT / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ds;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < NT / 32; ++i) t += red[i];
    atomic_add_f64(acc + c, t);
  }
}

__global__ void f64_to_f32_kernel(const double* __restrict__ a, float* __restrict__ o, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = (float)a[i];
}

inline bool vec4_ok(int C, int cs, int co) { return C % 4 == 0 && cs % 4 == 0 && co % 4 == 0; }

}  // namespace

extern "C" int crn_brn_stats(const float* x, int64_t rows, int32_t C, int32_t x_cs, int32_t x_co,
                             int32_t relu_in, double* acc, void* stream) {
  CRN_REQUIRE(x && acc && rows > 0 && C > 0, "crn_brn_stats: bad args");
  cudaStream_t st = crn_stream(stream);
  if (vec4_ok(C, x_cs, x_co)) {
    ColGrid g = col_grid(rows, C / 4);
    launch_clustered(brn_stats_kernel<4>, g.grid, g.cluster, st, x, rows, C, x_cs, x_co, relu_in, acc, g.txc);
  } else {
    ColGrid g = col_grid(rows, C);
    launch_clustered(brn_stats_kernel<1>, g.grid, g.cluster, st, x, rows, C, x_cs, x_co, relu_in, acc, g.txc);
  }
  CRN_LAUNCH_CHECK("brn_stats");
  return CRN_OK;
}

extern "C" int crn_brn_finalize(const double* acc, int64_t rows, int32_t C, const float* weight,
                                const float* bias, float* running_mean, float* running_var,
                                int64_t* num_batches_tracked, float eps, float momentum,
                                int32_t training, float* coef, void* stream) {
  CRN_REQUIRE(weight && bias && running_mean && running_var && coef && C > 0, "crn_brn_finalize: bad args");
  CRN_REQUIRE(!training || (acc && num_batches_tracked && rows > 0), "crn_brn_finalize: bad training args");
  cudaStream_t st = crn_stream(stream);
  const int threads = C >= 1024 ? 1024 : ((C + 31) / 32 * 32);
  brn_finalize_kernel<<<1, threads, 0, st>>>(acc, rows, C, weight, bias, running_mean, running_var,
                                             num_batches_tracked, eps, momentum, training, coef);
  CRN_LAUNCH_CHECK("brn_finalize");
  return CRN_OK;
}

extern "C" int crn_brn_apply(const float* x, int64_t rows, int32_t C, int32_t x_cs, int32_t x_co,
                             const float* coef, const float* res, int32_t relu_in, int32_t relu_out,
                             float* y, int32_t y_cs, int32_t y_co, float* y_pre, void* stream) {
  CRN_REQUIRE(x && coef && y && rows > 0 && C > 0, "crn_brn_apply: bad args");
  cudaStream_t st = crn_stream(stream);
  const bool v4 = vec4_ok(C, x_cs, x_co) && y_cs % 4 == 0 && y_co % 4 == 0;
  const int64_t total = rows * (v4 ? C / 4 : C);
  int64_t blocks = crn_ceil_div(total, NT);
  if (blocks > 16LL * kNumSMs) blocks = 16LL * kNumSMs;
  if (v4)
    brn_apply_kernel<4><<<(unsigned)blocks, NT, 0, st>>>(x, rows, C, x_cs, x_co, coef, res, relu_in,
                                                         relu_out, y, y_cs, y_co, y_pre);
  else
    brn_apply_kernel<1><<<(unsigned)blocks, NT, 0, st>>>(x, rows, C, x_cs, x_co, coef, res, relu_in,
                                                         relu_out, y, y_cs, y_co, y_pre);
  CRN_LAUNCH_CHECK("brn_apply");
  return CRN_OK;
}

extern "C" int crn_brn_bwd_reduce(const float* dy, int32_t dy_cs, int32_t dy_co, const float* y_act,
                                  const float* g_extra, const float* x, int32_t x_cs, int32_t x_co,
                                  int64_t rows, int32_t C, const float* coef, int32_t relu_in,
                                  int32_t relu_out, float* g_out, double* acc, void* stream) {
  CRN_REQUIRE(dy && x && coef && acc && rows > 0 && C > 0, "crn_brn_bwd_reduce: bad args");
  CRN_REQUIRE(!relu_out || y_act, "crn_brn_bwd_reduce: relu_out needs y_act");
  cudaStream_t st = crn_stream(stream);
  if (vec4_ok(C, x_cs, x_co) && dy_cs % 4 == 0 && dy_co % 4 == 0) {
    ColGrid g = col_grid(rows, C / 4);
    launch_clustered(brn_bwd_reduce_kernel<4>, g.grid, g.cluster, st, dy, dy_cs, dy_co, y_act, g_extra, x, x_cs, x_co, rows,
                                                    C, coef, relu_in, relu_out, g_out, acc, g.txc);
  } else {
    ColGrid g = col_grid(rows, C);
    launch_clustered(brn_bwd_reduce_kernel<1>, g.grid, g.cluster, st, dy, dy_cs, dy_co, y_act, g_extra, x, x_cs, x_co, rows,
                                                    C, coef, relu_in, relu_out, g_out, acc, g.txc);
  }
  CRN_LAUNCH_CHECK("brn_bwd_reduce");
  return CRN_OK;
}

extern "C" int crn_brn_bwd_dx(const float* g, int32_t g_cs, int32_t g_co, const float* x, int32_t x_cs,
                              int32_t x_co, int64_t rows, int32_t C, const float* coef,
                              const double* acc, const float* weight, int32_t relu_in,
                              int32_t training, float* dx, int32_t dx_cs, int32_t dx_co,
                              int32_t dx_accumulate, float* dweight, float* dbias, double* dxsum,
                              void* stream) {
  (void)weight;
  CRN_REQUIRE(g && x && coef && acc && dx && rows > 0 && C > 0, "crn_brn_bwd_dx: bad args");
  cudaStream_t st = crn_stream(stream);
  if (vec4_ok(C, x_cs, x_co) && g_cs % 4 == 0 && g_co % 4 == 0 && dx_cs % 4 == 0 && dx_co % 4 == 0) {
    ColGrid cg = col_grid(rows, C / 4);
    launch_clustered(brn_bwd_dx_kernel<4>, cg.grid, cg.cluster, st, g, g_cs, g_co, x, x_cs, x_co, rows, C, coef, acc, relu_in,
                                                 training, dx, dx_cs, dx_co, dx_accumulate, dweight, dbias,
                                                 dxsum, cg.txc);
  } else {
    ColGrid cg = col_grid(rows, C);
    launch_clustered(brn_bwd_dx_kernel<1>, cg.grid, cg.cluster, st, g, g_cs, g_co, x, x_cs, x_co, rows, C, coef, acc, relu_in,
                                                 training, dx, dx_cs, dx_co, dx_accumulate, dweight, dbias,
                                                 dxsum, cg.txc);
  }
  CRN_LAUNCH_CHECK("brn_bwd_dx");
  return CRN_OK;
}

extern "C" int crn_colsum(const float* x, int64_t rows, int32_t C, int32_t x_cs, int32_t x_co,
                          float* out, double* scratch, void* stream) {
  CRN_REQUIRE(x && out && scratch && rows > 0 && C > 0, "crn_colsum: bad args");
  cudaStream_t st = crn_stream(stream);
  cudaMemsetAsync(scratch, 0, sizeof(double) * C, st);
  if (vec4_ok(C, x_cs, x_co)) {
    ColGrid g = col_grid(rows, C / 4);
    launch_clustered(colsum_kernel<4>, g.grid, g.cluster, st, x, rows, C, x_cs, x_co, scratch, g.txc);
  } else {
    ColGrid g = col_grid(rows, C);
    launch_clustered(colsum_kernel<1>, g.grid, g.cluster, st, x, rows, C, x_cs, x_co, scratch, g.txc);
  }
  f64_to_f32_kernel<<<(C + 127) / 128, 128, 0, st>>>(scratch, out, C);
  crn_count_launches(1);
  CRN_LAUNCH_CHECK("colsum");
  return CRN_OK;
}

extern "C" int crn_colsum_planar(const float* x, int32_t N, int32_t C, int64_t S, float* out,
                                 double* scratch, void* stream) {
  CRN_REQUIRE(x && out && scratch && N > 0 && C > 0 && S > 0, "crn_colsum_planar: bad args");
  cudaStream_t st = crn_stream(stream);
  cudaMemsetAsync(scratch, 0, sizeof(double) * C, st);
  int64_t chunks = crn_ceil_div(S, NT * 16);
  if (chunks > 64) chunks = 64;
  colsum_planar_kernel<<<dim3((unsigned)chunks, C, N), NT, 0, st>>>(x, N, C, S, scratch);
  f64_to_f32_kernel<<<(C + 127) / 128, 128, 0, st>>>(scratch, out, C);
  crn_count_launches(1);
  CRN_LAUNCH_CHECK("colsum_planar");
  return CRN_OK;
}
