// BatchRenorm for SMALL activations in ONE launch per direction (model/batch_renorm.py:33-62 and its backward).
//
// The three-kernel path of brn.cu (stats -> finalize -> apply; reduce -> dx) is right for the big decoder tensors,
// where every pass runs near the HBM rate.  The 53 encoder instances (and the coarse decoder stages) normalise maps of
// 0.1 - 4 M elements: there each launch is a ~5-17 us latency floor (cluster launch, fp64 atomics, dependent launches)
// and the step pays ~320 of them.  Here a thread-block CLUSTER owns a group of 8 channels (32 B of every row)
// completely: its blocks split the rows, reduce through shared memory + distributed shared memory in a FIXED order
// (no atomics: the statistics are bit-reproducible), every block derives the coefficients, and the same blocks apply
// them to the rows they just read (second read served by L2).  Backward likewise: reduce, then dx, in one launch.
//
// num_batches_tracked: blocks of one launch may start after others have finished, so the counter cannot be read and
// incremented inside this kernel; crn_brn_nbt_snapshot copies the counters of all fused instances of a forward pass
// to a snapshot array and increments the originals in ONE launch, the fused kernels read the snapshot.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NT = 256;
constexpr int CG = 8;      // channels per cluster: 2 float4 per row
constexpr int RPS = NT / 2;  // rows per sweep of a block
constexpr int UNR = 4;      // forward: row loads in flight per thread
constexpr int UNB = 2;      // backward: 3-4 streams per row already; more rows in flight cost occupancy (184 registers)

struct BrnFwd {
  const float* x; long long rows; int C, x_cs, x_co, relu_in;
  const float* weight; const float* bias; float* running_mean; float* running_var; const long long* nt;
  float eps, momentum; int training;
  const float* res; int relu_out; float* y; int y_cs, y_co; float* y_pre; float* coef;
};

struct BrnBwd {
  const float* dy; int dy_cs, dy_co; const float* y_act; const float* g_extra; const float* x; int x_cs, x_co;
  long long rows; int C; const float* coef; int relu_in, relu_out, training;
  float* g_out; float* dx; int dx_cs, dx_co, dx_accumulate; float* dweight; float* dbias; double* dxsum;
};

__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4_plain(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}

// Sums NV x 4 per-thread doubles over the threads with the same tx (column quad) of the block, then over the blocks
// of the cluster in rank order.  Every block ends up with the same totals in tot[q][tx * 4 + e].
template <int NV>
__device__ __forceinline__ void cluster_col_sum(double (&part)[NV][4], int tx, double (*blk)[CG], double (*tot)[CG],
                                                cg::cluster_group& cluster) {
  __shared__ double red[NT / 32][2][NV][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < NV; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      double v = part[q][e];
#pragma unroll
      for (int o = 2; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);   // lanes of equal parity = equal tx
      if (lane < 2) red[warp][lane][q][e] = v;
    }
  __syncthreads();
  if (threadIdx.x < NV * CG) {
    const int q = threadIdx.x / CG, c = threadIdx.x % CG;
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) s += red[w][c >> 2][q][c & 3];
    blk[q][c] = s;
  }
  cluster.sync();                                // every block's blk[][] is written
  if (threadIdx.x < NV * CG) {
    const int q = threadIdx.x / CG, c = threadIdx.x % CG;
    double s = 0.0;
    const unsigned nb = cluster.num_blocks();
    for (unsigned rk = 0; rk < nb; ++rk) s += *cluster.map_shared_rank(&blk[q][c], rk);
    tot[q][c] = s;
  }
  __syncthreads();
  (void)tx;
}

__global__ void __launch_bounds__(NT) brn_fused_fwd_kernel(const BrnFwd p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int tx = threadIdx.x & 1, ty = threadIdx.x >> 1;
  const int c0 = blockIdx.y * CG + tx * 4;
  const long long per = (p.rows + CS - 1) / CS;
  const long long r0 = rank * per, r1 = (r0 + per < p.rows) ? r0 + per : p.rows;
  __shared__ double blk[2][CG], tot[2][CG];
  __shared__ float cf[3][CG];                    // a, b, mean of the cluster's channels
  if (p.training) {
    float shift[4];
    ld4(p.x + p.x_co + c0, shift);               // row 0 as the shift (conditioning of the variance)
    if (p.relu_in) {
#pragma unroll
      for (int e = 0; e < 4; ++e) shift[e] = fmaxf(shift[e], 0.f);
    }
    double part[2][4];
    float s1[4], s2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { part[0][e] = part[1][e] = 0.0; s1[e] = s2[e] = 0.f; }
    int cnt = 0;
    auto accum = [&](const float (&v)[4]) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float u = p.relu_in ? fmaxf(v[e], 0.f) : v[e];
        u -= shift[e];
        s1[e] += u;
        s2[e] = fmaf(u, u, s2[e]);
      }
    };
    // UNR predicated row loads in flight per thread (a block owns only 128..2048 rows: latency, not bandwidth)
    for (long long r = r0 + ty; r < r1; r += UNR * RPS) {
      float v[UNR][4];
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        if (r + u * RPS < r1) ld4(p.x + (r + u * RPS) * p.x_cs + p.x_co + c0, v[u]);
#pragma unroll
      for (int u = 0; u < UNR; ++u)
        if (r + u * RPS < r1) accum(v[u]);
      if (++cnt == 64) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { part[0][e] += s1[e]; part[1][e] += s2[e]; s1[e] = s2[e] = 0.f; }
        cnt = 0;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { part[0][e] += s1[e]; part[1][e] += s2[e]; }
    cluster_col_sum<2>(part, tx, blk, tot, cluster);
  }
  if (threadIdx.x < CG) {
    // the arithmetic of brn_finalize_kernel (brn.cu), per channel
    const int c = blockIdx.y * CG + threadIdx.x;
    const int C = p.C;
    const float w = p.weight[c], bz = p.bias[c];
    const float rm = p.running_mean[c], rv = p.running_var[c];
    const float rstd = sqrtf(rv + p.eps);
    float a, b, mean, invstd, r, d;
    if (p.training) {
      const long long nt = *p.nt;
      float dmax = 5.0f * (float)(nt - 5000) / 20000.0f;
      dmax = fminf(fmaxf(dmax, 0.f), 5.f);
      float rmax = 2.0f * (float)(nt - 5000) / 35000.0f;
      rmax = 1.0f + fminf(fmaxf(rmax, 0.f), 2.f);
      const double R = (double)p.rows;
      const double m1 = tot[0][threadIdx.x] / R;
      double var_d = tot[1][threadIdx.x] / R - m1 * m1;
      if (var_d < 0.0) var_d = 0.0;
      float sh = __ldg(p.x + p.x_co + c);
      if (p.relu_in) sh = fmaxf(sh, 0.f);
      mean = (float)((double)sh + m1);
      const float var = (float)var_d;
      const float std = sqrtf(var + p.eps);
      r = fminf(fmaxf(std / rstd, 1.0f / rmax), rmax);
      d = fminf(fmaxf((mean - rm) / rstd, -dmax), dmax);
      invstd = 1.0f / std;
      a = w * r / std;
      b = fmaf(w, d, bz);
      if (rank == 0) {                           // running statistics (batch_renorm.py:54-57, channel-count "Bessel")
        const float unbiased = var * (float)C / (float)(C - 1);
        p.running_var[c] = rv + p.momentum * (unbiased - rv);
        p.running_mean[c] = rm + p.momentum * (mean - rm);
      }
    } else {
      mean = rm; invstd = 1.0f / rstd; r = 1.f; d = 0.f;
      a = w / rstd;
      b = bz;
    }
    cf[0][threadIdx.x] = a; cf[1][threadIdx.x] = b; cf[2][threadIdx.x] = mean;
    if (rank == 0) {
      p.coef[c] = a; p.coef[C + c] = b; p.coef[2 * C + c] = mean; p.coef[3 * C + c] = invstd;
      p.coef[4 * C + c] = r; p.coef[5 * C + c] = d;
    }
  }
  __syncthreads();
  float a[4], b[4], mu[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) { a[e] = cf[0][tx * 4 + e]; b[e] = cf[1][tx * 4 + e]; mu[e] = cf[2][tx * 4 + e]; }
  auto apply = [&](long long rr, const float (&v)[4], const float (&rs)[4]) {
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float u = p.relu_in ? fmaxf(v[e], 0.f) : v[e];
      o[e] = fmaf(a[e], u - mu[e], b[e]);
    }
    if (p.res) {
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] += rs[e];
    }
    if (p.y_pre) st4(p.y_pre + rr * p.y_cs + p.y_co + c0, o);
    if (p.relu_out) {
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
    }
    st4(p.y + rr * p.y_cs + p.y_co + c0, o);
  };
  for (long long r = r0 + ty; r < r1; r += UNR * RPS) {
    float v[UNR][4], rs[UNR][4];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (r + u * RPS < r1) {
        ld4(p.x + (r + u * RPS) * p.x_cs + p.x_co + c0, v[u]);
        if (p.res) ld4(p.res + (r + u * RPS) * p.y_cs + p.y_co + c0, rs[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (r + u * RPS < r1) apply(r + u * RPS, v[u], rs[u]);
  }
  if (p.training) cluster.sync();                // nobody leaves while its blk[][] can still be read
}

__global__ void __launch_bounds__(NT) brn_fused_bwd_kernel(const BrnBwd p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int tx = threadIdx.x & 1, ty = threadIdx.x >> 1;
  const int c0 = blockIdx.y * CG + tx * 4;
  const int C = p.C;
  const long long per = (p.rows + CS - 1) / CS;
  const long long r0 = rank * per, r1 = (r0 + per < p.rows) ? r0 + per : p.rows;
  __shared__ double blk[2][CG], tot[2][CG], blk3[1][CG], tot3[1][CG];
  float aa[4], mean[4], invstd[4];
  ld4(p.coef + c0, aa);
  ld4(p.coef + 2 * C + c0, mean);
  ld4(p.coef + 3 * C + c0, invstd);
  // ---- pass 1: g = dy * [y > 0] + g_extra (stored if asked), S1 = sum g, S2 = sum g * xhat
  auto load_g = [&](long long rr, float (&g)[4]) {
    const long long go = rr * p.dy_cs + p.dy_co + c0;
    ld4_plain(p.dy + go, g);
    if (p.relu_out) {
      float ya[4];
      ld4(p.y_act + go, ya);
#pragma unroll
      for (int e = 0; e < 4; ++e) g[e] = ya[e] > 0.f ? g[e] : 0.f;
    }
    if (p.g_extra) {
      float ge[4];
      ld4(p.g_extra + go, ge);
#pragma unroll
      for (int e = 0; e < 4; ++e) g[e] += ge[e];
    }
  };
  {
    double part[2][4];
    float s1[4], s2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { part[0][e] = part[1][e] = 0.0; s1[e] = s2[e] = 0.f; }
    int cnt = 0;
    for (long long r = r0 + ty; r < r1; r += UNB * RPS) {
      float g[UNB][4], v[UNB][4];
#pragma unroll
      for (int u = 0; u < UNB; ++u) {
        if (r + u * RPS < r1) {
          load_g(r + u * RPS, g[u]);
          ld4(p.x + (r + u * RPS) * p.x_cs + p.x_co + c0, v[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < UNB; ++u) {
        if (r + u * RPS < r1) {
          if (p.g_out) st4(p.g_out + (r + u * RPS) * p.dy_cs + p.dy_co + c0, g[u]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float uu = p.relu_in ? fmaxf(v[u][e], 0.f) : v[u][e];
            const float xh = (uu - mean[e]) * invstd[e];
            s1[e] += g[u][e];
            s2[e] = fmaf(g[u][e], xh, s2[e]);
          }
        }
      }
      if (++cnt == 64) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { part[0][e] += s1[e]; part[1][e] += s2[e]; s1[e] = s2[e] = 0.f; }
        cnt = 0;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { part[0][e] += s1[e]; part[1][e] += s2[e]; }
    cluster_col_sum<2>(part, tx, blk, tot, cluster);
  }
  if (rank == 0 && threadIdx.x < CG) {
    const int c = blockIdx.y * CG + threadIdx.x;
    const double S1 = tot[0][threadIdx.x], S2 = tot[1][threadIdx.x];
    const float r = p.coef[4 * C + c], d = p.coef[5 * C + c];
    if (p.dweight) p.dweight[c] = (float)(r * S2 + d * S1);
    if (p.dbias) p.dbias[c] = (float)S1;
  }
  // ---- pass 2: dx
  float m1[4], m2[4];
  const double invR = 1.0 / (double)p.rows;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    m1[e] = p.training ? (float)(tot[0][tx * 4 + e] * invR) : 0.f;
    m2[e] = p.training ? (float)(tot[1][tx * 4 + e] * invR) : 0.f;
  }
  double part3[1][4];
  float sdx[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) { part3[0][e] = 0.0; sdx[e] = 0.f; }
  int cnt = 0;
  for (long long r = r0 + ty; r < r1; r += UNB * RPS) {
    float g[UNB][4], v[UNB][4], old[UNB][4];
#pragma unroll
    for (int u = 0; u < UNB; ++u) {
      const long long rr = r + u * RPS;
      if (rr < r1) {
        if (p.g_out) ld4_plain(p.g_out + rr * p.dy_cs + p.dy_co + c0, g[u]);   // this thread's own store of pass 1
        else load_g(rr, g[u]);
        ld4(p.x + rr * p.x_cs + p.x_co + c0, v[u]);
        if (p.dx_accumulate) ld4_plain(p.dx + rr * p.dx_cs + p.dx_co + c0, old[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UNB; ++u) {
      const long long rr = r + u * RPS;
      if (rr < r1) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float uu = p.relu_in ? fmaxf(v[u][e], 0.f) : v[u][e];
          const float xh = (uu - mean[e]) * invstd[e];
          float tt = aa[e] * (g[u][e] - m1[e] - xh * m2[e]);
          if (p.relu_in && !(v[u][e] > 0.f)) tt = 0.f;
          o[e] = tt;
          sdx[e] += tt;
        }
        if (p.dx_accumulate) {
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] += old[u][e];
        }
        st4(p.dx + rr * p.dx_cs + p.dx_co + c0, o);
      }
    }
    if (++cnt == 64) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { part3[0][e] += sdx[e]; sdx[e] = 0.f; }
      cnt = 0;
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) part3[0][e] += sdx[e];
  if (p.dxsum) {                                  // uniform across the grid
    cluster_col_sum<1>(part3, tx, blk3, tot3, cluster);
    if (rank == 0 && threadIdx.x < CG) p.dxsum[blockIdx.y * CG + threadIdx.x] = tot3[0][threadIdx.x];
  }
  cluster.sync();                                 // keep shared memory alive until every rank has read it
}

__global__ void brn_nbt_snapshot_kernel(long long* const* ptrs, int n, long long* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const long long v = *ptrs[i];
    out[i] = v;
    *ptrs[i] = v + 1;
  }
}

template <typename P>
cudaError_t launch_fused(void (*kern)(const P), const P& p, long long rows, int C, cudaStream_t st) {
  // blocks per cluster (row split): as many as keep every block at >= 128 rows AND the whole grid within one wave of
  // two blocks per SM -- these blocks are latency-bound (a handful of dependent round trips each), so a second wave
  // of tiny blocks doubles the kernel time (wide layers: C/8 clusters already fill the machine)
  const int groups = C / CG;
  int cs = 1;
  while (cs < 8 && rows / (2 * cs) >= RPS && 2 * cs * groups <= 2 * kNumSMs) cs <<= 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)cs, (unsigned)(C / CG), 1);
  cfg.blockDim = dim3(NT, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

inline bool aligned4(int a, int b) { return a % 4 == 0 && b % 4 == 0; }
}  // namespace

// 1 when the fused single-launch kernels take this instance: channels in groups of 8, small enough that the second
// read of a pass is served by L2 and that C/8 clusters of <= 8 blocks fill the machine.
extern "C" int crn_brn_fused_supported(int64_t rows, int32_t C) {
  // at most 16384 rows: a cluster splits the rows over <= 8 blocks, and beyond ~2048 rows per block the three-kernel
  // path of brn.cu (row-parallel grids) is faster (measured: 65536 x 64 took 32 / 93 us fused vs ~20 / 25 us)
  return rows >= 1 && rows <= 16384 && C >= CG && C % CG == 0 && rows * (int64_t)C <= (int64_t)5 * 1024 * 1024;
}

extern "C" int crn_brn_nbt_snapshot(int64_t* const* counters, int32_t n, int64_t* snapshot, void* stream) {
  CRN_REQUIRE(counters && snapshot && n > 0, "crn_brn_nbt_snapshot: bad args");
  brn_nbt_snapshot_kernel<<<(n + 127) / 128, 128, 0, crn_stream(stream)>>>(
      reinterpret_cast<long long* const*>(counters), n, reinterpret_cast<long long*>(snapshot));
  CRN_LAUNCH_CHECK("brn_nbt_snapshot");
  return CRN_OK;
}

extern "C" int crn_brn_fwd_fused(const float* x, int64_t rows, int32_t C, int32_t x_cs, int32_t x_co, int32_t relu_in,
                                 const float* weight, const float* bias, float* running_mean, float* running_var,
                                 const int64_t* nt_snapshot, float eps, float momentum, int32_t training,
                                 const float* res, int32_t relu_out, float* y, int32_t y_cs, int32_t y_co,
                                 float* y_pre, float* coef, void* stream) {
  CRN_REQUIRE(x && weight && bias && running_mean && running_var && y && coef, "crn_brn_fwd_fused: null pointer");
  CRN_REQUIRE(crn_brn_fused_supported(rows, C) && aligned4(x_cs, x_co) && aligned4(y_cs, y_co),
              "crn_brn_fwd_fused: unsupported shape (see crn_brn_fused_supported)");
  CRN_REQUIRE(!training || nt_snapshot, "crn_brn_fwd_fused: training needs the counter snapshot");
  BrnFwd p{x, rows, C, x_cs, x_co, relu_in, weight, bias, running_mean, running_var,
           reinterpret_cast<const long long*>(nt_snapshot), eps, momentum, training, res, relu_out, y, y_cs, y_co,
           y_pre, coef};
  if (launch_fused(brn_fused_fwd_kernel, p, rows, C, crn_stream(stream)) != cudaSuccess) {
    crn_set_error("crn_brn_fwd_fused: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return CRN_ERR_LAUNCH;
  }
  CRN_LAUNCH_CHECK("brn_fwd_fused");
  return CRN_OK;
}

extern "C" int crn_brn_bwd_fused(const float* dy, int32_t dy_cs, int32_t dy_co, const float* y_act, const float* g_extra,
                                 const float* x, int32_t x_cs, int32_t x_co, int64_t rows, int32_t C, const float* coef,
                                 int32_t relu_in, int32_t relu_out, int32_t training, float* g_out, float* dx,
                                 int32_t dx_cs, int32_t dx_co, int32_t dx_accumulate, float* dweight, float* dbias,
                                 double* dxsum, void* stream) {
  CRN_REQUIRE(dy && x && coef && dx, "crn_brn_bwd_fused: null pointer");
  CRN_REQUIRE(!relu_out || y_act, "crn_brn_bwd_fused: relu_out needs y_act");
  CRN_REQUIRE(crn_brn_fused_supported(rows, C) && aligned4(x_cs, x_co) && aligned4(dy_cs, dy_co) && aligned4(dx_cs, dx_co),
              "crn_brn_bwd_fused: unsupported shape (see crn_brn_fused_supported)");
  BrnBwd p{dy, dy_cs, dy_co, y_act, g_extra, x, x_cs, x_co, rows, C, coef, relu_in, relu_out, training, g_out, dx,
           dx_cs, dx_co, dx_accumulate, dweight, dbias, dxsum};
  if (launch_fused(brn_fused_bwd_kernel, p, rows, C, crn_stream(stream)) != cudaSuccess) {
    crn_set_error("crn_brn_bwd_fused: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return CRN_ERR_LAUNCH;
  }
  CRN_LAUNCH_CHECK("brn_bwd_fused");
  return CRN_OK;
}
