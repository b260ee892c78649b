// Generic fp32 implicit-GEMM convolution engine (CUDA-core FFMA, sm_100a).
//
// One "gather-GEMM" kernel covers conv fwd, transposed-conv fwd, and both
// dgrads, in 2-D and 3-D, for any stride:
//
//   out[r][n] = sum_{tap} sum_{k} in[row(r, tap)][k] * W[tap][k][n]
//
//  * DIRECT mode  (conv fwd, convT dgrad):   in_pos = idx*s - pad + k
//  * CLASS  mode  (convT fwd, conv dgrad):   output positions are split into
//    s^3 parity classes (blockIdx.z); inside a class the op is a stride-1
//    gather with the taps k == (c+pad) mod s:   in_pos = idx + (c+pad-k)/s
//
// and one wgrad kernel (rows are the GEMM K dimension, split-K + atomics).
// fp32 accumulate in fp32 FFMA: this is the precision-safe path (DESIGN.md
// "Precision": the net amplifies operand rounding ~1000x at init, so TF32
// single-pass tensor-core math cannot meet the 1e-3 parity budget).
//
// Replaces the ATen/cuDNN convolution calls at model/resnet50.py:61-70,94-108,
// 122-131 and model/reconstruction_decoder.py:47-95 of the reference.
#include "common.cuh"

namespace {

constexpr int KC = 16;    // K-chunk (channels per smem stage)
constexpr int NT = 256;   // threads per block

struct Axis3 { int v[3]; };

struct GatherParams {
  const float* in;
  const float* w;
  const float* bias;
  float* out;
  int N;
  int iD[3];        // input spatial dims  (z,y,x)
  int oD[3];        // output spatial dims (z,y,x)
  int Kd[3];        // kernel dims
  int s[3];         // per-axis stride
  int pad[3];       // per-axis pad
  int gK, gN;       // GEMM K (per tap) and N extents (logical)
  int wK, wN;       // packed weight dims per tap
  int in_cs, in_co, out_cs, out_co;
  int class_mode;   // 0 DIRECT, 1 CLASS
  int accumulate, planar, bias_n_stride;
  int ksplit;       // >1: the (tap, K-chunk) loop is split over blockIdx.z and partial sums are added atomically
};

struct AxisPlan {
  int l0, lstep, lext;           // output lattice
  int istep;                     // input index step per lattice idx
  int nk, k0, kstep, off0, offstep;
};

__device__ __forceinline__ AxisPlan make_axis(const GatherParams& p, int a, int c) {
  AxisPlan q;
  if (!p.class_mode) {
    q.l0 = 0; q.lstep = 1; q.lext = p.oD[a];
    q.istep = p.s[a];
    q.nk = p.Kd[a]; q.k0 = 0; q.kstep = 1; q.off0 = -p.pad[a]; q.offstep = 1;
  } else {
    const int s = p.s[a];
    q.l0 = c; q.lstep = s; q.lext = (p.oD[a] - c + s - 1) / s;
    if (q.lext < 0) q.lext = 0;
    q.istep = 1;
    q.k0 = (c + p.pad[a]) % s;
    q.kstep = s;
    q.nk = q.k0 < p.Kd[a] ? (p.Kd[a] - q.k0 + s - 1) / s : 0;
    q.off0 = (c + p.pad[a] - q.k0) / s;
    q.offstep = -1;
  }
  return q;
}

template <int TM, int TN, int RM, int RN>
__global__ void __launch_bounds__(NT, (RM * RN <= 32) ? 2 : 1) gather_gemm_kernel(const GatherParams p) {
  constexpr int TX = TN / RN;
  constexpr int TY = TM / RM;
  static_assert(TX * TY == NT, "thread layout");
  static_assert(RM % 4 == 0, "RM");
  constexpr int LDA = TM + 4;
  __shared__ __align__(16) float As[KC][LDA];
  __shared__ __align__(16) float Bs[KC][TN];

  const int tid = threadIdx.x;
  // ---- per-class axis plans
  int cz = 0, cy = 0, cx = 0;
  const int split = blockIdx.z % p.ksplit;
  if (p.class_mode) {
    int c = blockIdx.z / p.ksplit;
    cx = c % p.s[2]; c /= p.s[2];
    cy = c % p.s[1]; c /= p.s[1];
    cz = c;
  }
  const AxisPlan az = make_axis(p, 0, cz), ay = make_axis(p, 1, cy), ax = make_axis(p, 2, cx);
  const long long rows = (long long)p.N * az.lext * ay.lext * ax.lext;
  const long long r0 = (long long)blockIdx.x * TM;
  if (r0 >= rows) return;
  const int c0 = blockIdx.y * TN;

  const int ntaps = az.nk * ay.nk * ax.nk;
  const int nchunks = (p.gK + KC - 1) / KC;
  const int niter_all = ntaps * nchunks;
  const int it_begin = (int)((long long)niter_all * split / p.ksplit);
  const int niter = (int)((long long)niter_all * (split + 1) / p.ksplit);   // loop runs [it_begin, niter)
  const bool atomic_out = p.ksplit > 1;

  // ---- A-load slots: this thread loads float4 (row a_row0 + s*64, chans a_kq*4..+3)
  constexpr int LA = (TM * KC / 4) / NT;
  static_assert(LA >= 1, "LA");
  const int a_kq = tid & 3;
  const int a_row0 = tid >> 2;
  int sz[LA], sy[LA], sx[LA];
  long long sbase[LA];
#pragma unroll
  for (int s = 0; s < LA; ++s) {
    long long r = r0 + a_row0 + s * (NT / 4);
    if (r < rows) {
      int ix = (int)(r % ax.lext); long long q = r / ax.lext;
      int iy = (int)(q % ay.lext); q /= ay.lext;
      int iz = (int)(q % az.lext); int n = (int)(q / az.lext);
      sz[s] = iz * az.istep; sy[s] = iy * ay.istep; sx[s] = ix * ax.istep;
      sbase[s] = (long long)n * p.iD[0];
    } else {
      sz[s] = -(1 << 28); sy[s] = 0; sx[s] = 0; sbase[s] = 0;   // always out of bounds
    }
  }
  // ---- B-load slots
  constexpr int B_F4 = KC * TN / 4;
  constexpr int LB = (B_F4 + NT - 1) / NT;

  float4 ra[LA];
  float4 rb[LB];

  auto load_global = [&](int it) {
    const int tap = it / nchunks;
    const int chunk = it - tap * nchunks;
    int jx = tap % ax.nk; int q = tap / ax.nk;
    int jy = q % ay.nk; int jz = q / ay.nk;
    const int dz = az.off0 + az.offstep * jz, dy = ay.off0 + ay.offstep * jy,
              dx = ax.off0 + ax.offstep * jx;
    const int kz = az.k0 + az.kstep * jz, ky = ay.k0 + ay.kstep * jy, kx = ax.k0 + ax.kstep * jx;
    const int wtap = (kz * p.Kd[1] + ky) * p.Kd[2] + kx;
    const int kc = chunk * KC + a_kq * 4;
#pragma unroll
    for (int s = 0; s < LA; ++s) {
      const int pz = sz[s] + dz, py = sy[s] + dy, px = sx[s] + dx;
      const bool ok = (unsigned)pz < (unsigned)p.iD[0] && (unsigned)py < (unsigned)p.iD[1] &&
                      (unsigned)px < (unsigned)p.iD[2] && kc < p.gK;
      if (ok) {
        const long long off = (((sbase[s] + pz) * p.iD[1] + py) * p.iD[2] + px) * p.in_cs + p.in_co + kc;
        ra[s] = __ldg(reinterpret_cast<const float4*>(p.in + off));
      } else {
        ra[s] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float* wt = p.w + (long long)wtap * p.wK * p.wN;
#pragma unroll
    for (int s = 0; s < LB; ++s) {
      const int f = tid + s * NT;
      const int kk = f / (TN / 4), cq = f - kk * (TN / 4);
      const int k = chunk * KC + kk, c = c0 + cq * 4;
      if (f < B_F4 && k < p.wK && c < p.wN) {
        rb[s] = __ldg(reinterpret_cast<const float4*>(wt + (long long)k * p.wN + c));
      } else {
        rb[s] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  auto store_smem = [&]() {
#pragma unroll
    for (int s = 0; s < LA; ++s) {
      const int row = a_row0 + s * (NT / 4);
      As[a_kq * 4 + 0][row] = ra[s].x;
      As[a_kq * 4 + 1][row] = ra[s].y;
      As[a_kq * 4 + 2][row] = ra[s].z;
      As[a_kq * 4 + 3][row] = ra[s].w;
    }
#pragma unroll
    for (int s = 0; s < LB; ++s) {
      const int f = tid + s * NT;
      if (f < B_F4) {
        const int kk = f / (TN / 4), cq = f - kk * (TN / 4);
        *reinterpret_cast<float4*>(&Bs[kk][cq * 4]) = rb[s];
      }
    }
  };

  const int tx = tid % TX, ty = tid / TX;
  float acc[RM][RN];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;

  if (niter > it_begin) {
    load_global(it_begin);
    store_smem();
  }
  __syncthreads();
  for (int it = it_begin; it < niter; ++it) {
    if (it + 1 < niter) load_global(it + 1);
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      float a[RM], b[RN];
#pragma unroll
      for (int i4 = 0; i4 < RM / 4; ++i4) {
        const float4 v = *reinterpret_cast<const float4*>(&As[k][i4 * (TM / (RM / 4)) + ty * 4]);
        a[i4 * 4 + 0] = v.x; a[i4 * 4 + 1] = v.y; a[i4 * 4 + 2] = v.z; a[i4 * 4 + 3] = v.w;
      }
      if constexpr (RN % 4 == 0) {
#pragma unroll
        for (int j4 = 0; j4 < RN / 4; ++j4) {
          const float4 v = *reinterpret_cast<const float4*>(&Bs[k][j4 * (TN / (RN / 4)) + tx * 4]);
          b[j4 * 4 + 0] = v.x; b[j4 * 4 + 1] = v.y; b[j4 * 4 + 2] = v.z; b[j4 * 4 + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < RN; ++j) b[j] = Bs[k][tx * RN + j];
      }
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    if (it + 1 < niter) {
      store_smem();
      __syncthreads();
    }
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int lrow = (i / 4) * (TM / (RM / 4)) + ty * 4 + (i % 4);
    const long long r = r0 + lrow;
    if (r >= rows) continue;
    int ix = (int)(r % ax.lext); long long q = r / ax.lext;
    int iy = (int)(q % ay.lext); q /= ay.lext;
    int iz = (int)(q % az.lext); const int n = (int)(q / az.lext);
    const int oz = az.l0 + iz * az.lstep, oy = ay.l0 + iy * ay.lstep, ox = ax.l0 + ix * ax.lstep;
    const long long pos = (((long long)n * p.oD[0] + oz) * p.oD[1] + oy) * p.oD[2] + ox;
    if constexpr (RN % 4 == 0) {
      if (!p.planar) {
        // vectorised channels-last store (channel offsets/strides are multiples of 4)
#pragma unroll
        for (int j4 = 0; j4 < RN / 4; ++j4) {
          const int c = c0 + j4 * (TN / (RN / 4)) + tx * 4;
          if (c >= p.gN) continue;
          float* dst = p.out + pos * p.out_cs + p.out_co + c;
          float v[4] = {acc[i][j4 * 4 + 0], acc[i][j4 * 4 + 1], acc[i][j4 * 4 + 2], acc[i][j4 * 4 + 3]};
          if (atomic_out) {
            // split-K: partial sums meet in memory (out was zeroed, or holds the value to accumulate onto)
            for (int e = 0; e < 4 && c + e < p.gN; ++e) {
              float u = v[e];
              if (split == 0 && !p.accumulate && p.bias) u += p.bias[(long long)n * p.bias_n_stride + c + e];
              atomicAdd(dst + e, u);
            }
          } else if (c + 3 < p.gN) {
            if (p.accumulate) {
              const float4 o = *reinterpret_cast<const float4*>(dst);
              v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
            } else if (p.bias) {
              const float* bp = p.bias + (long long)n * p.bias_n_stride + c;
              v[0] += bp[0]; v[1] += bp[1]; v[2] += bp[2]; v[3] += bp[3];
            }
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
            for (int e = 0; e < 4 && c + e < p.gN; ++e) {
              float u = v[e];
              if (p.accumulate) u += dst[e];
              else if (p.bias) u += p.bias[(long long)n * p.bias_n_stride + c + e];
              dst[e] = u;
            }
          }
        }
        continue;
      }
    }
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      int lcol;
      if constexpr (RN % 4 == 0) lcol = (j / 4) * (TN / (RN / 4)) + tx * 4 + (j % 4);
      else lcol = tx * RN + j;
      const int c = c0 + lcol;
      if (c >= p.gN) continue;
      float v = acc[i][j];
      long long o;
      if (p.planar) {
        const long long S = (long long)p.oD[0] * p.oD[1] * p.oD[2];
        o = ((long long)n * p.gN + c) * S + (((long long)oz * p.oD[1] + oy) * p.oD[2] + ox);
      } else {
        o = pos * p.out_cs + p.out_co + c;
      }
      if (atomic_out) {
        if (split == 0 && !p.accumulate && p.bias) v += p.bias[(long long)n * p.bias_n_stride + c];
        atomicAdd(p.out + o, v);
        continue;
      }
      if (p.accumulate) {
        v += p.out[o];
      } else if (p.bias) {
        v += p.bias[(long long)n * p.bias_n_stride + c];
      }
      p.out[o] = v;
    }
  }
}

// --------------------------------------------------------------------------
// wgrad:  dW[tap][m][n] += sum_rows P[rowP(r,tap)][m] * Q[rowQ(r,tap)][n]
// --------------------------------------------------------------------------
struct WgradParams {
  const float* P;
  const float* Q;
  float* dw;
  int N;
  int pD[3], qD[3];     // spatial dims of P / Q tensors
  int lext[3];          // row lattice
  int Kd[3];
  int pstep[3], poff0[3], poffstep[3];
  int qstep[3], qoff0[3], qoffstep[3];
  int gM, gN;           // logical Cin, Cout
  int wK, wN;           // packed dims
  int p_cs, p_co, q_cs, q_co;
  long long rows;
  long long rows_per_split;
  int mtiles, ntiles;
};

constexpr int WT = 64;   // wgrad tile (both M and N)
constexpr int WR = 16;   // rows per chunk

// TM = M (Cin) extent of the tile: 64, or 16 for the RGB stem (Cin = 3) where a 64-wide tile is 95% padding.
template <int TM>
__global__ void __launch_bounds__(NT) wgrad_kernel(const WgradParams p) {
  constexpr int RM = TM / 16;              // m values per thread
  __shared__ __align__(16) float Ps[WR][TM];
  __shared__ __align__(16) float Qs[WR][WT];
  const int tid = threadIdx.x;
  const int tap = blockIdx.z;
  const int mt = blockIdx.y / p.ntiles, nt = blockIdx.y - mt * p.ntiles;
  const int m0 = mt * TM, n0 = nt * WT;
  const int rows = (int)p.rows;            // host guarantees rows < 2^31
  const int rbeg = (int)(blockIdx.x * p.rows_per_split);
  int rend = rbeg + (int)p.rows_per_split;
  if (rend > rows) rend = rows;
  if (rbeg >= rend) return;

  int kx = tap % p.Kd[2]; int q = tap / p.Kd[2];
  int ky = q % p.Kd[1]; int kz = q / p.Kd[1];
  const int pdz = p.poff0[0] + p.poffstep[0] * kz, pdy = p.poff0[1] + p.poffstep[1] * ky,
            pdx = p.poff0[2] + p.poffstep[2] * kx;
  const int qdz = p.qoff0[0] + p.qoffstep[0] * kz, qdy = p.qoff0[1] + p.qoffstep[1] * ky,
            qdx = p.qoff0[2] + p.qoffstep[2] * kx;

  const int lrow = tid >> 4;        // 0..15
  const int lc4 = (tid & 15) * 4;   // 0..60
  const int tx = tid & 15, ty = tid >> 4;
  float acc[RM][4];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 rp, rq;
  auto load_global = [&](int rchunk) {
    const int r = rchunk + lrow;
    rp = make_float4(0.f, 0.f, 0.f, 0.f);
    rq = rp;
    if (r < rend) {
      int ix = r % p.lext[2]; int t = r / p.lext[2];
      int iy = t % p.lext[1]; t /= p.lext[1];
      int iz = t % p.lext[0]; const int n = t / p.lext[0];
      const int pz = iz * p.pstep[0] + pdz, py = iy * p.pstep[1] + pdy, px = ix * p.pstep[2] + pdx;
      const int qz = iz * p.qstep[0] + qdz, qy = iy * p.qstep[1] + qdy, qx = ix * p.qstep[2] + qdx;
      const bool okp = (unsigned)pz < (unsigned)p.pD[0] && (unsigned)py < (unsigned)p.pD[1] &&
                       (unsigned)px < (unsigned)p.pD[2];
      const bool okq = (unsigned)qz < (unsigned)p.qD[0] && (unsigned)qy < (unsigned)p.qD[1] &&
                       (unsigned)qx < (unsigned)p.qD[2];
      if (okp && okq) {
        if (lc4 < TM && m0 + lc4 < p.gM) {
          const long long off = ((((long long)n * p.pD[0] + pz) * p.pD[1] + py) * p.pD[2] + px) * p.p_cs +
                                p.p_co + m0 + lc4;
          rp = __ldg(reinterpret_cast<const float4*>(p.P + off));
        }
        if (n0 + lc4 < p.gN) {
          const long long off = ((((long long)n * p.qD[0] + qz) * p.qD[1] + qy) * p.qD[2] + qx) * p.q_cs +
                                p.q_co + n0 + lc4;
          rq = __ldg(reinterpret_cast<const float4*>(p.Q + off));
        }
      }
    }
  };

  load_global(rbeg);
  for (int rc = rbeg; rc < rend; rc += WR) {
    if (lc4 < TM) *reinterpret_cast<float4*>(&Ps[lrow][lc4]) = rp;
    *reinterpret_cast<float4*>(&Qs[lrow][lc4]) = rq;
    __syncthreads();
    if (rc + WR < rend) load_global(rc + WR);
#pragma unroll
    for (int k = 0; k < WR; ++k) {
      float av[RM];
      if constexpr (RM == 4) {
        const float4 a = *reinterpret_cast<const float4*>(&Ps[k][ty * 4]);
        av[0] = a.x; av[1] = a.y; av[2] = a.z; av[3] = a.w;
      } else {
        av[0] = Ps[k][ty];
      }
      const float4 b = *reinterpret_cast<const float4*>(&Qs[k][tx * 4]);
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* dwt = p.dw + (long long)tap * p.wK * p.wN;
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int m = m0 + ty * RM + i;
    if (m >= p.gM) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.gN) continue;
      atomicAdd(dwt + (long long)m * p.wN + n, acc[i][j]);
    }
  }
}

// --------------------------------------------------------------------------
// weight pack / unpack (multi-layer, one launch)
// --------------------------------------------------------------------------
__device__ __forceinline__ int find_item(const int64_t* offsets, int n, int64_t e) {
  int lo = 0, hi = n;   // offsets[lo] <= e < offsets[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (offsets[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void pack_weights_kernel(const crn_pack_item* items, const int64_t* offsets, int n,
                                    int64_t total) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int li = find_item(offsets, n, e);
    const crn_pack_item it = items[li];
    const int64_t l = e - offsets[li];              // index in [tap][CinP][CoutP]
    const int co = (int)(l % it.CoutP); const int64_t q = l / it.CoutP;
    const int ci = (int)(q % it.CinP); const int tap = (int)(q / it.CinP);
    float v = 0.f;
    if (ci < it.Cin && co < it.Cout) {
      const int64_t s = it.src_is_transposed ? ((int64_t)ci * it.Cout + co) * it.taps + tap
                                             : ((int64_t)co * it.Cin + ci) * it.taps + tap;
      v = it.src[s];
    }
    if (it.dst_fwd) it.dst_fwd[l] = v;
    if (it.dst_dgrad) it.dst_dgrad[((int64_t)tap * it.CoutP + co) * it.CinP + ci] = v;
  }
}

// Packed gradient [tap][CinP][CoutP] -> PyTorch layout ([Cout][Cin][taps], ConvTranspose: [Cin][Cout][taps]) as a
// shared-memory tiled transpose: both sides move whole 128-byte lines.  (The first version iterated over the
// destination and read the packed side 4 bytes at a time: 824 GB/s, 0.35 ms per step for the 94 MB of parameters.)
//   tile = 32 co x CI ci x all taps; reads: per (tap, ci) 32 consecutive co; writes: per co (ci, tap) is one contiguous
//   run of CI * taps >= 64 floats (CI = 64 for 1x1, 8 for 3x3, ..., 1 for >= 64 taps); transposed: per ci one run of 32 * taps.
constexpr int UNPACK_TILE_FLOATS = 32 * 345;           // 32 co x 343 taps (+ padding)
__device__ __forceinline__ int unpack_ci_per_tile(const crn_unpack_item& it) {
  return it.dst_is_transposed ? 1 : (it.taps >= 64 ? 1 : (64 + it.taps - 1) / it.taps);
}
__device__ __forceinline__ int unpack_tiles(const crn_unpack_item& it) {
  if (it.taps > 343) return 64;                         // generic element-wise path, 64 strided blocks
  const int CI = unpack_ci_per_tile(it);
  return ((it.Cout + 31) / 32) * ((it.Cin + CI - 1) / CI);
}
// The tiles of all items form one flat list that the blocks stride over (the items differ by 4 orders of magnitude
// in size); a block walks the item list once, its tile index only grows.
__global__ void __launch_bounds__(256) unpack_wgrads_kernel(const crn_unpack_item* items, int n) {
  __shared__ float tile[UNPACK_TILE_FLOATS];
  int li = 0;
  long long base = 0;                                   // flat index of tile 0 of item li
  crn_unpack_item it = items[0];
  int nt = unpack_tiles(it);
  for (long long tg = blockIdx.x;; tg += gridDim.x) {
    while (li < n && tg >= base + nt) {
      base += nt;
      if (++li < n) { it = items[li]; nt = unpack_tiles(it); }
    }
    if (li >= n) break;
    const int tl = (int)(tg - base);
    const int taps = it.taps;
    if (taps > 343) {                                   // no layer of the model; generic element-wise path
      const int64_t total = (int64_t)it.Cin * it.Cout * taps;
      for (int64_t l = (int64_t)tl * blockDim.x + threadIdx.x; l < total; l += (int64_t)64 * blockDim.x) {
        const int tap = (int)(l % taps); const int64_t q = l / taps;
        int ci, co;
        if (it.dst_is_transposed) { co = (int)(q % it.Cout); ci = (int)(q / it.Cout); }
        else { ci = (int)(q % it.Cin); co = (int)(q / it.Cin); }
        it.dst[l] = it.src_packed[((int64_t)tap * it.CinP + ci) * it.CoutP + co];
      }
      continue;
    }
    const int CI = unpack_ci_per_tile(it);
    const int run = CI * taps;                          // floats per co in the tile
    const int ld = run | 1;                             // odd row stride: conflict-free transposed stores
    const int tiles_co = (it.Cout + 31) / 32;
    const int co0 = (tl % tiles_co) * 32, ci0 = (tl / tiles_co) * CI;
    __syncthreads();
    // 16-byte loads (CoutP is a multiple of 4, co0 of 32), several in flight per thread: with 4-byte loads the kernel
    // was latency-bound at 1.2 TB/s
#pragma unroll 4
    for (int i = threadIdx.x; i < run * 8; i += blockDim.x) {
      const int c4 = (i & 7) * 4, r = i >> 3;           // r = tap * CI + ci_l
      const int tap = r / CI, ci_l = r - tap * CI;
      const int co = co0 + c4, ci = ci0 + ci_l;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (co < it.CoutP && ci < it.Cin)
        v = __ldg(reinterpret_cast<const float4*>(it.src_packed + ((int64_t)tap * it.CinP + ci) * it.CoutP + co));
      float* dst = tile + c4 * ld + ci_l * taps + tap;
      dst[0] = v.x; dst[ld] = v.y; dst[2 * ld] = v.z; dst[3 * ld] = v.w;
    }
    __syncthreads();
    if (it.dst_is_transposed) {                         // [ci][co][tap]: one run of 32 * taps floats per ci
      float* dst = it.dst + ((int64_t)ci0 * it.Cout + co0) * taps;
      const int lim = (it.Cout - co0 < 32 ? it.Cout - co0 : 32) * taps;
      for (int i = threadIdx.x; i < lim; i += blockDim.x) {
        const int co_l = i / taps, tap = i - co_l * taps;
        dst[i] = tile[co_l * ld + tap];
      }
    } else {                                            // [co][ci][tap]: per co one run of CI * taps floats
      const int lim = (it.Cin - ci0 < CI ? it.Cin - ci0 : CI) * taps;
      for (int i = threadIdx.x; i < run * 32; i += blockDim.x) {
        const int co_l = i / run, j = i - co_l * run;
        const int co = co0 + co_l;
        if (co < it.Cout && j < lim) it.dst[((int64_t)co * it.Cin + ci0) * taps + j] = tile[co_l * ld + j];
      }
    }
  }
}

// --------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------
bool check_desc(const crn_conv_desc* d) {
  if (!d) return false;
  if (d->N <= 0 || d->Cin <= 0 || d->Cout <= 0) return false;
  if (d->x_cs % 4 || d->x_co % 4 || d->y_cs % 4 || d->y_co % 4) return false;
  if (d->CinP % 4 || d->CoutP % 4 || d->CinP < d->Cin || d->CoutP < d->Cout) return false;
  if (d->stride < 1 || d->kD < 1 || d->kH < 1 || d->kW < 1) return false;
  return true;
}

void axis_setup(const crn_conv_desc* d, int s[3], int pad[3], int K[3], int iD[3], int oD[3]) {
  K[0] = d->kD; K[1] = d->kH; K[2] = d->kW;
  iD[0] = d->iD; iD[1] = d->iH; iD[2] = d->iW;
  oD[0] = d->oD; oD[1] = d->oH; oD[2] = d->oW;
  for (int a = 0; a < 3; ++a) {
    const bool trivial = (K[a] == 1 && iD[a] == 1 && oD[a] == 1);
    s[a] = trivial ? 1 : d->stride;
    pad[a] = (K[a] > 1) ? d->pad : 0;
  }
}

template <int TM, int TN, int RM, int RN>
int launch_gather_cfg(GatherParams p, long long max_rows, int nclasses, int taps, cudaStream_t st) {
  const long long blocks = crn_ceil_div(max_rows, TM) * crn_ceil_div(p.gN, TN) * nclasses;
  // split-K when the output tile grid cannot fill the machine (small grids with a huge tap x channel
  // reduction: encoder stages 4-5 at small batch, decoder stages 1-3).  Needs a dense, exclusively owned
  // output so that it can be zeroed first.
  p.ksplit = 1;
  const long long niter = (long long)taps * crn_ceil_div(p.gK, KC);
  const bool dense_out = !p.planar && p.out_co == 0 && p.out_cs == (p.gN + 3) / 4 * 4;
  if (blocks < kNumSMs && niter >= 8 && dense_out && !(crn_get_flags() & 16)) {
    long long ks = crn_ceil_div(2LL * kNumSMs, blocks);
    if (ks > niter / 4) ks = niter / 4;
    if (ks > 64) ks = 64;
    if (ks > 1) {
      p.ksplit = (int)ks;
      if (!p.accumulate) {
        long long out_rows = p.N;
        for (int a = 0; a < 3; ++a) out_rows *= p.oD[a];
        cudaMemsetAsync(p.out, 0, sizeof(float) * out_rows * p.out_cs, st);
      }
    }
  }
  dim3 grid((unsigned)crn_ceil_div(max_rows, TM), (unsigned)crn_ceil_div(p.gN, TN), (unsigned)(nclasses * p.ksplit));
  gather_gemm_kernel<TM, TN, RM, RN><<<grid, NT, 0, st>>>(p);
  CRN_LAUNCH_CHECK("gather_gemm");
  return CRN_OK;
}

int launch_gather(const GatherParams& p, cudaStream_t st) {
  // rows of the largest class, taps of the largest class
  long long max_rows = p.N;
  int nclasses = 1, taps = 1;
  for (int a = 0; a < 3; ++a) {
    if (p.class_mode) {
      max_rows *= (p.oD[a] + p.s[a] - 1) / p.s[a];
      nclasses *= p.s[a];
      taps *= (p.Kd[a] + p.s[a] - 1) / p.s[a];
    } else {
      max_rows *= p.oD[a];
      taps *= p.Kd[a];
    }
  }
  if (max_rows <= 0) return CRN_OK;
  const int n = p.gN;
  if (n <= 16) return launch_gather_cfg<256, 16, 8, 2>(p, max_rows, nclasses, taps, st);
  if (n <= 32) return launch_gather_cfg<256, 32, 8, 4>(p, max_rows, nclasses, taps, st);
  // the largest register tile the problem can use; split-K restores the parallelism on small grids
  if (max_rows <= 64) return launch_gather_cfg<64, 64, 4, 4>(p, max_rows, nclasses, taps, st);
  if (n > 64 && max_rows >= 256) return launch_gather_cfg<128, 128, 8, 8>(p, max_rows, nclasses, taps, st);
  return launch_gather_cfg<128, 64, 8, 4>(p, max_rows, nclasses, taps, st);
}

}  // namespace

int crn_rowdirect_try(const crn_conv_desc* d, int kind, const float* in, const float* w, const float* bias,
                      float* out, int accumulate, cudaStream_t st);

extern "C" int crn_conv_fwd(const crn_conv_desc* d, const float* x, const float* w_fwd,
                            const float* bias, float* y, int32_t accumulate, void* stream) {
  CRN_REQUIRE(check_desc(d), "crn_conv_fwd: bad descriptor");
  CRN_REQUIRE(x && w_fwd && y, "crn_conv_fwd: null pointer");
  {   // small-channel 3-D decoder layers: row-direct kernel (conv_rowdirect.cu)
    const int rc = crn_rowdirect_try(d, 0, x, w_fwd, bias, y, accumulate, crn_stream(stream));
    if (rc != CRN_ERR_UNSUPPORTED) return rc;
  }
  GatherParams p{};
  p.in = x; p.w = w_fwd; p.bias = accumulate ? nullptr : bias; p.out = y;
  p.N = d->N;
  axis_setup(d, p.s, p.pad, p.Kd, p.iD, p.oD);
  p.gK = d->Cin; p.gN = d->Cout; p.wK = d->CinP; p.wN = d->CoutP;
  p.in_cs = d->x_cs; p.in_co = d->x_co; p.out_cs = d->y_cs; p.out_co = d->y_co;
  p.class_mode = d->transposed ? 1 : 0;
  p.accumulate = accumulate; p.planar = d->y_planar; p.bias_n_stride = d->bias_n_stride;
  return launch_gather(p, crn_stream(stream));
}

extern "C" int crn_conv_dgrad(const crn_conv_desc* d, const float* dy, const float* w_dgrad,
                              float* dx, int32_t accumulate, void* stream) {
  CRN_REQUIRE(check_desc(d), "crn_conv_dgrad: bad descriptor");
  CRN_REQUIRE(dy && w_dgrad && dx, "crn_conv_dgrad: null pointer");
  CRN_REQUIRE(!d->y_planar, "crn_conv_dgrad: planar dy unsupported");
  {
    const int rc = crn_rowdirect_try(d, 1, dy, w_dgrad, nullptr, dx, accumulate, crn_stream(stream));
    if (rc != CRN_ERR_UNSUPPORTED) return rc;
  }
  GatherParams p{};
  p.in = dy; p.w = w_dgrad; p.bias = nullptr; p.out = dx;
  p.N = d->N;
  int iD[3], oD[3];
  axis_setup(d, p.s, p.pad, p.Kd, iD, oD);
  for (int a = 0; a < 3; ++a) { p.iD[a] = oD[a]; p.oD[a] = iD[a]; }   // roles swap
  p.gK = d->Cout; p.gN = d->Cin; p.wK = d->CoutP; p.wN = d->CinP;
  p.in_cs = d->y_cs; p.in_co = d->y_co; p.out_cs = d->x_cs; p.out_co = d->x_co;
  p.class_mode = d->transposed ? 0 : 1;
  p.accumulate = accumulate; p.planar = 0; p.bias_n_stride = 0;
  return launch_gather(p, crn_stream(stream));
}

int crn_wgrad_row_try(const crn_conv_desc* d, const float* x, const float* dy, float* dw, cudaStream_t st);
int crn_wgrad_stem_try(const crn_conv_desc* d, const float* x, const float* dy, float* dw, cudaStream_t st);

extern "C" int crn_conv_wgrad(const crn_conv_desc* d, const float* x, const float* dy,
                              float* dw_packed, void* stream) {
  CRN_REQUIRE(check_desc(d), "crn_conv_wgrad: bad descriptor");
  CRN_REQUIRE(x && dy && dw_packed, "crn_conv_wgrad: null pointer");
  CRN_REQUIRE(!d->y_planar, "crn_conv_wgrad: planar dy unsupported");
  {   // RGB stem (conv_wgrad_stem.cu)
    const int rc = crn_wgrad_stem_try(d, x, dy, dw_packed, crn_stream(stream));
    if (rc != CRN_ERR_UNSUPPORTED) return rc;
  }
  {   // small-channel 3-D decoder layers: tap-row kernel (conv_wgrad_row.cu)
    const int rc = crn_wgrad_row_try(d, x, dy, dw_packed, crn_stream(stream));
    if (rc != CRN_ERR_UNSUPPORTED) return rc;
  }
  WgradParams p{};
  p.P = x; p.Q = dy; p.dw = dw_packed; p.N = d->N;
  int s[3], pad[3], K[3], iD[3], oD[3];
  axis_setup(d, s, pad, K, iD, oD);
  for (int a = 0; a < 3; ++a) {
    p.pD[a] = iD[a]; p.qD[a] = oD[a]; p.Kd[a] = K[a];
    if (!d->transposed) {   // rows = output positions; x at o*s - pad + k
      p.lext[a] = oD[a];
      p.pstep[a] = s[a]; p.poff0[a] = -pad[a]; p.poffstep[a] = 1;
      p.qstep[a] = 1; p.qoff0[a] = 0; p.qoffstep[a] = 0;
    } else {                // rows = input positions; dy at i*s - pad + k
      p.lext[a] = iD[a];
      p.pstep[a] = 1; p.poff0[a] = 0; p.poffstep[a] = 0;
      p.qstep[a] = s[a]; p.qoff0[a] = -pad[a]; p.qoffstep[a] = 1;
    }
  }
  p.gM = d->Cin; p.gN = d->Cout; p.wK = d->CinP; p.wN = d->CoutP;
  p.p_cs = d->x_cs; p.p_co = d->x_co; p.q_cs = d->y_cs; p.q_co = d->y_co;
  p.rows = (long long)p.N * p.lext[0] * p.lext[1] * p.lext[2];
  if (p.rows <= 0) return CRN_OK;
  CRN_REQUIRE(p.rows < 0x7fffffffLL, "crn_conv_wgrad: too many rows");
  const bool narrow = p.gM <= 16;            // RGB stem: a 16-wide M tile
  p.mtiles = narrow ? 1 : (int)crn_ceil_div(p.gM, WT);
  p.ntiles = (int)crn_ceil_div(p.gN, WT);
  const int taps = K[0] * K[1] * K[2];
  const long long tiles = (long long)p.mtiles * p.ntiles * taps;
  long long nsplit = crn_ceil_div(4LL * kNumSMs, tiles);
  const long long max_split = crn_ceil_div(p.rows, 4 * WR);
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  p.rows_per_split = crn_ceil_div(crn_ceil_div(p.rows, nsplit), WR) * WR;
  nsplit = crn_ceil_div(p.rows, p.rows_per_split);
  CRN_REQUIRE(taps <= 65535 && (long long)p.mtiles * p.ntiles <= 65535, "crn_conv_wgrad: grid too large");
  dim3 grid((unsigned)nsplit, (unsigned)(p.mtiles * p.ntiles), (unsigned)taps);
  if (narrow) wgrad_kernel<16><<<grid, NT, 0, crn_stream(stream)>>>(p);
  else wgrad_kernel<64><<<grid, NT, 0, crn_stream(stream)>>>(p);
  CRN_LAUNCH_CHECK("wgrad");
  return CRN_OK;
}

extern "C" int crn_pack_weights(const crn_pack_item* items, const int64_t* offsets, int32_t n,
                                int64_t total, void* stream) {
  CRN_REQUIRE(items && offsets && n > 0 && total > 0, "crn_pack_weights: bad args");
  const int blocks = (int)(crn_ceil_div(total, 256) < 8 * kNumSMs ? crn_ceil_div(total, 256) : 8 * kNumSMs);
  pack_weights_kernel<<<blocks, 256, 0, crn_stream(stream)>>>(items, offsets, n, total);
  CRN_LAUNCH_CHECK("pack_weights");
  return CRN_OK;
}

extern "C" int crn_unpack_wgrads(const crn_unpack_item* items, const int64_t* offsets, int32_t n,
                                 int64_t total, void* stream) {
  CRN_REQUIRE(items && offsets && n > 0 && total > 0, "crn_unpack_wgrads: bad args");
  (void)offsets; (void)total;                           // kept in the signature: element prefix sums of the items
  unpack_wgrads_kernel<<<4 * kNumSMs, 256, 0, crn_stream(stream)>>>(items, n);
  CRN_LAUNCH_CHECK("unpack_wgrads");
  return CRN_OK;
}
