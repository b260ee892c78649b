// Direct ("row") convolution for the small-channel 3-D decoder layers, fp32 FFMA.
//
// Covers, with one kernel template, for Conv3d k=5 s=1 and ConvTranspose3d k=7 s=2
// (model/reconstruction_decoder.py:57-95 of the reference):
//   conv fwd, conv dgrad           -> stride-1 gather, 5 taps per axis
//   convT fwd (8 parity classes)   -> stride-1 gather inside a class, 3 or 4 taps per axis
//   convT dgrad                    -> stride-2 gather, 7 taps per axis
//
// The implicit-GEMM kernel (conv_generic.cu) re-stages the input once per tap
// (125-343x) and pads N to >= 16; with Cout = 2..32 that wastes most of its FMAs and L2
// traffic.  Here a block stages one input z-plane slab [rows][x][K-chunk] ONCE per
// (kz, K-chunk) in its global channels-last layout together with that stage's weights;
// a thread owns RV consecutive outputs along x times RC output channels and slides a
// register window along x, so all x-taps reuse the same shared-memory loads:
// ~20 FMAs per LDS, FFMA-bound.
#include "common.cuh"

namespace {

struct RDParams {
  const float* in;
  const float* w;
  const float* bias;
  float* out;
  int N;
  int iD[3], oD[3], Kd[3], s[3], pad[3];
  int gK, gN, wK, wN;
  int in_cs, in_co, out_cs, out_co;
  int class_mode, accumulate, planar, bias_n_stride;
  int TY, XG, NCG;     // lattice rows per tile, thread groups along x, cout groups per block
  int KCH;             // K channels staged per step (multiple of KW)
  int g_shift;         // log2(KCH / KW) or -1
  int SR, SWd, SWs;    // staged rows, staged width (voxels), skewed width
};

struct Axis {
  int l0, lstep, lext, istep, nk, k0, kstep, off0, offstep;
};

__device__ __forceinline__ Axis make_axis(const RDParams& p, int a, int c) {
  Axis q;
  if (!p.class_mode) {
    q.l0 = 0; q.lstep = 1; q.lext = p.oD[a];
    q.istep = p.s[a];
    q.nk = p.Kd[a]; q.k0 = 0; q.kstep = 1; q.off0 = -p.pad[a]; q.offstep = 1;
  } else {
    const int s = p.s[a];
    q.l0 = c; q.lstep = s; q.lext = (p.oD[a] - c + s - 1) / s;
    q.istep = 1;
    q.k0 = (c + p.pad[a]) % s;
    q.kstep = s;
    q.nk = q.k0 < p.Kd[a] ? (p.Kd[a] - q.k0 + s - 1) / s : 0;
    q.off0 = (c + p.pad[a] - q.k0) / s;
    q.offstep = -1;
  }
  return q;
}

// column c of a staged row lives at c + c/8 (one pad voxel per 8): kills the 8-way bank
// conflict between the threads of a quarter-warp, whose windows start 8 voxels apart.
__host__ __device__ __forceinline__ constexpr int skew(int c) { return c + (c >> 3); }

template <int KW>
__device__ __forceinline__ void lds_group(const float* p, float (&v)[KW]) {
  if constexpr (KW == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  }
}

template <int KW>
__device__ __forceinline__ void cp_async_group(float* dst, const float* src, bool ok) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  if constexpr (KW == 4) {
    const int sz = ok ? 16 : 0;                    // src-size 0 => zeros, nothing read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(sz) : "memory");
  } else {
    const int sz = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(sz) : "memory");
  }
}

// RV outputs along x per thread, RC output channels per thread, KW = K-group width,
// NKX = taps along x (window columns), ISTEP = input step per output (y and x).
template <int RV, int RC, int KW, int NKX, int ISTEP>
__global__ void __launch_bounds__(256) rowdirect_kernel(const RDParams p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int WN = (RV - 1) * ISTEP + NKX;        // window columns
  const int tid = threadIdx.x;
  int cz = 0, cy = 0, cx = 0;
  if (p.class_mode) {
    int c = blockIdx.z;
    cx = c % p.s[2]; c /= p.s[2];
    cy = c % p.s[1]; c /= p.s[1];
    cz = c;
  }
  const Axis az = make_axis(p, 0, cz), ay = make_axis(p, 1, cy), ax = make_axis(p, 2, cx);
  const int yblocks = (ay.lext + p.TY - 1) / p.TY;
  int t = blockIdx.x;
  const int yb = t % yblocks; t /= yblocks;
  const int iz = t % az.lext; const int n = t / az.lext;
  const int y0 = yb * p.TY;
  const int ncols = p.NCG * RC;                      // output channels per block
  const int co_blk = blockIdx.y * ncols;

  const int xg = tid % p.XG;
  const int yy = (tid / p.XG) % p.TY;
  const int cgl = tid / (p.XG * p.TY);
  const bool active = cgl < p.NCG && (y0 + yy) < ay.lext;

  // smallest input offset per axis; staged column dxc <-> offset xmin + dxc <-> tap jx
  const int ymin = ay.offstep > 0 ? ay.off0 : ay.off0 - (ay.nk - 1);
  const int xmin = ax.offstep > 0 ? ax.off0 : ax.off0 - (ax.nk - 1);

  float* inS = smem;                                          // [SR][SWs][KCH]
  float* wS = smem + (((size_t)p.SR * p.SWs * p.KCH + 3) & ~(size_t)3);   // [nky][NKX][KCH][ncols]

  float acc[RV][RC];
#pragma unroll
  for (int j = 0; j < RV; ++j)
#pragma unroll
    for (int c = 0; c < RC; ++c) acc[j][c] = 0.f;

  const int groups = p.KCH / KW;
  for (int jz = 0; jz < az.nk; ++jz) {
    const int pz = iz * az.istep + az.off0 + az.offstep * jz;
    if (pz < 0 || pz >= p.iD[0]) continue;            // block-uniform
    const int kz = az.k0 + az.kstep * jz;
    for (int kc0 = 0; kc0 < p.gK; kc0 += p.KCH) {
      __syncthreads();
      // ---- stage the input slab with cp.async (zero-fill outside the grid = conv padding / missing taps);
      // (row, unit) advance incrementally per thread: no per-element divisions
      {
        const int upr = p.SWd * groups;               // K-group units per staged row
        int rr = tid / upr, cu = tid - rr * upr;
        const long long plane = ((long long)n * p.iD[0] + pz) * p.iD[1];
        const int py0 = y0 * ISTEP + ymin;
        while (rr < p.SR) {
          int c, g;
          if (p.g_shift >= 0) { c = cu >> p.g_shift; g = cu & (groups - 1); } else { c = cu / groups; g = cu - c * groups; }
          const int py = py0 + rr, px = xmin + c;
          const int k = kc0 + g * KW;
          const bool ok = (unsigned)py < (unsigned)p.iD[1] && (unsigned)px < (unsigned)p.iD[2] && k < p.gK;
          const long long off = ok ? ((plane + py) * p.iD[2] + px) * p.in_cs + p.in_co + k : 0;
          float* dst = inS + ((size_t)rr * p.SWs + skew(c)) * p.KCH + g * KW;
          cp_async_group<KW>(dst, p.in + off, ok);
          cu += blockDim.x;
          while (cu >= upr) { cu -= upr; ++rr; }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
      }
      // ---- stage this step's weights in window order: wS[ryc][dxc][k][co]; a warp per (ryc, dxc) tap
      {
        const int per_tap = p.KCH * ncols;
        const int ntap = ay.nk * NKX;
        for (int tp = tid >> 5; tp < ntap; tp += (int)(blockDim.x >> 5)) {
          const int ryc = tp / NKX, dxc = tp - ryc * NKX;
          const int jy = ay.offstep > 0 ? ryc : ay.nk - 1 - ryc;
          const int jx = ax.offstep > 0 ? dxc : ax.nk - 1 - dxc;
          const int ky = ay.k0 + ay.kstep * jy, kx = ax.k0 + ax.kstep * jx;
          const long long tap = ((long long)kz * p.Kd[1] + ky) * p.Kd[2] + kx;
          const float* wsrc = p.w + (tap * p.wK + kc0) * p.wN + co_blk;
          float* wdst = wS + (size_t)tp * per_tap;
          const bool tap_ok = dxc < ax.nk;
          for (int e = (int)(tid & 31); e < per_tap; e += 32) {
            const int k = e / ncols, c = e - k * ncols;
            float v = 0.f;
            if (tap_ok && kc0 + k < p.wK && co_blk + c < p.wN) v = __ldg(wsrc + (long long)k * p.wN + c);
            wdst[e] = v;
          }
        }
      }
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      __syncthreads();
      if (!active) continue;
      for (int ryc = 0; ryc < ay.nk; ++ryc) {
        const float* srow = inS + (size_t)(yy * ISTEP + ryc) * p.SWs * p.KCH;
        const int colbase = xg * RV * ISTEP;           // multiple of 8
        const int sbase = colbase + (colbase >> 3);
        for (int g = 0; g < groups; ++g) {
          float win[WN][KW];
#pragma unroll
          for (int c = 0; c < WN; ++c) lds_group<KW>(srow + (size_t)(sbase + skew(c)) * p.KCH + g * KW, win[c]);
          const float* wrow = wS + ((size_t)(ryc * NKX) * p.KCH + g * KW) * ncols + cgl * RC;
#pragma unroll
          for (int dxc = 0; dxc < NKX; ++dxc) {
#pragma unroll
            for (int kk = 0; kk < KW; ++kk) {
              float wv[RC];
              const float* wp = wrow + ((size_t)dxc * p.KCH + kk) * ncols;
              if constexpr (RC % 4 == 0) {
#pragma unroll
                for (int c4 = 0; c4 < RC / 4; ++c4) {
                  const float4 q = *reinterpret_cast<const float4*>(wp + c4 * 4);
                  wv[c4 * 4 + 0] = q.x; wv[c4 * 4 + 1] = q.y; wv[c4 * 4 + 2] = q.z; wv[c4 * 4 + 3] = q.w;
                }
              } else {
#pragma unroll
                for (int c = 0; c < RC; ++c) wv[c] = wp[c];
              }
#pragma unroll
              for (int j = 0; j < RV; ++j)
#pragma unroll
                for (int c = 0; c < RC; ++c) acc[j][c] = fmaf(win[j * ISTEP + dxc][kk], wv[c], acc[j][c]);
            }
          }
        }
      }
    }
  }
  if (!active) return;
  // ---- epilogue
  const int oz = az.l0 + iz * az.lstep, oy = ay.l0 + (y0 + yy) * ay.lstep;
  const long long S = (long long)p.oD[0] * p.oD[1] * p.oD[2];
#pragma unroll
  for (int j = 0; j < RV; ++j) {
    const int ix = xg * RV + j;
    const int ox = ax.l0 + ix * ax.lstep;
    const long long pos = (((long long)n * p.oD[0] + oz) * p.oD[1] + oy) * p.oD[2] + ox;
#pragma unroll
    for (int c = 0; c < RC; ++c) {
      const int co = co_blk + cgl * RC + c;
      if (co >= p.gN) continue;
      float v = acc[j][c];
      const long long o = p.planar ? ((long long)n * p.gN + co) * S + (((long long)oz * p.oD[1] + oy) * p.oD[2] + ox)
                                   : pos * p.out_cs + p.out_co + co;
      if (p.accumulate) v += p.out[o];
      else if (p.bias) v += p.bias[(long long)n * p.bias_n_stride + co];
      p.out[o] = v;
    }
  }
}

template <int RV, int RC, int KW, int NKX, int ISTEP>
int launch(RDParams p, int lext_x, int lext_y, int lext_z, int nclasses, int max_nky, cudaStream_t st) {
  auto kern = rowdirect_kernel<RV, RC, KW, NKX, ISTEP>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    configured = true;
  }
  p.XG = lext_x / RV;
  // cout groups per block: all of them if they fit in 256 threads with >= 4 rows
  const int total_groups = (p.gN + RC - 1) / RC;
  int ncg = total_groups;
  while (ncg > 1 && p.XG * 4 * ncg > 256) ncg = (ncg + 1) / 2;
  p.NCG = ncg;
  int ty = 256 / (p.XG * ncg);
  if (ty > lext_y) ty = lext_y;
  if (ty > 16) ty = 16;
  if (ty < 1) return CRN_ERR_UNSUPPORTED;
  // K chunk: as much as fits in ~96 KB together with the weights
  int kch = ((p.gK + KW - 1) / KW) * KW;
  if (kch > 16) kch = 16;
  size_t bytes = 0;
  for (;;) {
    p.TY = ty; p.KCH = kch;
    p.SR = (ty - 1) * ISTEP + max_nky;
    p.SWd = (lext_x - 1) * ISTEP + NKX;
    p.SWs = skew(p.SWd) + 1;
    bytes = sizeof(float) * ((((size_t)p.SR * p.SWs * kch + 3) & ~(size_t)3) + (size_t)max_nky * NKX * kch * ncg * RC);
    if (bytes <= 100 * 1024) break;
    if (kch > KW && kch > 4) kch = (kch / 2 + KW - 1) / KW * KW;
    else if (ty > 1) ty = (ty + 1) / 2;
    else return CRN_ERR_UNSUPPORTED;
  }
  {
    const int g = p.KCH / KW;
    int sh = 0;
    while ((1 << sh) < g) ++sh;
    p.g_shift = (1 << sh) == g ? sh : -1;
  }
  const int threads = ((p.XG * ty * ncg + 31) / 32) * 32;
  if (threads > 256) return CRN_ERR_UNSUPPORTED;
  const long long tiles = (long long)p.N * lext_z * ((lext_y + ty - 1) / ty);
  const int gy = (total_groups + ncg - 1) / ncg;
  if (tiles > 0x7fffffffLL) return CRN_ERR_UNSUPPORTED;
  kern<<<dim3((unsigned)tiles, (unsigned)gy, (unsigned)nclasses), threads, bytes, st>>>(p);
  CRN_LAUNCH_CHECK("rowdirect");
  return CRN_OK;
}

}  // namespace

// kind: 0 = forward (conv or convT), 1 = dgrad.  Returns CRN_ERR_UNSUPPORTED outside the envelope.
int crn_rowdirect_try(const crn_conv_desc* d, int kind, const float* in, const float* w, const float* bias,
                      float* out, int accumulate, cudaStream_t st) {
  if (crn_get_flags() & 1) return CRN_ERR_UNSUPPORTED;
  const bool conv5 = !d->transposed && d->kD == 5 && d->kH == 5 && d->kW == 5 && d->stride == 1 && d->pad == 2;
  const bool convT7 = d->transposed && d->kD == 7 && d->kH == 7 && d->kW == 7 && d->stride == 2 && d->pad == 3;
  if (!conv5 && !convT7) return CRN_ERR_UNSUPPORTED;
  RDParams p{};
  p.in = in; p.w = w; p.bias = accumulate ? nullptr : bias; p.out = out;
  p.N = d->N; p.accumulate = accumulate; p.bias_n_stride = d->bias_n_stride;
  const int iD[3] = {d->iD, d->iH, d->iW}, oD[3] = {d->oD, d->oH, d->oW};
  for (int a = 0; a < 3; ++a) {
    p.Kd[a] = d->kD; p.s[a] = d->stride; p.pad[a] = d->pad;
    p.iD[a] = kind == 0 ? iD[a] : oD[a];
    p.oD[a] = kind == 0 ? oD[a] : iD[a];
  }
  if (kind == 0) {
    p.gK = d->Cin; p.gN = d->Cout; p.wK = d->CinP; p.wN = d->CoutP;
    p.in_cs = d->x_cs; p.in_co = d->x_co; p.out_cs = d->y_cs; p.out_co = d->y_co;
    p.class_mode = d->transposed ? 1 : 0; p.planar = d->y_planar;
  } else {
    if (d->y_planar) return CRN_ERR_UNSUPPORTED;
    p.gK = d->Cout; p.gN = d->Cin; p.wK = d->CoutP; p.wN = d->CinP;
    p.in_cs = d->y_cs; p.in_co = d->y_co; p.out_cs = d->x_cs; p.out_co = d->x_co;
    p.class_mode = d->transposed ? 0 : 1; p.planar = 0;
  }
  // small channel counts only: the implicit-GEMM kernel is the better tool for wide layers
  if (p.gN > 64 || p.gK > 64) return CRN_ERR_UNSUPPORTED;
  if (p.in_cs % 4 || p.in_co % 4) return CRN_ERR_UNSUPPORTED;
  const int s = p.class_mode ? d->stride : 1;
  int lext[3];
  for (int a = 0; a < 3; ++a) {
    if (p.class_mode && p.oD[a] % s) return CRN_ERR_UNSUPPORTED;
    lext[a] = p.oD[a] / s;
  }
  const int nclasses = p.class_mode ? s * s * s : 1;
  const bool strided_gather = !p.class_mode && d->stride == 2;      // convT dgrad
  // K-group width: 2 when the gathered tensor only has 2 (padded to 4) channels
  const bool k2 = p.gK <= 2;
  // Dispatch envelope = where this kernel beat the implicit-GEMM kernel on a B200 (profiles/r01_layers.md):
  // every tiny-Cout layer, the k=5 forward convs, the narrow transposed forward convs, and the dgrad of
  // the logits layer.  Everything else stays on conv_generic.cu.
  if (conv5) {                       // fwd (direct) or dgrad (class mode, one class): 5 taps, step 1
    if (lext[2] % 8 || lext[2] > 64) return CRN_ERR_UNSUPPORTED;
    if (p.gN <= 4) return launch<8, 4, 4, 5, 1>(p, lext[2], lext[1], lext[0], nclasses, 5, st);
    if (kind != 0 || p.gN > 32) return CRN_ERR_UNSUPPORTED;
    return launch<8, 8, 4, 5, 1>(p, lext[2], lext[1], lext[0], nclasses, 5, st);
  }
  if (!strided_gather) {             // convT fwd: 8 classes, 3 or 4 taps per axis, step 1
    if (lext[2] % 8 || lext[2] > 64) return CRN_ERR_UNSUPPORTED;
    if (p.gN <= 2) return launch<8, 2, 4, 4, 1>(p, lext[2], lext[1], lext[0], nclasses, 4, st);
    if (p.gN <= 4) return launch<8, 4, 4, 4, 1>(p, lext[2], lext[1], lext[0], nclasses, 4, st);
    if (p.gN > 16 || p.gK > 32) return CRN_ERR_UNSUPPORTED;
    return launch<8, 4, 4, 4, 1>(p, lext[2], lext[1], lext[0], nclasses, 4, st);
  }
  // convT dgrad: 7 taps, input step 2
  if (lext[2] % 4 || lext[2] > 64) return CRN_ERR_UNSUPPORTED;
  if (k2) return launch<4, 8, 2, 7, 2>(p, lext[2], lext[1], lext[0], nclasses, 7, st);
  if (crn_get_flags() & 8) return launch<4, 8, 4, 7, 2>(p, lext[2], lext[1], lext[0], nclasses, 7, st);
  return CRN_ERR_UNSUPPORTED;
}
