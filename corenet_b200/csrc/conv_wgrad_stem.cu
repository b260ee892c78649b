// Weight gradient of the ResNet-50 stem: Conv2d(3 -> 64, k=7, s=2, p=3) on the 256x256 image
// (model/resnet50.py:122-131).  dW[ky][kx][ci][co] = sum_{n, oy, ox} x[n][2oy-3+ky][2ox-3+kx][ci] * dy[n][oy][ox][co].
// Only 4 (padded RGB) input channels: no tensor-core tile fits, but the problem is tiny (0.8 GMAC at B = 4) -- the
// generic split-K GEMM spent 0.74 ms on it because its 16-wide M tile is mostly padding.  Here:
//   * CTA = one output image row at a time (grid-stride over N * oH rows): the 7 input rows (zero-padded borders)
//     and the dy row are staged in shared memory once;
//   * thread = (ky, pair of output channels): 7 kx x 4 ci x 2 co = 56 register accumulators, the 7-pixel input
//     window slides along x in registers (2 new float4 per output pixel, all threads of a warp read the same
//     address -> broadcast), dy is read as float2 (conflict-free);
//   * one atomic flush of the 12544 sums per CTA at the end.
#include "common.cuh"

namespace {

constexpr int SW_THREADS = 7 * 32;     // (ky, co pair)

template <int OW>
__global__ void __launch_bounds__(SW_THREADS) wgrad_stem_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                float* __restrict__ dw, int N, int iH, int oH, int x_cs,
                                                                int y_cs, int y_co, int CinP, int CoutP) {
  constexpr int IW = 2 * OW, XW = IW + 8;                 // staged input row: pixels -3 .. IW+4
  constexpr int OC = 64;                                  // output pixels per staged dy chunk (static smem < 48 KB)
  __shared__ float4 xs[7][XW];
  __shared__ float ds[OC][64];
  const int tid = threadIdx.x, ky = tid >> 5, cp = tid & 31;
  float acc[7][4][2];
#pragma unroll
  for (int a = 0; a < 7; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.f;
  for (int row = blockIdx.x; row < N * oH; row += gridDim.x) {
    const int n = row / oH, oy = row - n * oH;
    __syncthreads();
    for (int i = tid; i < 7 * XW; i += SW_THREADS) {
      const int r = i / XW, px = i - r * XW - 3;
      const int iy = 2 * oy - 3 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((unsigned)iy < (unsigned)iH && (unsigned)px < (unsigned)IW)
        v = __ldg(reinterpret_cast<const float4*>(x + (((long long)n * iH + iy) * IW + px) * x_cs));
      xs[r][px + 3] = v;
    }
    for (int ox0 = 0; ox0 < OW; ox0 += OC) {
      if (ox0) __syncthreads();
      for (int i = tid; i < OC * 16; i += SW_THREADS) {
        const int ox = i >> 4, c4 = i & 15;
        *reinterpret_cast<float4*>(&ds[ox][c4 * 4]) =
            __ldg(reinterpret_cast<const float4*>(dy + (((long long)n * oH + oy) * OW + ox0 + ox) * y_cs + y_co + c4 * 4));
      }
      __syncthreads();
      // window w[kx] = x[2ox - 3 + kx]; staged index = 2ox + kx
      float4 w[7];
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) w[kx] = xs[ky][2 * ox0 + kx];
#pragma unroll 2
      for (int o = 0; o < OC; ++o) {
        const float2 g = *reinterpret_cast<const float2*>(&ds[o][cp * 2]);
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          acc[kx][0][0] = fmaf(w[kx].x, g.x, acc[kx][0][0]); acc[kx][0][1] = fmaf(w[kx].x, g.y, acc[kx][0][1]);
          acc[kx][1][0] = fmaf(w[kx].y, g.x, acc[kx][1][0]); acc[kx][1][1] = fmaf(w[kx].y, g.y, acc[kx][1][1]);
          acc[kx][2][0] = fmaf(w[kx].z, g.x, acc[kx][2][0]); acc[kx][2][1] = fmaf(w[kx].z, g.y, acc[kx][2][1]);
          acc[kx][3][0] = fmaf(w[kx].w, g.x, acc[kx][3][0]); acc[kx][3][1] = fmaf(w[kx].w, g.y, acc[kx][3][1]);
        }
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) w[kx] = w[kx + 2];
        w[5] = xs[ky][2 * (ox0 + o) + 7];
        w[6] = xs[ky][2 * (ox0 + o) + 8];
      }
    }
  }
#pragma unroll
  for (int kx = 0; kx < 7; ++kx)
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      if (ci >= CinP) continue;
      float* dst = dw + ((long long)(ky * 7 + kx) * CinP + ci) * CoutP + cp * 2;
      atomicAdd(dst, acc[kx][ci][0]);
      atomicAdd(dst + 1, acc[kx][ci][1]);
    }
}

}  // namespace

// Returns CRN_ERR_UNSUPPORTED unless this is the stem shape (2-D k=7 s=2 p=3, <= 4 input channels stored with
// channel stride 4, 64 output channels, output width 128 or 64).
int crn_wgrad_stem_try(const crn_conv_desc* d, const float* x, const float* dy, float* dw, cudaStream_t st) {
  if (d->transposed || d->kD != 1 || d->kH != 7 || d->kW != 7 || d->stride != 2 || d->pad != 3) return CRN_ERR_UNSUPPORTED;
  if (d->iD != 1 || d->oD != 1 || d->Cin > 4 || d->x_cs != 4 || d->x_co != 0 || d->Cout != 64 || d->CinP != 4) return CRN_ERR_UNSUPPORTED;
  if (d->iW != 2 * d->oW || d->iH != 2 * d->oH || (d->oW != 128 && d->oW != 64) || d->y_cs % 4 || d->y_co % 4) return CRN_ERR_UNSUPPORTED;
  if (crn_get_flags() & 32) return CRN_ERR_UNSUPPORTED;
  const int rows = d->N * d->oH;
  const int grid = rows < 2 * kNumSMs ? rows : 2 * kNumSMs;
  if (d->oW == 128)
    wgrad_stem_kernel<128><<<grid, SW_THREADS, 0, st>>>(x, dy, dw, d->N, d->iH, d->oH, d->x_cs, d->y_cs, d->y_co, d->CinP, d->CoutP);
  else
    wgrad_stem_kernel<64><<<grid, SW_THREADS, 0, st>>>(x, dy, dw, d->N, d->iH, d->oH, d->x_cs, d->y_cs, d->y_co, d->CinP, d->CoutP);
  CRN_LAUNCH_CHECK("wgrad_stem");
  return CRN_OK;
}
