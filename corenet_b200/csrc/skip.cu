// Ray-traced skip connection: project every voxel centre through the camera,
// truncate to the nearest-lower pixel, gather the compressed 2-D feature map
// and write it straight into the channel slice of the 3-D concat buffer.
//
// Replaces model/ray_traced_skip_connection.py:91-142 (everything after the
// 1x1 compress conv) and the torch.cat at model/reconstruction_decoder.py:117.
// HBM-bound, write dominated: 4*C*g^3 bytes per scene and scale.
//
// Index arithmetic reproduces the reference's fp32 sequence exactly:
//   c   = float(idx) + offset                      (:99-100)
//   p_n = fma(m_n2,z, fma(m_n1,y, m_n0*x)) + m_n3  (einsum at transformations.py:133,
//                                                   = what torch's K=4 bmm does)
//   u   = (p_x / p_w) / 2 + 0.5                    (:109,112)
//   ix  = clamp(trunc(u * w) + 1, 0, w + 1)        (:121-122,131-132)
//   out = p_z >= 0 ? padded[iy][ix] : 0            (:135-142)
#include "common.cuh"

namespace {
constexpr int NT = 256;

struct Proj { int ix, iy; bool front; };

// float -> int64 like x86 cvttss2si (what torch's CPU .to(int64) does):
// NaN / out of range -> INT64_MIN ("integer indefinite").
__device__ __forceinline__ long long trunc_i64_x86(float v) {
  if (!(v > -9.2233720368547758e18f && v < 9.2233720368547758e18f)) return (long long)0x8000000000000000ULL;
  return (long long)v;
}

__device__ __forceinline__ Proj project(const float* __restrict__ m, const float* __restrict__ offs,
                                        int x, int y, int z, int w, int h) {
  const float cx = __fadd_rn((float)x, offs[0]);
  const float cy = __fadd_rn((float)y, offs[1]);
  const float cz = __fadd_rn((float)z, offs[2]);
  float p[4];
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const float t0 = __fmul_rn(m[n * 4 + 0], cx);
    const float t1 = __fmaf_rn(m[n * 4 + 1], cy, t0);
    const float t2 = __fmaf_rn(m[n * 4 + 2], cz, t1);
    p[n] = __fadd_rn(t2, m[n * 4 + 3]);
  }
  const float ux = __fadd_rn(__fmul_rn(__fdiv_rn(p[0], p[3]), 0.5f), 0.5f);
  const float uy = __fadd_rn(__fmul_rn(__fdiv_rn(p[1], p[3]), 0.5f), 0.5f);
  long long ix = trunc_i64_x86(__fmul_rn(ux, (float)w));
  long long iy = trunc_i64_x86(__fmul_rn(uy, (float)h));
  // +1 then clamp to [0, w+1]; written so that INT64 extremes cannot overflow
  ix = ix < -1 ? 0 : (ix > (long long)w ? (long long)w + 1 : ix + 1);
  iy = iy < -1 ? 0 : (iy > (long long)h ? (long long)h + 1 : iy + 1);
  Proj r;
  r.ix = (int)ix; r.iy = (int)iy;
  r.front = p[2] >= 0.f;
  return r;
}

__global__ void __launch_bounds__(NT) skip_fwd_kernel(const float* __restrict__ map, int N, int h, int w,
                                                      int C, int map_cs, const float* __restrict__ m,
                                                      const float* __restrict__ offs, int gD, int gH, int gW,
                                                      float* __restrict__ out, int out_cs, int out_co) {
  const int Q = C / 4;
  const int64_t V = (int64_t)gD * gH * gW;
  const int64_t total = (int64_t)N * V * Q;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int q = (int)(i % Q);
    const int64_t vox = i / Q;
    const int n = (int)(vox / V);
    int64_t v = vox - (int64_t)n * V;
    const int x = (int)(v % gW); v /= gW;
    const int y = (int)(v % gH); const int z = (int)(v / gH);
    const Proj pr = project(m + n * 16, offs + n * 3, x, y, z, w, h);
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pr.front && pr.ix >= 1 && pr.ix <= w && pr.iy >= 1 && pr.iy <= h) {
      const int64_t pix = ((int64_t)n * h + (pr.iy - 1)) * w + (pr.ix - 1);
      val = __ldg(reinterpret_cast<const float4*>(map + pix * map_cs) + q);
    }
    *reinterpret_cast<float4*>(out + vox * out_cs + out_co + q * 4) = val;
  }
}

__global__ void __launch_bounds__(NT) skip_bwd_kernel(const float* __restrict__ dout, int out_cs, int out_co,
                                                      int N, int h, int w, int C, int map_cs,
                                                      const float* __restrict__ m,
                                                      const float* __restrict__ offs, int gD, int gH, int gW,
                                                      float* __restrict__ dmap) {
  const int Q = C / 4;
  const int64_t V = (int64_t)gD * gH * gW;
  const int64_t total = (int64_t)N * V * Q;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int q = (int)(i % Q);
    const int64_t vox = i / Q;
    const int n = (int)(vox / V);
    int64_t v = vox - (int64_t)n * V;
    const int x = (int)(v % gW); v /= gW;
    const int y = (int)(v % gH); const int z = (int)(v / gH);
    const Proj pr = project(m + n * 16, offs + n * 3, x, y, z, w, h);
    if (pr.front && pr.ix >= 1 && pr.ix <= w && pr.iy >= 1 && pr.iy <= h) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(dout + vox * out_cs + out_co) + q);
      const int64_t pix = ((int64_t)n * h + (pr.iy - 1)) * w + (pr.ix - 1);
      float* d = dmap + pix * map_cs + q * 4;
      atomicAdd(reinterpret_cast<float4*>(d), g);   // red.global.add.v4.f32 on sm_90+
    }
  }
}

__global__ void __launch_bounds__(NT) skip_idx_kernel(int N, int h, int w, const float* __restrict__ m,
                                                      const float* __restrict__ offs, int gD, int gH, int gW,
                                                      int32_t* __restrict__ idx) {
  const int64_t V = (int64_t)gD * gH * gW;
  const int64_t total = (int64_t)N * V;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int n = (int)(i / V);
    int64_t v = i - (int64_t)n * V;
    const int x = (int)(v % gW); v /= gW;
    const int y = (int)(v % gH); const int z = (int)(v / gH);
    const Proj pr = project(m + n * 16, offs + n * 3, x, y, z, w, h);
    idx[i] = pr.front ? pr.iy * (w + 2) + pr.ix : -1;
  }
}

inline unsigned grid_for(int64_t total) {
  int64_t b = crn_ceil_div(total, NT);
  if (b > 16LL * kNumSMs) b = 16LL * kNumSMs;
  return (unsigned)(b < 1 ? 1 : b);
}
}  // namespace

extern "C" int crn_skip_sample_fwd(const float* map, int32_t N, int32_t h, int32_t w, int32_t C,
                                   int32_t map_cs, const float* m, const float* offs, int32_t gD,
                                   int32_t gH, int32_t gW, float* out, int32_t out_cs, int32_t out_co,
                                   void* stream) {
  CRN_REQUIRE(map && m && offs && out, "crn_skip_sample_fwd: null pointer");
  CRN_REQUIRE(C % 4 == 0 && map_cs % 4 == 0 && out_cs % 4 == 0 && out_co % 4 == 0,
              "crn_skip_sample_fwd: channel counts/strides must be multiples of 4");
  skip_fwd_kernel<<<grid_for((int64_t)N * gD * gH * gW * (C / 4)), NT, 0, crn_stream(stream)>>>(
      map, N, h, w, C, map_cs, m, offs, gD, gH, gW, out, out_cs, out_co);
  CRN_LAUNCH_CHECK("skip_fwd");
  return CRN_OK;
}

extern "C" int crn_skip_sample_bwd(const float* dout, int32_t out_cs, int32_t out_co, int32_t N, int32_t h,
                                   int32_t w, int32_t C, int32_t map_cs, const float* m, const float* offs,
                                   int32_t gD, int32_t gH, int32_t gW, float* dmap, void* stream) {
  CRN_REQUIRE(dout && m && offs && dmap, "crn_skip_sample_bwd: null pointer");
  CRN_REQUIRE(C % 4 == 0 && map_cs % 4 == 0 && out_cs % 4 == 0 && out_co % 4 == 0,
              "crn_skip_sample_bwd: channel counts/strides must be multiples of 4");
  skip_bwd_kernel<<<grid_for((int64_t)N * gD * gH * gW * (C / 4)), NT, 0, crn_stream(stream)>>>(
      dout, out_cs, out_co, N, h, w, C, map_cs, m, offs, gD, gH, gW, dmap);
  CRN_LAUNCH_CHECK("skip_bwd");
  return CRN_OK;
}

extern "C" int crn_skip_indices(int32_t N, int32_t h, int32_t w, const float* m, const float* offs,
                                int32_t gD, int32_t gH, int32_t gW, int32_t* idx, void* stream) {
  CRN_REQUIRE(m && offs && idx, "crn_skip_indices: null pointer");
  skip_idx_kernel<<<grid_for((int64_t)N * gD * gH * gW), NT, 0, crn_stream(stream)>>>(N, h, w, m, offs, gD,
                                                                                     gH, gW, idx);
  CRN_LAUNCH_CHECK("skip_idx");
  return CRN_OK;
}
