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
#include <cub/device/device_radix_sort.cuh>

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

// Forward: a warp takes 32 consecutive voxels.  Each lane projects ONE voxel (two IEEE divides per voxel instead of
// per float4), then the warp copies the 32 x Q float4 of those voxels cooperatively: lane j handles float4 j, j+32, ...
// and gets its voxel's pixel by shuffle, so consecutive lanes read consecutive 16 B of a map row and write consecutive
// 16 B of a concat row (coalesced within a voxel's channel slice).
template <int QT>
__global__ void __launch_bounds__(NT) skip_fwd_kernel(const float* __restrict__ map, int N, int h, int w,
                                                      int C, int map_cs, const float* __restrict__ m,
                                                      const float* __restrict__ offs, int gD, int gH, int gW,
                                                      float* __restrict__ out, int out_cs, int out_co) {
  const int Q = QT > 0 ? QT : C / 4;
  const int64_t V = (int64_t)gD * gH * gW;
  const int64_t total = (int64_t)N * V;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * NT + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * NT) >> 5;
  for (int64_t base = warp0 * 32; base < total; base += nwarps * 32) {
    const int64_t vox = base + lane;
    int64_t pix = -1;
    if (vox < total) {
      const int n = (int)(vox / V);
      int64_t v = vox - (int64_t)n * V;
      const int x = (int)(v % gW); v /= gW;
      const int y = (int)(v % gH); const int z = (int)(v / gH);
      const Proj pr = project(m + n * 16, offs + n * 3, x, y, z, w, h);
      if (pr.front && pr.ix >= 1 && pr.ix <= w && pr.iy >= 1 && pr.iy <= h)
        pix = ((int64_t)n * h + (pr.iy - 1)) * w + (pr.ix - 1);
    }
    const int cnt = (int)((total - base) < 32 ? (total - base) : 32);
#pragma unroll 4
    for (int j0 = 0; j0 < cnt * Q; j0 += 32) {            // warp-uniform trip count (the shuffle needs all lanes)
      const int j = j0 + lane;
      const bool active = j < cnt * Q;
      const int v = active ? j / Q : 0, q = j - v * Q;
      const int64_t pv = __shfl_sync(0xffffffffu, pix, v);
      if (active) {
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pv >= 0) val = __ldg(reinterpret_cast<const float4*>(map + pv * map_cs) + q);
        *reinterpret_cast<float4*>(out + (base + v) * out_cs + out_co + q * 4) = val;
      }
    }
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

// ---- deterministic backward: voxels sorted by the pixel they sample (stable radix sort), one thread per
// (pixel, float4) sums its voxels in list order.  No atomics: bit-reproducible, and dmap is written, not accumulated.
__global__ void __launch_bounds__(NT) skip_keys_kernel(int N, int h, int w, const float* __restrict__ m,
                                                       const float* __restrict__ offs, int gD, int gH, int gW,
                                                       uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const int64_t V = (int64_t)gD * gH * gW;
  const int64_t total = (int64_t)N * V;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int n = (int)(i / V);
    int64_t v = i - (int64_t)n * V;
    const int x = (int)(v % gW); v /= gW;
    const int y = (int)(v % gH); const int z = (int)(v / gH);
    const Proj pr = project(m + n * 16, offs + n * 3, x, y, z, w, h);
    uint32_t key = (uint32_t)(N * h * w);          // sentinel: samples nothing (outside / behind the camera)
    if (pr.front && pr.ix >= 1 && pr.ix <= w && pr.iy >= 1 && pr.iy <= h)
      key = (uint32_t)((n * h + (pr.iy - 1)) * w + (pr.ix - 1));
    keys[i] = key;
    vals[i] = (int32_t)i;
  }
}

// starts[p] = first position in the sorted key list with key >= p, p in [0, P]
__global__ void __launch_bounds__(NT) skip_starts_kernel(const uint32_t* __restrict__ keys, int64_t total, int P,
                                                         int32_t* __restrict__ starts) {
  const int p = blockIdx.x * NT + threadIdx.x;
  if (p > P) return;
  int64_t lo = 0, hi = total;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < (uint32_t)p) lo = mid + 1; else hi = mid;
  }
  starts[p] = (int32_t)lo;
}

__global__ void __launch_bounds__(NT) skip_bwd_sorted_kernel(const float* __restrict__ dout, int out_cs, int out_co,
                                                             int P, int Q, int map_cs,
                                                             const int32_t* __restrict__ sorted_vox,
                                                             const int32_t* __restrict__ starts,
                                                             float* __restrict__ dmap) {
  const int64_t total = (int64_t)P * Q;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int p = (int)(i / Q), q = (int)(i - (int64_t)p * Q);
    const int s0 = __ldg(starts + p), s1 = __ldg(starts + p + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = s0;
    for (; k + 4 <= s1; k += 4) {                       // 4 independent gathers in flight, summed in list order
      float4 g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t vox = __ldg(sorted_vox + k + u);
        g[u] = __ldg(reinterpret_cast<const float4*>(dout + vox * out_cs + out_co) + q);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { acc.x += g[u].x; acc.y += g[u].y; acc.z += g[u].z; acc.w += g[u].w; }
    }
    for (; k < s1; ++k) {
      const int64_t vox = __ldg(sorted_vox + k);
      const float4 g = __ldg(reinterpret_cast<const float4*>(dout + vox * out_cs + out_co) + q);
      acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
    }
    *reinterpret_cast<float4*>(dmap + (int64_t)p * map_cs + q * 4) = acc;
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
  const unsigned grid = grid_for((int64_t)N * gD * gH * gW);
  cudaStream_t st = crn_stream(stream);
#define CRN_SKIP_FWD(QT) \
  skip_fwd_kernel<QT><<<grid, NT, 0, st>>>(map, N, h, w, C, map_cs, m, offs, gD, gH, gW, out, out_cs, out_co)
  switch (C / 4) {
    case 3: CRN_SKIP_FWD(3); break;
    case 6: CRN_SKIP_FWD(6); break;
    case 12: CRN_SKIP_FWD(12); break;
    case 24: CRN_SKIP_FWD(24); break;
    default: CRN_SKIP_FWD(0); break;
  }
#undef CRN_SKIP_FWD
  CRN_LAUNCH_CHECK("skip_fwd");
  return CRN_OK;
}

namespace {
inline int skip_key_bits(int64_t P) {
  int bits = 1;
  while (((int64_t)1 << bits) <= P) ++bits;      // keys take values 0..P (P = sentinel)
  return bits;
}
inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
}  // namespace

extern "C" int64_t crn_skip_lists_workspace_bytes(int32_t N, int32_t h, int32_t w, int32_t gD, int32_t gH,
                                                  int32_t gW) {
  const int64_t n = (int64_t)N * gD * gH * gW;
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, 0,
                                  skip_key_bits((int64_t)N * h * w));
  return (int64_t)(3 * align256(n * 4) + align256(temp));
}

extern "C" int crn_skip_build_lists(int32_t N, int32_t h, int32_t w, const float* m, const float* offs, int32_t gD,
                                    int32_t gH, int32_t gW, void* workspace, int64_t workspace_bytes,
                                    int32_t* sorted_vox, int32_t* starts, void* stream) {
  CRN_REQUIRE(m && offs && workspace && sorted_vox && starts, "crn_skip_build_lists: null pointer");
  const int64_t n = (int64_t)N * gD * gH * gW;
  const int64_t P = (int64_t)N * h * w;
  CRN_REQUIRE(n > 0 && n < ((int64_t)1 << 31) && P < ((int64_t)1 << 31), "crn_skip_build_lists: sizes out of range");
  CRN_REQUIRE(workspace_bytes >= crn_skip_lists_workspace_bytes(N, h, w, gD, gH, gW),
              "crn_skip_build_lists: workspace too small");
  cudaStream_t st = crn_stream(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  uint32_t* keys_in = reinterpret_cast<uint32_t*>(ws);
  uint32_t* keys_out = reinterpret_cast<uint32_t*>(ws + align256(n * 4));
  int32_t* vals_in = reinterpret_cast<int32_t*>(ws + 2 * align256(n * 4));
  void* temp = ws + 3 * align256(n * 4);
  size_t temp_bytes = (size_t)workspace_bytes - 3 * align256(n * 4);
  skip_keys_kernel<<<grid_for(n), NT, 0, st>>>(N, h, w, m, offs, gD, gH, gW, keys_in, vals_in);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, sorted_vox, (int)n, 0,
                                                  skip_key_bits(P), st);
  if (e != cudaSuccess) {
    crn_set_error("crn_skip_build_lists: radix sort failed: %s", cudaGetErrorString(e));
    return CRN_ERR_LAUNCH;
  }
  skip_starts_kernel<<<(unsigned)crn_ceil_div(P + 1, NT), NT, 0, st>>>(keys_out, n, (int)P, starts);
  crn_count_launches(2);
  CRN_LAUNCH_CHECK("skip_build_lists");
  return CRN_OK;
}

extern "C" int crn_skip_sample_bwd_sorted(const float* dout, int32_t out_cs, int32_t out_co, int32_t N, int32_t h,
                                          int32_t w, int32_t C, int32_t map_cs, const int32_t* sorted_vox,
                                          const int32_t* starts, float* dmap, void* stream) {
  CRN_REQUIRE(dout && sorted_vox && starts && dmap, "crn_skip_sample_bwd_sorted: null pointer");
  CRN_REQUIRE(C % 4 == 0 && map_cs % 4 == 0 && out_cs % 4 == 0 && out_co % 4 == 0,
              "crn_skip_sample_bwd_sorted: channel counts/strides must be multiples of 4");
  const int P = N * h * w, Q = C / 4;
  skip_bwd_sorted_kernel<<<grid_for((int64_t)P * Q), NT, 0, crn_stream(stream)>>>(dout, out_cs, out_co, P, Q, map_cs,
                                                                                 sorted_vox, starts, dmap);
  CRN_LAUNCH_CHECK("skip_bwd_sorted");
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
