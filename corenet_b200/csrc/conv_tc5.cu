// Conv3d k=5, stride 1, pad 2 (forward and dgrad) on the 5th-gen tensor cores: tcgen05.mma with the
// accumulators in TMEM, 3xTF32 operand splitting for fp32-class accuracy.
// Replaces the cuDNN calls behind nn.Conv3d(k=5) at model/reconstruction_decoder.py:66,74,82,91.
//
// Implicit GEMM without im2col copies:
//   * one CTA owns an (8 x 16) xy tile (M = 128 voxels) and ZT = 8 output z-planes: 8 accumulators of
//     128 x N fp32 live in TMEM (double buffered across work items);
//   * each input z-plane of the tile (+2 halo) is staged ONCE per K-pass (8 channels) in shared memory as
//     [hi|lo][k-chunk][y 20][x 12][4 ch]: that is the canonical no-swizzle K-major UMMA layout in which
//     the 8 x-neighbours of a row group are one 128 B core matrix, so a filter tap (ky, kx) is just a
//     +(ky*12+kx)*16 B shift of the descriptor start address -- every tap reads the same bytes;
//   * loop order kz -> ky -> z-plane: for a fixed kz the 8 output planes need 8 consecutive input planes,
//     so the plane ring slides by one per kz (9 slots) and a weight row (5 taps) is reused for 8 planes;
//   * weights are pre-split/pre-packed per (pass, tap) in the UMMA layout and streamed by cp.async.bulk
//     (TMA bulk copy) into a 3-stage ring with mbarrier transaction counts;
//   * warp roles: 4 epilogue warps (TMEM -> registers -> +bias -> global), 4 producer warps (halo gather,
//     hi/lo split, st.shared, fence.proxy.async), 1 MMA warp (single elected thread issues tcgen05.mma and
//     tcgen05.commit), 1 weight-copy warp.  All waits are bounded (a protocol bug cannot hang the GPU).
//
// a*b ~= hi_a*hi_b + lo_a*hi_b + hi_a*lo_b with hi = rna_tf32(a), lo = a - hi: relative error ~4e-7
// (measured by crn_tc_probe), i.e. the parity budget of the fp32 path, at 3 MMAs per product.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TX = 8, TY = 16;                   // tile: 8 x 16 voxels per plane (ZT planes, template)
constexpr int XS = TX + 4, YS = TY + 4;          // staged plane with halo
constexpr int CHUNK_BYTES = YS * XS * 16;        // one 4-channel chunk of a plane  (3840)
constexpr int PART_BYTES = 2 * CHUNK_BYTES;      // 8 channels                       (7680)
constexpr int PLANE_BYTES = 2 * PART_BYTES;      // hi + lo                          (15360)
constexpr int MAXSLOT = 9;                       // plane ring capacity (ZT + 1 slots used)
constexpr int WSTAGES = 3;
constexpr int NTHREADS = 320;                    // 4 epilogue + 4 producer + MMA + weight warps

struct TC5Params {
  const float* in;      // [N, D, H, W, in_cs] (+ in_co)
  const float* wtc;     // packed weights [P][125] x ([2][hi|lo][NPAD][4] (ND) or [hi|lo][2][NPAD][4])
  const float* bias;    // [gN] or null
  float* out;           // [N, D, H, W, out_cs] (+ out_co)
  int* status;          // set to 1 on a barrier timeout
  int N, D, H, W;
  int gK, gN;           // logical channels in / out
  int in_cs, in_co, out_cs, out_co;
  int P;                // K passes of 8 channels
  int cout_cls;         // scatter mode (ConvTranspose3d k7 s2): output channels per parity class (gN = 8 * cout_cls)
  int planar;           // scatter mode: out is [N, Cout, 2D, 2H, 2W] instead of channels-last
  int tiles_x, tiles_y, tiles_z;
  int nitems;
  int single;           // 1: single-pass TF32 (the A_lo / W_lo products are not issued)
};

struct __align__(8) Barriers {
  uint64_t plane_full[MAXSLOT], plane_empty[MAXSLOT];
  uint64_t w_full[WSTAGES], w_empty[WSTAGES];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  int abort_flag;
};

template <int ZT>
__device__ __forceinline__ void decode_item(const TC5Params& p, int item, int& n, int& z0, int& y0, int& x0) {
  int t = item;
  x0 = (t % p.tiles_x) * TX; t /= p.tiles_x;
  y0 = (t % p.tiles_y) * TY; t /= p.tiles_y;
  z0 = (t % p.tiles_z) * ZT; n = t / p.tiles_z;
}

// ND ("N doubling"): the hi*hi and hi*lo products share their A operand, so they are issued as ONE MMA
// against the concatenated B = [hi rows | lo rows] (N = 2*NPAD) into 2*NPAD accumulator columns that the
// epilogue adds; with the narrow N of these layers the MMA cost is set by the A fetch, so 2 MMAs per
// (tap, plane) instead of 3.
//
// KT = taps per axis: 5 (Conv3d k=5, offsets -2..2) or 4 (ConvTranspose3d k=7 s=2 p=3 seen from the INPUT grid:
// output voxel 2i+c gathers inputs i-1..i+2 with tap k = c+3-2o, so all 8 parity classes are one stride-1
// 4x4x4-tap convolution with N = 8*Cout "class channels" -- the A tiles are shared by every class).
// SCAT: epilogue scatters column (class, co) of voxel i to output voxel 2i+class.
// GATH: the dgrad of that layer -- the same 4^3-tap convolution run backwards: K = 8*Cout class channels gathered
// from dY (class channel (c, co) of coarse voxel i is dY[2i+c][co]), taps at offsets -2..1, N = Cin.
template <int NPAD, int ZT, bool ND, int KT, bool SCAT, bool GATH>
__global__ void __launch_bounds__(NTHREADS, 1) conv_tc5_kernel(const TC5Params p) {
  constexpr int HLO = (KT == 5 || GATH) ? 2 : 1;           // most negative tap offset
  constexpr int KZG = (KT == 4) ? 2 : 1;                   // kz planes per flush group (chain <= 96 MMAs)
  // epilogue keeps the running sums of all ZT planes in registers (16 columns: ZT * 16 floats per thread) -> no
  // global read-modify-write per flush group, one store per item (measured neutral on the N = 16 forward layer,
  // 2.25 -> 2.22 ms: that layer is bound by the MMA issue of its ~78-cycle 128x32x8 instructions, not the epilogue).
  constexpr bool RACC = NPAD == 16 && (SCAT || ZT <= 8);
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NSLOT = ZT + 1;                            // plane ring
  constexpr int NPLANE = ZT + KT - 1;                      // planes loaded per pass
  constexpr int WTAP_BYTES = 2 * 2 * NPAD * 16;            // one tap: hi|lo x 2 k-chunks x NPAD x 16 B
  constexpr int WROW_BYTES = KT * WTAP_BYTES;              // one (kz, ky) row of KT taps
  constexpr int ACOLS = ND ? 2 * NPAD : NPAD;              // accumulator columns per output plane
  constexpr int ASTG = (2 * ZT * ACOLS <= 512) ? 2 : 1;    // accumulator stages (epilogue overlap when 2)
  constexpr int TMEM_COLS = (ASTG * ZT * ACOLS <= 256) ? 256 : 512;
  static_assert(ASTG * ZT * ACOLS <= 512, "TMEM budget");
  uint8_t* ring = smem;
  uint8_t* wring = smem + NSLOT * PLANE_BYTES;
  Barriers* B = reinterpret_cast<Barriers*>(wring + WSTAGES * WROW_BYTES);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)tc::uniform_u32((uint32_t)tid >> 5);   // warp-uniform for the compiler (role branches)
  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) { tc::mbar_init(&B->plane_full[i], 128); tc::mbar_init(&B->plane_empty[i], 1); }
    for (int i = 0; i < WSTAGES; ++i) { tc::mbar_init(&B->w_full[i], 1); tc::mbar_init(&B->w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&B->acc_full[i], 1); tc::mbar_init(&B->acc_empty[i], 128); }
    B->abort_flag = 0;
    tc::mbar_fence_init();
  }
  if (warp == 8) tc::tmem_alloc(&B->tmem_base, TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tc::uniform_u32(B->tmem_base);
  const uint32_t ring_u32 = tc::smem_u32(ring), wring_u32 = tc::smem_u32(wring);

  auto fail = [&]() { B->abort_flag = 1; *p.status = 1; };
  volatile int* ab = &B->abort_flag;

  if (warp < 4) {
    // ============================ EPILOGUE: one flush per (pass, kz) group.
    // The tensor core's fp32 accumulate truncates, so a 375-MMA chain drifts by ~3e-5; flushing every
    // 75 MMAs and summing the groups in fp32 registers / global (round-to-nearest) keeps the result at
    // fp32-FFMA accuracy.  out (+)= acc with the same thread owning the same addresses across groups.
    long long G = 0;
    bool dead = false;
    for (int item = blockIdx.x; item < p.nitems && !dead; item += gridDim.x) {
      int n, z0, y0, x0;
      decode_item<ZT>(p, item, n, z0, y0, x0);
      const int m = warp * 32 + lane;              // row of the tile = TMEM lane
      const int y = y0 + (m >> 3), x = x0 + (m & 7);
      // scatter one 16-column chunk of plane zz (SCAT): all loads are issued before the stores (the compiler
      // cannot reorder them itself: the scattered addresses may alias), one global round trip per chunk
      auto scatter_chunk = [&](int zz, int c0, const float (&v)[16], bool first) {
        const int OH = 2 * p.H, OW = 2 * p.W;
        const long long oS = (long long)(2 * p.D) * OH * OW;
        if (!p.planar && (p.cout_cls & 3) == 0) {             // a column quad stays inside one class
          float* dst[4];
          float4 old[4];
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            const int cc = c0 + qd * 4;
            dst[qd] = nullptr;
            if (cc >= p.gN) continue;
            const int cls = cc / p.cout_cls, co = cc - cls * p.cout_cls;
            const int oz = 2 * (z0 + zz) + (cls >> 2), oy = 2 * y + ((cls >> 1) & 1), ox = 2 * x + (cls & 1);
            const long long sp = ((long long)oz * OH + oy) * OW + ox;
            dst[qd] = p.out + ((long long)n * oS + sp) * p.out_cs + p.out_co + co;
            if (first) {
              old[qd] = p.bias ? make_float4(__ldg(p.bias + co), __ldg(p.bias + co + 1), __ldg(p.bias + co + 2),
                                             __ldg(p.bias + co + 3))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
              old[qd] = *reinterpret_cast<const float4*>(dst[qd]);
            }
          }
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            if (!dst[qd]) continue;
            float4 o = old[qd];
            o.x += v[qd * 4]; o.y += v[qd * 4 + 1]; o.z += v[qd * 4 + 2]; o.w += v[qd * 4 + 3];
            *reinterpret_cast<float4*>(dst[qd]) = o;
          }
        } else if (first && p.planar && p.cout_cls == 2 && p.gN == 16 && c0 == 0) {
          // FG_BG logits (2 planes): the px = 0 / px = 1 classes of a coarse voxel are neighbouring fine voxels of the
          // same plane row -> one 8-byte store per (pz, py, channel); consecutive lanes (coarse x) write consecutive
          // 8 bytes, so every 32-byte sector is written whole by one instruction
#pragma unroll
          for (int q = 0; q < 4; ++q) {                         // q = pz * 2 + py
            const int oz = 2 * (z0 + zz) + (q >> 1), oy = 2 * y + (q & 1);
            const long long sp = ((long long)oz * OH + oy) * OW + 2 * x;
#pragma unroll
            for (int co = 0; co < 2; ++co) {
              const float b = p.bias ? __ldg(p.bias + co) : 0.f;
              *reinterpret_cast<float2*>(p.out + ((long long)n * 2 + co) * oS + sp) =
                  make_float2(v[q * 4 + co] + b, v[q * 4 + 2 + co] + b);
            }
          }
        } else if (first) {                                     // nothing to read: plain stores, no batching
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int cc = c0 + e;
            if (cc >= p.gN) continue;
            const int cls = cc / p.cout_cls, co = cc - cls * p.cout_cls;
            const int oz = 2 * (z0 + zz) + (cls >> 2), oy = 2 * y + ((cls >> 1) & 1), ox = 2 * x + (cls & 1);
            const long long sp = ((long long)oz * OH + oy) * OW + ox;
            float* dst = p.planar ? p.out + ((long long)n * p.cout_cls + co) * oS + sp
                                  : p.out + ((long long)n * oS + sp) * p.out_cs + p.out_co + co;
            *dst = v[e] + (p.bias ? __ldg(p.bias + co) : 0.f);
          }
        } else {
          float* dst[16];
          float old[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int cc = c0 + e;
            dst[e] = nullptr;
            if (cc >= p.gN) continue;
            const int cls = cc / p.cout_cls, co = cc - cls * p.cout_cls;
            const int oz = 2 * (z0 + zz) + (cls >> 2), oy = 2 * y + ((cls >> 1) & 1), ox = 2 * x + (cls & 1);
            const long long sp = ((long long)oz * OH + oy) * OW + ox;
            dst[e] = p.planar ? p.out + ((long long)n * p.cout_cls + co) * oS + sp
                              : p.out + ((long long)n * oS + sp) * p.out_cs + p.out_co + co;
            old[e] = *dst[e];
          }
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (dst[e]) *dst[e] = old[e] + v[e];
        }
      };
      // RACC: with 16 columns the running sums of all ZT planes fit in registers -> no global read-modify-write,
      // one store per item
      float sum[RACC ? ZT : 1][16];
      if constexpr (RACC) {
#pragma unroll
        for (int zz = 0; zz < ZT; ++zz)
#pragma unroll
          for (int e = 0; e < 16; ++e) sum[zz][e] = 0.f;
      }
      for (int pass = 0; pass < p.P && !dead; ++pass) {
        for (int kg = 0; kg < KT; kg += KZG, ++G) {
          const int st = (int)(G % ASTG);
          if (!tc::mbar_wait(&B->acc_full[st], (uint32_t)(G / ASTG) & 1, ab)) { fail(); dead = true; break; }
          tc::fence_after_sync();
          const bool first = pass == 0 && kg == 0;
          const bool last = pass == p.P - 1 && kg + KZG >= KT;
#pragma unroll
          for (int zz = 0; zz < ZT; ++zz) {
            bool valid = false;                      // did any kz of this group write accumulator zz?
#pragma unroll
            for (int kz = kg; kz < kg + KZG; ++kz) {
              const int q = z0 + zz + kz - HLO;
              valid = valid || (kz < KT && q >= 0 && q < p.D);
            }
            if constexpr (RACC) {
              if (valid) {
                const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + st * (ZT * ACOLS) + zz * ACOLS;
                float v[16];
                tc::tmem_ld16(ta, v);
#pragma unroll
                for (int e = 0; e < 16; ++e) sum[zz][e] += v[e];
                if constexpr (ND) {
                  tc::tmem_ld16(ta + NPAD, v);
#pragma unroll
                  for (int e = 0; e < 16; ++e) sum[zz][e] += v[e];
                }
              }
              if (last) {
                if constexpr (SCAT) {
                  scatter_chunk(zz, 0, sum[zz], true);
                } else {
                  const long long pos = (((long long)n * p.D + (z0 + zz)) * p.H + y) * p.W + x;
                  float* dst = p.out + pos * p.out_cs + p.out_co;
#pragma unroll
                  for (int qd = 0; qd < 4; ++qd) {
                    const int c = qd * 4;
                    if (c < p.gN) {                        // gN is a multiple of 4
                      float4 o = make_float4(sum[zz][c], sum[zz][c + 1], sum[zz][c + 2], sum[zz][c + 3]);
                      if (p.bias) {
                        o.x += __ldg(p.bias + c); o.y += __ldg(p.bias + c + 1);
                        o.z += __ldg(p.bias + c + 2); o.w += __ldg(p.bias + c + 3);
                      }
                      *reinterpret_cast<float4*>(dst + c) = o;
                    }
                  }
                }
#pragma unroll
                for (int e = 0; e < 16; ++e) sum[zz][e] = 0.f;
              }
              continue;
            }
            if (!valid && !first) continue;
            if constexpr (!SCAT) {
            const long long pos = (((long long)n * p.D + (z0 + zz)) * p.H + y) * p.W + x;
            float* dst = p.out + pos * p.out_cs + p.out_co;
#pragma unroll
            for (int c0 = 0; c0 < NPAD; c0 += 16) {
              float v[16];
              if (valid) {
                const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + st * (ZT * ACOLS) + zz * ACOLS + c0;
                tc::tmem_ld16(ta, v);
                if constexpr (ND) {
                  float v2[16];
                  tc::tmem_ld16(ta + NPAD, v2);
#pragma unroll
                  for (int e = 0; e < 16; ++e) v[e] += v2[e];
                }
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = 0.f;
              }
#pragma unroll
              for (int qd = 0; qd < 4; ++qd) {
                const int c = c0 + qd * 4;
                if (c < p.gN) {                          // gN is a multiple of 4
                  float4 o = make_float4(v[qd * 4], v[qd * 4 + 1], v[qd * 4 + 2], v[qd * 4 + 3]);
                  if (first) {
                    if (p.bias) {
                      o.x += __ldg(p.bias + c); o.y += __ldg(p.bias + c + 1);
                      o.z += __ldg(p.bias + c + 2); o.w += __ldg(p.bias + c + 3);
                    }
                  } else {
                    const float4 old = *reinterpret_cast<const float4*>(dst + c);
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                  }
                  *reinterpret_cast<float4*>(dst + c) = o;
                }
              }
            }
            } else {
            // scatter: column cc = class * cout_cls + co of voxel (z, y, x) -> output voxel (2z+cz, 2y+cy, 2x+cx)
#pragma unroll 1
            for (int c0 = 0; c0 < NPAD; c0 += 16) {
              if (c0 >= p.gN) break;
              float v[16];
              if (valid) {
                const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + st * (ZT * ACOLS) + zz * ACOLS + c0;
                tc::tmem_ld16(ta, v);
                if constexpr (ND) {
                  float v2[16];
                  tc::tmem_ld16(ta + NPAD, v2);
#pragma unroll
                  for (int e = 0; e < 16; ++e) v[e] += v2[e];
                }
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = 0.f;
              }
              scatter_chunk(zz, c0, v, first);
            }
            }
          }
          tc::fence_before_sync();
          tc::mbar_arrive(&B->acc_empty[st]);
        }
      }
    }
  } else if (warp < 8) {
    // ============================ PRODUCERS: halo gather + hi/lo split into the plane ring
    const int pt = tid - 128;                      // 0..127
    long long L = 0;                               // running plane-load index (slot = L % NSLOT)
    bool dead = false;
    for (int item = blockIdx.x; item < p.nitems && !dead; item += gridDim.x) {
      int n, z0, y0, x0;
      decode_item<ZT>(p, item, n, z0, y0, x0);
      for (int pass = 0; pass < p.P && !dead; ++pass) {
        for (int r = 0; r < NPLANE; ++r, ++L) {
          const int slot = (int)(L % NSLOT);
          const uint32_t use = (uint32_t)(L / NSLOT);
          if (use > 0 && !tc::mbar_wait(&B->plane_empty[slot], (use - 1) & 1, ab)) { fail(); dead = true; break; }
          const int q = z0 - HLO + r;
          if (q >= 0 && q < p.D) {
            uint8_t* dst = ring + slot * PLANE_BYTES;
            // all gathers of this thread are issued before the first split / store: NU independent loads in flight
            constexpr int NU = (YS * XS * 2 + 127) / 128;
            float4 a[NU];
#pragma unroll
            for (int i = 0; i < NU; ++i) {
              const int u = pt + i * 128;
              a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (u < YS * XS * 2) {
                const int kc = u & 1; const int v = u >> 1;
                const int xs = v % XS, ys = v / XS;
                const int y = y0 - HLO + ys, x = x0 - HLO + xs;
                const int k = pass * 8 + kc * 4;
                if ((unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W && k < p.gK) {
                  long long off;
                  if constexpr (GATH) {              // class channel k = (cls, co) lives at fine voxel 2i + cls
                    const int cls = k / p.cout_cls, co = k - cls * p.cout_cls;
                    off = ((((long long)n * (2 * p.D) + 2 * q + (cls >> 2)) * (2 * p.H) + 2 * y + ((cls >> 1) & 1)) *
                               (2 * p.W) + 2 * x + (cls & 1)) * p.in_cs + p.in_co + co;
                  } else {
                    off = ((((long long)n * p.D + q) * p.H + y) * p.W + x) * p.in_cs + p.in_co + k;
                  }
                  a[i] = __ldg(reinterpret_cast<const float4*>(p.in + off));
                }
              }
            }
#pragma unroll
            for (int i = 0; i < NU; ++i) {
              const int u = pt + i * 128;
              if (u < YS * XS * 2) {
                const int kc = u & 1; const int v = u >> 1;
                float4 hi, lo;
                tc::split_tf32(a[i].x, hi.x, lo.x); tc::split_tf32(a[i].y, hi.y, lo.y);
                tc::split_tf32(a[i].z, hi.z, lo.z); tc::split_tf32(a[i].w, hi.w, lo.w);
                const int o = kc * CHUNK_BYTES + v * 16;
                *reinterpret_cast<float4*>(dst + o) = hi;
                *reinterpret_cast<float4*>(dst + PART_BYTES + o) = lo;
              }
            }
            tc::fence_async_smem();
          }
          tc::mbar_arrive(&B->plane_full[slot]);
        }
      }
    }
  } else if (warp == 8) {
    // ============================ MMA ISSUER: the whole warp runs the loop converged, one elected lane issues
    // (tc_common.cuh elect_one: an `if (lane == 0)` region turns every UTCHMMA into a vote loop)
    {
      const int p_single = p.single;
      constexpr uint32_t idesc = tc::make_idesc_tf32(128, NPAD, 0, 0);
      constexpr uint32_t idesc2 = tc::make_idesc_tf32(128, 2 * NPAD, 0, 0);
      constexpr uint32_t A_DESC_HI = (uint32_t)((XS * 16) >> 4) | (1u << 14);   // SBO field | version 1 (bit 46)
      long long L0 = 0;                             // plane-load index of r = 0 of the current pass
      long long Wn = 0;                             // running weight-row index
      long long G = 0;                              // running (pass, kz) group index -> accumulator stage
      bool dead = false;
      for (int item = blockIdx.x; item < p.nitems && !dead; item += gridDim.x) {
        int n, z0, y0, x0;
        decode_item<ZT>(p, item, n, z0, y0, x0);
        for (int pass = 0; pass < p.P && !dead; ++pass, L0 += NPLANE) {
          int st = 0;
          uint32_t started = 0;                     // bit zz: accumulator zz already written in this group
          for (int kz = 0; kz < KT && !dead; ++kz) {
            if (kz % KZG == 0) {                    // a new flush group starts
              st = (int)(G % ASTG);
              if (G >= ASTG && !tc::mbar_wait(&B->acc_empty[st], (uint32_t)((G / ASTG) - 1) & 1, ab)) { fail(); dead = true; break; }
              tc::fence_after_sync();
              started = 0;
            }
            // planes r = kz .. kz+7 must be resident: r <= 7 are awaited at kz = 0, then one new per kz
            const int rlo = kz == 0 ? 0 : kz + ZT - 1, rhi = kz + ZT - 1;
            for (int r = rlo; r <= rhi; ++r) {
              const long long L = L0 + r;
              if (!tc::mbar_wait(&B->plane_full[L % NSLOT], (uint32_t)(L / NSLOT) & 1, ab)) { fail(); dead = true; break; }
            }
            if (dead) break;
            tc::fence_after_sync();
            uint32_t abase[ZT];
            uint32_t valid = 0;
#pragma unroll
            for (int zz = 0; zz < ZT; ++zz) {
              const int q = z0 + zz + kz - HLO;
              if (q >= 0 && q < p.D) valid |= 1u << zz;     // zero planes contribute nothing
              // low descriptor word of the plane: (addr >> 4) | LBO field; taps/parts add a constant to it
              abase[zz] = ((ring_u32 + (uint32_t)((L0 + zz + kz) % NSLOT) * PLANE_BYTES) >> 4) |
                          ((uint32_t)(CHUNK_BYTES >> 4) << 16);
            }
            for (int ky = 0; ky < KT; ++ky, ++Wn) {
              const int ws = (int)(Wn % WSTAGES);
              if (!tc::mbar_wait(&B->w_full[ws], (uint32_t)(Wn / WSTAGES) & 1, ab)) { fail(); dead = true; break; }
              tc::fence_after_sync();
              const uint32_t wbase = wring_u32 + ws * WROW_BYTES;
              // low descriptor words of the weight row (smem addresses >> 4 stay below 2^14: adding tap offsets never
              // carries into the LBO field)
              constexpr uint32_t B_DESC_HI = (uint32_t)(128 >> 4) | (1u << 14);       // SBO | version 1
              const uint32_t wlo_nd = ((wbase >> 4) & 0x3FFFu) | ((uint32_t)((2 * NPAD * 16) >> 4) << 16);
              const uint32_t wlo_w = ((wbase >> 4) & 0x3FFFu) | ((uint32_t)((NPAD * 16) >> 4) << 16);
              // Interleave the 8 output planes: consecutive tcgen05.mma go to DIFFERENT accumulators, so the
              // tensor pipe never waits on a read-after-write of the same TMEM tile.
#pragma unroll
              for (int kx = 0; kx < KT; ++kx) {
                if constexpr (ND) {
                  // tap layout [kc][hi rows | lo rows][16 B]: one descriptor, N = 2*NPAD or the first NPAD rows
                  const uint64_t db = ((uint64_t)B_DESC_HI << 32) |
                                      (wlo_nd + (uint32_t)((kx * WTAP_BYTES) >> 4));
#pragma unroll
                  for (int part = 0; part < 2; ++part) {
                    if (part && p_single) break;
#pragma unroll
                    for (int zz = 0; zz < ZT; ++zz) {
                      if (!((valid >> zz) & 1u)) continue;
                      const uint32_t alo = abase[zz] + (ky * XS + kx) + (part == 1 ? (PART_BYTES >> 4) : 0);
                      const uint64_t da = ((uint64_t)A_DESC_HI << 32) | alo;
                      const uint32_t acc = (kx | part) ? 1u : ((started >> zz) & 1u);
                      tc::mma_tf32_e(tmem + st * (ZT * ACOLS) + zz * ACOLS, da, db, part == 0 ? idesc2 : idesc, acc);
                    }
                  }
                } else {
                  const uint64_t dbh = ((uint64_t)B_DESC_HI << 32) | (wlo_w + (uint32_t)((kx * WTAP_BYTES) >> 4));
                  const uint64_t dbl = dbh + (uint64_t)((2 * NPAD * 16) >> 4);
#pragma unroll
                  for (int part = 0; part < 3; ++part) {
                    if (part && p_single) break;
#pragma unroll
                    for (int zz = 0; zz < ZT; ++zz) {
                      if (!((valid >> zz) & 1u)) continue;
                      const uint32_t alo = abase[zz] + (ky * XS + kx) + (part == 1 ? (PART_BYTES >> 4) : 0);
                      const uint64_t da = ((uint64_t)A_DESC_HI << 32) | alo;
                      const uint32_t acc = (kx | part) ? 1u : ((started >> zz) & 1u);
                      tc::mma_tf32_e(tmem + st * (ZT * ACOLS) + zz * ACOLS, da, part == 2 ? dbl : dbh, idesc, acc);
                    }
                  }
                }
              }
              started |= valid;
              tc::commit_e(&B->w_empty[ws]);          // weight row free once these MMAs have completed
            }
            if (dead) break;
            tc::commit_e(&B->plane_empty[(L0 + kz) % NSLOT]);      // plane r = kz is done
            if (kz % KZG == KZG - 1 || kz == KT - 1) {
              tc::commit_e(&B->acc_full[st]);                      // this group's partial sums are complete
              ++G;
            }
          }
          if (dead) break;
          for (int r = KT; r < NPLANE; ++r) tc::commit_e(&B->plane_empty[(L0 + r) % NSLOT]);
        }
      }
    }
  } else {
    // ============================ WEIGHT COPIES: one cp.async.bulk per (pass, kz, ky) row
    if (lane == 0) {
      long long Wn = 0;
      bool dead = false;
      for (int item = blockIdx.x; item < p.nitems && !dead; item += gridDim.x) {
        for (int pass = 0; pass < p.P && !dead; ++pass) {
          for (int row = 0; row < KT * KT; ++row, ++Wn) {
            const int ws = (int)(Wn % WSTAGES);
            const uint32_t use = (uint32_t)(Wn / WSTAGES);
            if (use > 0 && !tc::mbar_wait(&B->w_empty[ws], (use - 1) & 1, ab)) { fail(); dead = true; break; }
            const float* src = p.wtc + ((size_t)pass * (KT * KT) + row) * (WROW_BYTES / 4);
            const uint32_t bar = tc::smem_u32(&B->w_full[ws]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)WROW_BYTES)
                         : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    wring_u32 + ws * WROW_BYTES),
                "l"(src), "r"((uint32_t)WROW_BYTES), "r"(bar)
                : "memory");
          }
        }
      }
    }
  }
  // ---- teardown
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, TMEM_COLS);
}

// pack kernel: PyTorch conv weight [Cout][Cin][125] -> wtc[P][125][tap block] (see TC5Params::wtc)
//   fwd  : k = ci, n = co, tap t
//   dgrad: k = co, n = ci, tap 124 - t   (the gradient wrt x is a conv with the flipped kernel)
__global__ void tc5_pack_kernel(const float* __restrict__ w, int Cout, int Cin, int dgrad, int NPAD, int P,
                                int nd, float* __restrict__ out) {
  const long long total = (long long)P * 125 * 2 * NPAD * 4;     // (pass, tap, kc, n, e)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i & 3); long long r = i >> 2;
    const int n = (int)(r % NPAD); r /= NPAD;
    const int kc = (int)(r & 1); r >>= 1;
    const int t = (int)(r % 125); const int pass = (int)(r / 125);
    const int k = pass * 8 + kc * 4 + e;
    const int K = dgrad ? Cout : Cin, Nn = dgrad ? Cin : Cout;
    float v = 0.f;
    if (k < K && n < Nn) {
      const int co = dgrad ? k : n, ci = dgrad ? n : k;
      const int ts = dgrad ? 124 - t : t;
      v = w[((long long)co * Cin + ci) * 125 + ts];
    }
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    const float hi = __uint_as_float(h), lo = v - hi;
    const long long base = (((long long)pass * 125 + t) * 2) * 2 * NPAD * 4;      // start of this tap (hi part)
    if (nd) {            // [kc][hi rows | lo rows][4]
      const long long off = ((long long)kc * 2 * NPAD + n) * 4 + e;
      out[base + off] = hi;
      out[base + off + (long long)NPAD * 4] = lo;
    } else {             // [hi: [kc][NPAD][4]] [lo: [kc][NPAD][4]]
      const long long off = ((long long)kc * NPAD + n) * 4 + e;
      out[base + off] = hi;
      out[base + 2LL * NPAD * 4 + off] = lo;
    }
  }
}

// pack kernel: PyTorch ConvTranspose3d weight [Cin][Cout][7][7][7] -> wtc[P][64 taps][tap block], column
// n = class * Cout + co (class = cz*4 + cy*2 + cx), tap j = (jz*4 + jy)*4 + jx <-> input offset j - 1 per axis,
// filter index k = c + 3 - 2*(j - 1) (zero weight where k falls outside 0..6: class 0 has 3 taps per axis).
// dgrad != 0: the transposed operator, K = (class, co), N = ci, tap j <-> offset j - 2, k = c - 1 + 2*j.
__global__ void tct_pack_kernel(const float* __restrict__ w, int Cin, int Cout, int NPAD, int P, int nd, int dgrad,
                                float* __restrict__ out) {
  const long long total = (long long)P * 64 * 2 * NPAD * 4;      // (pass, tap, kc, n, e)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i & 3); long long r = i >> 2;
    const int n = (int)(r % NPAD); r /= NPAD;
    const int kc = (int)(r & 1); r >>= 1;
    const int t = (int)(r & 63); const int pass = (int)(r >> 6);
    const int kk = pass * 8 + kc * 4 + e;
    float v = 0.f;
    if (!dgrad) {
      const int ci = kk;
      if (ci < Cin && n < 8 * Cout) {
        const int cls = n / Cout, co = n - cls * Cout;
        const int kz = (cls >> 2) + 5 - 2 * (t >> 4), ky = ((cls >> 1) & 1) + 5 - 2 * ((t >> 2) & 3),
                  kx = (cls & 1) + 5 - 2 * (t & 3);
        if ((unsigned)kz < 7u && (unsigned)ky < 7u && (unsigned)kx < 7u)
          v = w[(((long long)ci * Cout + co) * 7 + kz) * 49 + ky * 7 + kx];
      }
    } else {
      const int ci = n;
      if (ci < Cin && kk < 8 * Cout) {
        const int cls = kk / Cout, co = kk - cls * Cout;
        const int kz = (cls >> 2) - 1 + 2 * (t >> 4), ky = ((cls >> 1) & 1) - 1 + 2 * ((t >> 2) & 3),
                  kx = (cls & 1) - 1 + 2 * (t & 3);
        if ((unsigned)kz < 7u && (unsigned)ky < 7u && (unsigned)kx < 7u)
          v = w[(((long long)ci * Cout + co) * 7 + kz) * 49 + ky * 7 + kx];
      }
    }
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    const float hi = __uint_as_float(h), lo = v - hi;
    const long long base = (((long long)pass * 64 + t) * 2) * 2 * NPAD * 4;
    if (nd) {
      const long long off = ((long long)kc * 2 * NPAD + n) * 4 + e;
      out[base + off] = hi;
      out[base + off + (long long)NPAD * 4] = lo;
    } else {
      const long long off = ((long long)kc * NPAD + n) * 4 + e;
      out[base + off] = hi;
      out[base + 2LL * NPAD * 4 + off] = lo;
    }
  }
}

template <int NPAD, int ZT, bool ND, int KT = 5, bool SCAT = false, bool GATH = false>
int launch_tc5(TC5Params p, cudaStream_t st) {
  constexpr int WROW_BYTES = KT * 2 * 2 * NPAD * 16;
  const size_t smem = (size_t)(ZT + 1) * PLANE_BYTES + (size_t)WSTAGES * WROW_BYTES + sizeof(Barriers) + 64;
  p.tiles_z = p.D / ZT;
  p.nitems = p.N * p.tiles_x * p.tiles_y * p.tiles_z;
  auto kern = conv_tc5_kernel<NPAD, ZT, ND, KT, SCAT, GATH>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      crn_set_error("conv_tc5: cannot set %zu bytes of dynamic shared memory", smem);
      return CRN_ERR_LAUNCH;
    }
    configured = true;
  }
  int grid = p.nitems < kNumSMs ? p.nitems : kNumSMs;
  kern<<<grid, NTHREADS, smem, st>>>(p);
  CRN_LAUNCH_CHECK("conv_tc5");
  return CRN_OK;
}

}  // namespace

// the packed-weight layout and the kernel variant must agree: both follow this switch (flag 64 = 3 narrow MMAs)
static bool tc5_use_nd(int N) {
  if (crn_get_flags() & 64) return false;
  return N <= 16;
}

extern "C" int64_t crn_tc5_packed_floats(int32_t K, int32_t N) {
  const int P = (K + 7) / 8;
  const int NPAD = N <= 16 ? 16 : (N <= 32 ? 32 : 64);
  return (int64_t)P * 125 * 2 * 2 * NPAD * 4;
}

extern "C" int crn_tc5_pack(const float* w, int32_t Cout, int32_t Cin, int32_t dgrad, float* out, void* stream) {
  CRN_REQUIRE(w && out && Cout > 0 && Cin > 0, "crn_tc5_pack: bad args");
  const int K = dgrad ? Cout : Cin, N = dgrad ? Cin : Cout;
  CRN_REQUIRE(N <= 64, "crn_tc5_pack: N > 64 unsupported");
  const int P = (K + 7) / 8;
  const int NPAD = N <= 16 ? 16 : (N <= 32 ? 32 : 64);
  const long long total = (long long)P * 125 * 2 * NPAD * 4;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  tc5_pack_kernel<<<blocks, 256, 0, crn_stream(stream)>>>(w, Cout, Cin, dgrad, NPAD, P, tc5_use_nd(N) ? 1 : 0, out);
  CRN_LAUNCH_CHECK("tc5_pack");
  return CRN_OK;
}

// kind 0: y = conv5(x) + bias (in = x, gK = Cin, gN = Cout); kind 1: dx = conv5^T(dy) (in = dy, gK = Cout, gN = Cin).
extern "C" int crn_conv5_tc(const crn_conv_desc* d, int32_t kind, const float* in, const float* wtc,
                            const float* bias, float* out, int32_t* status, void* stream) {
  CRN_REQUIRE(d && in && wtc && out && status, "crn_conv5_tc: null pointer");
  CRN_REQUIRE(!d->transposed && d->kD == 5 && d->kH == 5 && d->kW == 5 && d->stride == 1 && d->pad == 2,
              "crn_conv5_tc: only Conv3d k=5 s=1 p=2");
  CRN_REQUIRE(d->iD == d->oD && d->iH == d->oH && d->iW == d->oW, "crn_conv5_tc: shape mismatch");
  CRN_REQUIRE(d->iW % TX == 0 && d->iH % TY == 0 && d->iD % 8 == 0, "crn_conv5_tc: grid must tile by 8x16x8");
  CRN_REQUIRE(!d->y_planar && !d->bias_n_stride, "crn_conv5_tc: planar / per-scene bias unsupported");
  TC5Params p{};
  p.single = crn_single_pass();
  p.in = in; p.wtc = wtc; p.bias = kind == 0 ? bias : nullptr; p.out = out; p.status = status;
  p.N = d->N; p.D = d->iD; p.H = d->iH; p.W = d->iW;
  if (kind == 0) {
    p.gK = d->Cin; p.gN = d->Cout; p.in_cs = d->x_cs; p.in_co = d->x_co; p.out_cs = d->y_cs; p.out_co = d->y_co;
  } else {
    p.gK = d->Cout; p.gN = d->Cin; p.in_cs = d->y_cs; p.in_co = d->y_co; p.out_cs = d->x_cs; p.out_co = d->x_co;
  }
  CRN_REQUIRE(p.gN % 4 == 0 && p.gK % 4 == 0 && p.gN <= 64, "crn_conv5_tc: channels must be multiples of 4, N <= 64");
  CRN_REQUIRE(p.in_cs % 4 == 0 && p.in_co % 4 == 0 && p.out_cs % 4 == 0 && p.out_co % 4 == 0,
              "crn_conv5_tc: channel strides/offsets must be multiples of 4");
  p.P = (p.gK + 7) / 8;
  p.tiles_x = p.W / TX; p.tiles_y = p.H / TY;
  cudaStream_t st = crn_stream(stream);
  // 8 output planes per item while two accumulator stages fit in TMEM (2*ZT*NPAD <= 512 columns), else 4
  // N doubling pays where it keeps two accumulator stages (measured on a B200, scripts/tc5_test.py):
  //   N<=16: 3.09 -> 2.25 ms (stage-6 forward); N<=32 (ZT=8 or 4) and N<=64 do not gain.
  if (tc5_use_nd(p.gN)) {
    return launch_tc5<16, 8, true>(p, st);
  }
  if (p.gN <= 16) return launch_tc5<16, 8, false>(p, st);
  if (p.gN <= 32) return launch_tc5<32, 8, false>(p, st);
  // small grids (stage_4.c1 at 16^3, B = 4: 32 four-plane items): one output plane per item fills the machine
  if ((long long)p.N * p.tiles_x * p.tiles_y * (p.D / 4) <= kNumSMs / 2) return launch_tc5<64, 1, false>(p, st);
  return launch_tc5<64, 4, false>(p, st);
}

// ---------------------------------------------------------------------------------------------
// ConvTranspose3d k=7 s=2 p=3 op=1 forward on the same kernel (KT = 4, scatter epilogue).
// Replaces the cuDNN call behind nn.ConvTranspose3d at model/reconstruction_decoder.py:69,77,85,95.
static int tct_npad(int Cout) {
  const int n = 8 * Cout;
  return n <= 16 ? 16 : (n <= 32 ? 32 : (n <= 64 ? 64 : 128));
}

static int tc_npad64(int N) { return N <= 16 ? 16 : (N <= 32 ? 32 : 64); }

extern "C" int64_t crn_tct_packed_floats(int32_t Cin, int32_t Cout, int32_t dgrad) {
  const int P = ((dgrad ? 8 * Cout : Cin) + 7) / 8;
  const int NPAD = dgrad ? tc_npad64(Cin) : tct_npad(Cout);
  return (int64_t)P * 64 * 2 * 2 * NPAD * 4;
}

extern "C" int crn_tct_pack(const float* w, int32_t Cin, int32_t Cout, int32_t dgrad, float* out, void* stream) {
  CRN_REQUIRE(w && out && Cout > 0 && Cin > 0, "crn_tct_pack: bad args");
  if (dgrad) CRN_REQUIRE(Cin <= 64 && Cout % 4 == 0, "crn_tct_pack: dgrad needs Cin <= 64 and Cout % 4 == 0");
  else CRN_REQUIRE(8 * Cout <= 128, "crn_tct_pack: Cout > 16 unsupported");
  const int K = dgrad ? 8 * Cout : Cin, N = dgrad ? Cin : 8 * Cout;
  const int P = (K + 7) / 8;
  const int NPAD = dgrad ? tc_npad64(Cin) : tct_npad(Cout);
  const long long total = (long long)P * 64 * 2 * NPAD * 4;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  tct_pack_kernel<<<blocks, 256, 0, crn_stream(stream)>>>(w, Cin, Cout, NPAD, P, tc5_use_nd(N) ? 1 : 0, dgrad ? 1 : 0,
                                                          out);
  CRN_LAUNCH_CHECK("tct_pack");
  return CRN_OK;
}

extern "C" int crn_convt7_tc(const crn_conv_desc* d, const float* x, const float* wtc, const float* bias, float* y,
                             int32_t* status, void* stream) {
  CRN_REQUIRE(d && x && wtc && y && status, "crn_convt7_tc: null pointer");
  CRN_REQUIRE(d->transposed && d->kD == 7 && d->kH == 7 && d->kW == 7 && d->stride == 2 && d->pad == 3,
              "crn_convt7_tc: only ConvTranspose3d k=7 s=2 p=3");
  CRN_REQUIRE(d->oD == 2 * d->iD && d->oH == 2 * d->iH && d->oW == 2 * d->iW, "crn_convt7_tc: output must be 2x input");
  CRN_REQUIRE(d->iW % TX == 0 && d->iH % TY == 0 && d->iD % 8 == 0, "crn_convt7_tc: input grid must tile by 8x16x8");
  CRN_REQUIRE(!d->bias_n_stride, "crn_convt7_tc: per-scene bias unsupported");
  CRN_REQUIRE(d->Cin % 4 == 0 && d->x_cs % 4 == 0 && d->x_co % 4 == 0 && 8 * d->Cout <= 128,
              "crn_convt7_tc: Cin, x strides must be multiples of 4 and Cout <= 16");
  TC5Params p{};
  p.single = crn_single_pass();
  p.in = x; p.wtc = wtc; p.bias = bias; p.out = y; p.status = status;
  p.N = d->N; p.D = d->iD; p.H = d->iH; p.W = d->iW;
  p.gK = d->Cin; p.gN = 8 * d->Cout; p.cout_cls = d->Cout; p.planar = d->y_planar;
  p.in_cs = d->x_cs; p.in_co = d->x_co; p.out_cs = d->y_cs; p.out_co = d->y_co;
  if (!d->y_planar && d->Cout % 4 == 0)
    CRN_REQUIRE(d->y_cs % 4 == 0 && d->y_co % 4 == 0, "crn_convt7_tc: y strides must be multiples of 4");
  p.P = (p.gK + 7) / 8;
  p.tiles_x = p.W / TX; p.tiles_y = p.H / TY;
  cudaStream_t st = crn_stream(stream);
  if (p.gN <= 16) {
    // ZT = 4: the epilogue keeps 4 x 16 running sums per thread in registers (8 planes would spill)
    if (tc5_use_nd(p.gN)) return launch_tc5<16, 4, true, 4, true>(p, st);
    return launch_tc5<16, 4, false, 4, true>(p, st);
  }
  if (p.gN <= 32) return launch_tc5<32, 8, false, 4, true>(p, st);
  if (p.gN <= 64) return launch_tc5<64, 4, false, 4, true>(p, st);
  // two planes per item keep two accumulator stages in TMEM (2*2*128 columns): 0.55 ms vs 0.74 ms with ZT = 4
  // on the stage-5 layer (scripts/tct_test.py)
  // small grids (stage_4.t1 at 16^3, B = 4: 64 two-plane items on 148 SMs): one plane per item fills the machine
  if ((long long)p.N * p.tiles_x * p.tiles_y * (p.D / 2) <= kNumSMs / 2) return launch_tc5<128, 1, false, 4, true>(p, st);
  return launch_tc5<128, 2, false, 4, true>(p, st);
}

// dx = ConvTranspose3d^T(dy): the dgrad of the layer above (dy channels-last [N, 2D, 2H, 2W, y_cs]).
extern "C" int crn_convt7_tc_dgrad(const crn_conv_desc* d, const float* dy, const float* wtc, float* dx,
                                   int32_t* status, void* stream) {
  CRN_REQUIRE(d && dy && wtc && dx && status, "crn_convt7_tc_dgrad: null pointer");
  CRN_REQUIRE(d->transposed && d->kD == 7 && d->kH == 7 && d->kW == 7 && d->stride == 2 && d->pad == 3,
              "crn_convt7_tc_dgrad: only ConvTranspose3d k=7 s=2 p=3");
  CRN_REQUIRE(d->oD == 2 * d->iD && d->oH == 2 * d->iH && d->oW == 2 * d->iW,
              "crn_convt7_tc_dgrad: output must be 2x input");
  CRN_REQUIRE(d->iW % TX == 0 && d->iH % TY == 0 && d->iD % 8 == 0,
              "crn_convt7_tc_dgrad: input grid must tile by 8x16x8");
  CRN_REQUIRE(!d->y_planar, "crn_convt7_tc_dgrad: planar dy unsupported");
  CRN_REQUIRE(d->Cout % 4 == 0 && d->Cin % 4 == 0 && d->Cin <= 64, "crn_convt7_tc_dgrad: Cout % 4, Cin % 4, Cin <= 64");
  CRN_REQUIRE(d->x_cs % 4 == 0 && d->x_co % 4 == 0 && d->y_cs % 4 == 0 && d->y_co % 4 == 0,
              "crn_convt7_tc_dgrad: channel strides/offsets must be multiples of 4");
  TC5Params p{};
  p.single = crn_single_pass();
  p.in = dy; p.wtc = wtc; p.bias = nullptr; p.out = dx; p.status = status;
  p.N = d->N; p.D = d->iD; p.H = d->iH; p.W = d->iW;
  p.gK = 8 * d->Cout; p.gN = d->Cin; p.cout_cls = d->Cout;
  p.in_cs = d->y_cs; p.in_co = d->y_co; p.out_cs = d->x_cs; p.out_co = d->x_co;
  p.P = (p.gK + 7) / 8;
  p.tiles_x = p.W / TX; p.tiles_y = p.H / TY;
  cudaStream_t st = crn_stream(stream);
  if (p.gN <= 16) {
    if (tc5_use_nd(p.gN)) return launch_tc5<16, 8, true, 4, false, true>(p, st);
    return launch_tc5<16, 8, false, 4, false, true>(p, st);
  }
  if (p.gN <= 32) return launch_tc5<32, 8, false, 4, false, true>(p, st);
  // small grids (stage_4.t1 at 16^3: 32 work items with 4 planes each): one output plane per item fills the machine
  if ((long long)p.N * p.tiles_x * p.tiles_y * (p.D / 4) < kNumSMs / 2) return launch_tc5<64, 1, false, 4, false, true>(p, st);
  return launch_tc5<64, 4, false, 4, false, true>(p, st);
}
