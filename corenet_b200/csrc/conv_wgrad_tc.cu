// Weight gradients of the WIDE convolutions on the 5th-gen tensor cores (tcgen05, 3xTF32):
//   dW[tap][ci][co] += sum_rows x[src(row, tap)][ci] * dy[row][co]
// Replaces the cuDNN backward-filter calls behind the encoder Conv2d layers (model/resnet50.py:61-70,94-108,
// 122-131) and the coarse decoder Conv3d / ConvTranspose3d layers (model/reconstruction_decoder.py:66-77).
//
// The reduction index (rows = pixels / voxels) is the slow index of both operands in memory (channels-last), so
// both are MN-major UMMA operands.  For tf32 that exists only in the "128-byte swizzle, 32-byte base" layout
// (descriptor layout type 1; pinned on hardware by crn_tc_probe_mn): one reduction row is one 128-byte line holding
// 32 channels, 32-byte chunks XOR-swizzled with the row index -- i.e. channels-last rows almost as they are.
//
//   * CTA = (tap, 128-channel M tile, BN-channel N tile, row slice).  The M side is whichever operand the host
//     chose (x or dy); the other one is the N side.  One operand is read at the base rows, the other ("shifted")
//     at base*stride - pad + tap (conv: base = dy grid, shifted = x; transposed conv: base = x grid, shifted = dy).
//   * 4 producer warps stage 16 rows x (128 + BN) channels per stage: float4 gathers (zero outside the grid /
//     beyond C), hi/lo split, swizzled st.shared; loads run PF stages ahead in registers.
//   * one elected thread issues 3 tcgen05.mma per K=8 rows (hi*hi + lo*hi + hi*lo); the TMEM accumulator is folded
//     into fp32 registers every FLW stages (48 MMAs) by the 4 epilogue warps, like the forward kernel.
//   * epilogue: dW tile += sums (plain read-modify-write when the CTA owns the tile, atomics when rows are split).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int KR = 16;                           // rows per stage (two K=8 MMA steps)
constexpr int CB_BYTES = KR * 128;               // one 32-channel block of a stage (LBO)
constexpr int P_PART_BYTES = 4 * CB_BYTES;       // hi (or lo) of the 128-channel M operand
constexpr int P_STAGE_BYTES = 2 * P_PART_BYTES;  // 16 KB
constexpr int FLW = 8;                           // stages per accumulator flush group (48 MMAs)
constexpr int NTHREADS = 288;                    // 4 epilogue + 4 producer + MMA warps
constexpr int MAXSTAGE = 8;

struct WOperand {
  const float* ptr;
  int cs, co, C;      // channel stride / offset / count
  int shifted;        // 1: gathered at base*stride - pad + tap; 0: read at the base rows
  int D, H, W;        // its spatial dims
};

struct WTParams {
  WOperand P, Q;       // M-side, N-side
  float* dw;           // [tap][CinP][CoutP]
  int* status;
  int CinP, CoutP;
  int single;          // 1: single-pass TF32 (hi x hi only)
  int m_is_cin;        // D[m][n]: m = ci, n = co (1) or m = co, n = ci (0)
  int N, bD, bH, bW;   // base grid
  int kD, kH, kW, sD, sH, sW, pD, pH, pW;
  int mtiles, ntiles, rsplit;
  long long rows;      // base rows
  long long rows_per_split;   // multiple of KR
};

struct __align__(8) WTBarriers {
  uint64_t full[MAXSTAGE], empty[MAXSTAGE];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  int abort_flag;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1) wgrad_tc_kernel(const WTParams p) {
  constexpr int QB = BN / 32;                     // 32-channel blocks of the N operand
  constexpr int Q_PART_BYTES = QB * CB_BYTES;
  constexpr int Q_STAGE_BYTES = 2 * Q_PART_BYTES;
  constexpr int STAGE_BYTES = P_STAGE_BYTES + Q_STAGE_BYTES;
  constexpr int NSTAGE = BN == 128 ? 6 : 8;
  constexpr int PF = BN == 128 ? 2 : 3;
  constexpr int NL = 4 + QB;                      // float4 loads per thread and stage
  constexpr int TMEM_COLS = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms need 512 B alignment
  WTBarriers* B = reinterpret_cast<WTBarriers*>(smem + NSTAGE * STAGE_BYTES);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)tc::uniform_u32((uint32_t)tid >> 5);   // warp-uniform for the compiler (role branches)
  int bx = blockIdx.x;
  const int nt = bx % p.ntiles; bx /= p.ntiles;
  const int mt = bx % p.mtiles; const int tap = bx / p.mtiles;
  const long long rbeg = (long long)blockIdx.y * p.rows_per_split;
  long long rend = rbeg + p.rows_per_split;
  if (rend > p.rows) rend = p.rows;
  const int nst = rend > rbeg ? (int)((rend - rbeg + KR - 1) / KR) : 0;

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) { tc::mbar_init(&B->full[i], 128); tc::mbar_init(&B->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&B->acc_full[i], 1); tc::mbar_init(&B->acc_empty[i], 128); }
    B->abort_flag = 0;
    tc::mbar_fence_init();
  }
  if (warp == 8) tc::tmem_alloc(&B->tmem_base, TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tc::uniform_u32(B->tmem_base);
  const uint32_t smem_u32 = tc::smem_u32(smem);
  auto fail = [&]() { B->abort_flag = 1; *p.status = 1; };
  volatile int* ab = &B->abort_flag;

  if (warp < 4) {
    // ============================ EPILOGUE
    float sum[BN];
#pragma unroll
    for (int e = 0; e < BN; ++e) sum[e] = 0.f;
    const int G = (nst + FLW - 1) / FLW;
    bool dead = false;
    for (int g = 0; g < G && !dead; ++g) {
      const int st = g & 1;
      if (!tc::mbar_wait(&B->acc_full[st], (uint32_t)(g >> 1) & 1, ab)) { fail(); dead = true; break; }
      tc::fence_after_sync();
      const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + st * BN;
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tc::tmem_ld16(ta + c0, v);
#pragma unroll
        for (int e = 0; e < 16; ++e) sum[c0 + e] += v[e];
      }
      tc::fence_before_sync();
      tc::mbar_arrive(&B->acc_empty[st]);
    }
    const int m = mt * 128 + warp * 32 + lane;
    if (p.m_is_cin && !dead && nst > 0) {
      // dW rows are contiguous in co = n: transpose the warp's 32 x BN tile through the (now idle) operand ring and
      // add whole rows per instruction
      constexpr int LDS = BN + 4;
      float* tile = reinterpret_cast<float*>(smem) + warp * 32 * LDS;
#pragma unroll
      for (int c = 0; c < BN; c += 4)
        *reinterpret_cast<float4*>(tile + lane * LDS + c) = make_float4(sum[c], sum[c + 1], sum[c + 2], sum[c + 3]);
      __syncwarp();
      constexpr int RPI = 128 / BN;
      const int col = 4 * (lane % (BN / 4)), rsub = lane / (BN / 4);
      const int ncol = p.Q.C - nt * BN;
      if (col < ncol) {
#pragma unroll 4
        for (int r0 = 0; r0 < 32; r0 += RPI) {
          const int r = r0 + rsub;
          const int mm = mt * 128 + warp * 32 + r;
          if (mm >= p.P.C) break;
          const float4 o = *reinterpret_cast<const float4*>(tile + r * LDS + col);
          float* dst = p.dw + ((long long)tap * p.CinP + mm) * p.CoutP + nt * BN + col;
          if (p.rsplit > 1) {
            atomicAdd(reinterpret_cast<float4*>(dst), o);
          } else {
            float4 old = *reinterpret_cast<const float4*>(dst);
            old.x += o.x; old.y += o.y; old.z += o.z; old.w += o.w;
            *reinterpret_cast<float4*>(dst) = old;
          }
        }
      }
    }
    if (m < p.P.C && !dead && nst > 0) {
      const int ncol = p.Q.C - nt * BN;              // valid N channels of this tile (multiple of 4)
      if (p.m_is_cin) {
        // handled above (whole-warp transposed write)
      } else {
        float* dst = p.dw + ((long long)tap * p.CinP + nt * BN) * p.CoutP + m;
#pragma unroll
        for (int c = 0; c < BN; c += 4) {
          if (c >= ncol) break;
          float* d0 = dst + (long long)c * p.CoutP;
          if (p.rsplit > 1) {
            atomicAdd(d0, sum[c]); atomicAdd(d0 + p.CoutP, sum[c + 1]);
            atomicAdd(d0 + 2 * p.CoutP, sum[c + 2]); atomicAdd(d0 + 3 * p.CoutP, sum[c + 3]);
          } else {
            d0[0] += sum[c]; d0[p.CoutP] += sum[c + 1]; d0[2 * p.CoutP] += sum[c + 2]; d0[3 * p.CoutP] += sum[c + 3];
          }
        }
      }
    }
  } else if (warp < 8) {
    // ============================ PRODUCERS: thread = (row of the stage, 16-byte column of the 128-byte line)
    const int pt = tid - 128;
    const int r = pt >> 3, l8 = pt & 7;
    const int kx = tap % p.kW; const int t2 = tap / p.kW;
    const int ky = t2 % p.kH, kz = t2 / p.kH;
    // swizzled byte offset of this thread's 16 bytes inside a 32-channel block
    const uint32_t sw = (uint32_t)r * 128 + (uint32_t)((((l8 >> 1) ^ (r & 3)) << 5) + ((l8 & 1) << 4));
    auto row_offsets = [&](long long row, long long& offp, long long& offq) {
      offp = -1; offq = -1;
      if (row >= rend) return;
      long long q = row;
      const int x = (int)(q % p.bW); q /= p.bW;
      const int y = (int)(q % p.bH); q /= p.bH;
      const int zz = (int)(q % p.bD); const int n = (int)(q / p.bD);
      auto at = [&](const WOperand& o) -> long long {
        if (!o.shifted) return row * o.cs + o.co;
        const int iz = zz * p.sD - p.pD + kz, iy = y * p.sH - p.pH + ky, ix = x * p.sW - p.pW + kx;
        if ((unsigned)iz >= (unsigned)o.D || (unsigned)iy >= (unsigned)o.H || (unsigned)ix >= (unsigned)o.W) return -1;
        return ((((long long)n * o.D + iz) * o.H + iy) * o.W + ix) * o.cs + o.co;
      };
      offp = at(p.P); offq = at(p.Q);
    };
    auto load_stage = [&](int i, float4 (&v)[NL]) {
      long long offp, offq;
      row_offsets(rbeg + (long long)i * KR + r, offp, offq);
      if (offp < 0 || offq < 0) { offp = -1; offq = -1; }      // a zero on either side zeroes the product
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = mt * 128 + j * 32 + l8 * 4;
        v[j] = (offp >= 0 && c < p.P.C) ? __ldg(reinterpret_cast<const float4*>(p.P.ptr + offp + c))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < QB; ++j) {
        const int c = nt * BN + j * 32 + l8 * 4;
        v[4 + j] = (offq >= 0 && c < p.Q.C) ? __ldg(reinterpret_cast<const float4*>(p.Q.ptr + offq + c))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    float4 buf[PF][NL];
#pragma unroll
    for (int j = 0; j < PF; ++j)
      if (j < nst) load_stage(j, buf[j]);
    bool dead = false;
    for (int i0 = 0; i0 < nst && !dead; i0 += PF) {
#pragma unroll
      for (int j = 0; j < PF; ++j) {
        const int i = i0 + j;
        if (i >= nst || dead) continue;
        const int slot = i % NSTAGE;
        const uint32_t use = (uint32_t)(i / NSTAGE);
        if (use > 0 && !tc::mbar_wait(&B->empty[slot], (use - 1) & 1, ab)) { fail(); dead = true; continue; }
        uint8_t* sp = smem + slot * STAGE_BYTES + sw;
#pragma unroll
        for (int q = 0; q < NL; ++q) {
          float4 hi, lo;
          tc::split_tf32(buf[j][q].x, hi.x, lo.x); tc::split_tf32(buf[j][q].y, hi.y, lo.y);
          tc::split_tf32(buf[j][q].z, hi.z, lo.z); tc::split_tf32(buf[j][q].w, hi.w, lo.w);
          uint8_t* d = q < 4 ? sp + q * CB_BYTES : sp + P_STAGE_BYTES + (q - 4) * CB_BYTES;
          const int part = q < 4 ? P_PART_BYTES : Q_PART_BYTES;
          *reinterpret_cast<float4*>(d) = hi;
          *reinterpret_cast<float4*>(d + part) = lo;
        }
        tc::fence_async_smem();
        tc::mbar_arrive(&B->full[slot]);
        if (i + PF < nst) load_stage(i + PF, buf[j]);
      }
    }
  } else {
    // ============================ MMA ISSUER (one elected thread)
    {  // converged warp, elected lane issues (tc_common.cuh elect_one)
      const int p_single = p.single;
      constexpr uint32_t idesc = tc::make_idesc_tf32(128, BN, 1, 1);
      int st = 0;
      bool dead = false;
      for (int i = 0; i < nst && !dead; ++i) {
        const int slot = i % NSTAGE;
        const uint32_t ph = (uint32_t)(i / NSTAGE) & 1;
        if (i % FLW == 0) {
          const int g = i / FLW;
          st = g & 1;
          if (g >= 2 && !tc::mbar_wait(&B->acc_empty[st], (uint32_t)((g >> 1) - 1) & 1, ab)) { fail(); dead = true; break; }
          tc::fence_after_sync();
        }
        if (!tc::mbar_wait(&B->full[slot], ph, ab)) { fail(); dead = true; break; }
        tc::fence_after_sync();
        const uint32_t p_base = smem_u32 + slot * STAGE_BYTES, q_base = p_base + P_STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t dph = tc::make_desc_mn32(p_base + ks * 1024, CB_BYTES, 512);
          const uint64_t dpl = tc::make_desc_mn32(p_base + P_PART_BYTES + ks * 1024, CB_BYTES, 512);
          const uint64_t dqh = tc::make_desc_mn32(q_base + ks * 1024, CB_BYTES, 512);
          const uint64_t dql = tc::make_desc_mn32(q_base + Q_PART_BYTES + ks * 1024, CB_BYTES, 512);
          const uint32_t d = tmem + st * BN;
          tc::mma_tf32_e(d, dph, dqh, idesc, (i % FLW == 0 && ks == 0) ? 0u : 1u);
          if (!p_single) {
            tc::mma_tf32_e(d, dpl, dqh, idesc, 1u);
            tc::mma_tf32_e(d, dph, dql, idesc, 1u);
          }
        }
        tc::commit_e(&B->empty[slot]);
        if (i % FLW == FLW - 1 || i == nst - 1) tc::commit_e(&B->acc_full[st]);
      }
    }
  }
  // ---- teardown
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, TMEM_COLS);
}

template <int BN>
int launch_wt(const WTParams& p, int taps, cudaStream_t st) {
  constexpr int NSTAGE = BN == 128 ? 6 : 8;
  constexpr int STAGE_BYTES = P_STAGE_BYTES + 2 * (BN / 32) * CB_BYTES;
  const size_t smem = (size_t)NSTAGE * STAGE_BYTES + sizeof(WTBarriers) + 1024 + 64;
  auto kern = wgrad_tc_kernel<BN>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      crn_set_error("conv_wgrad_tc: cannot set %zu bytes of dynamic shared memory", smem);
      return CRN_ERR_LAUNCH;
    }
    configured = true;
  }
  dim3 grid((unsigned)(taps * p.mtiles * p.ntiles), (unsigned)p.rsplit);
  kern<<<grid, NTHREADS, smem, st>>>(p);
  CRN_LAUNCH_CHECK("conv_wgrad_tc");
  return CRN_OK;
}

}  // namespace

// dWf[tap][ci][co] += sum_rows x * dy  (same contract as crn_conv_wgrad: dw zeroed by the caller before the first call)
extern "C" int crn_conv_wgrad_tc(const crn_conv_desc* d, const float* x, const float* dy, float* dw_packed,
                                 int32_t* status, void* stream) {
  CRN_REQUIRE(d && x && dy && dw_packed && status, "crn_conv_wgrad_tc: null pointer");
  CRN_REQUIRE(!d->y_planar, "crn_conv_wgrad_tc: planar dy unsupported");
  CRN_REQUIRE(d->Cin % 4 == 0 && d->Cout % 4 == 0 && d->x_cs % 4 == 0 && d->x_co % 4 == 0 && d->y_cs % 4 == 0 &&
                  d->y_co % 4 == 0 && d->CoutP % 4 == 0,
              "crn_conv_wgrad_tc: channels, strides and offsets must be multiples of 4");
  WTParams p{};
  p.dw = dw_packed; p.status = status; p.CinP = d->CinP; p.CoutP = d->CoutP; p.single = crn_single_pass();
  p.N = d->N; p.kD = d->kD; p.kH = d->kH; p.kW = d->kW;
  const int K3[3] = {d->kD, d->kH, d->kW}, I3[3] = {d->iD, d->iH, d->iW}, O3[3] = {d->oD, d->oH, d->oW};
  int s3[3], p3[3];
  for (int a = 0; a < 3; ++a) {
    const bool trivial = K3[a] == 1 && I3[a] == 1 && O3[a] == 1;
    s3[a] = trivial ? 1 : d->stride;
    p3[a] = K3[a] > 1 ? d->pad : 0;
  }
  p.sD = s3[0]; p.sH = s3[1]; p.sW = s3[2]; p.pD = p3[0]; p.pH = p3[1]; p.pW = p3[2];
  WOperand X{x, d->x_cs, d->x_co, d->Cin, d->transposed ? 0 : 1, d->iD, d->iH, d->iW};
  WOperand Y{dy, d->y_cs, d->y_co, d->Cout, d->transposed ? 1 : 0, d->oD, d->oH, d->oW};
  // base grid: the operand that is NOT shifted (conv: dy positions; transposed conv: x positions)
  if (d->transposed) { p.bD = d->iD; p.bH = d->iH; p.bW = d->iW; }
  else { p.bD = d->oD; p.bH = d->oH; p.bW = d->oW; }
  p.rows = (long long)p.N * p.bD * p.bH * p.bW;
  if (p.rows <= 0) return CRN_OK;
  // M side: the operand whose channel count wastes less of the 128-row tiles; N side gets BN = 64 / 128
  auto waste = [](int c) { return (double)((c + 127) / 128 * 128) / c; };
  const bool m_is_cin = waste(d->Cin) <= waste(d->Cout);
  p.m_is_cin = m_is_cin ? 1 : 0;
  p.P = m_is_cin ? X : Y; p.Q = m_is_cin ? Y : X;
  const int BN = p.Q.C <= 64 ? 64 : 128;
  p.mtiles = (p.P.C + 127) / 128; p.ntiles = (p.Q.C + BN - 1) / BN;
  const int taps = d->kD * d->kH * d->kW;
  const long long tiles = (long long)taps * p.mtiles * p.ntiles;
  long long nsplit = crn_ceil_div(2LL * kNumSMs, tiles);
  const long long max_split = crn_ceil_div(p.rows, 8 * KR);        // at least 8 stages per CTA
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  p.rows_per_split = crn_ceil_div(crn_ceil_div(p.rows, nsplit), KR) * KR;
  p.rsplit = (int)crn_ceil_div(p.rows, p.rows_per_split);
  cudaStream_t st = crn_stream(stream);
  return BN == 64 ? launch_wt<64>(p, taps, st) : launch_wt<128>(p, taps, st);
}
