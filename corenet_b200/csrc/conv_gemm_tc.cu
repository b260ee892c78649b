// Implicit-GEMM convolution forward / dgrad on the 5th-gen tensor cores for the WIDE layers (>= 32 channels on
// both sides): the whole ResNet-50 encoder (1x1 and 3x3 Conv2d, model/resnet50.py:61-70,94-108,122-131) and the
// coarse decoder Conv3d layers (model/reconstruction_decoder.py:66,74).  Replaces the cuDNN calls behind them.
//
//   out[row][n] (+)= sum_{tap, c} in[src(row, tap)][c] * W[tap][c][n]          3xTF32, fp32 accumulate
//
//   * CTA tile: 128 output rows (M, one TMEM lane each) x BN output channels; grid = (row tiles, channel tiles,
//     split-K slices).  K is walked in stages of 16 channels of one filter tap.
//   * A (activations): 4 producer warps, thread = output row.  The thread gathers the 64 contiguous bytes of its
//     source row for the stage (zero outside the image / beyond Cin), splits every value into hi = rna_tf32(a),
//     lo = a - hi and stores both in the canonical no-swizzle K-major UMMA layout; loads run 3 stages ahead of
//     the stores (register ring), stores go through a shared-memory ring of NSTAGE stages.
//   * B (weights): pre-split / pre-packed per (channel tile, stage) in the UMMA layout by crn_gemm_tc_pack and
//     streamed by one cp.async.bulk (TMA bulk copy) per stage with an mbarrier transaction count.
//   * one elected thread issues 3 tcgen05.mma per K=8 step (hi*hi + lo*hi + hi*lo) into a TMEM accumulator;
//     every FL stages (24 MMAs) it moves to the other accumulator stage and the 4 epilogue warps fold the finished
//     group into fp32 registers (the tensor core's accumulate truncates: short chains + round-to-nearest sums keep
//     fp32-FFMA accuracy).  Epilogue: + bias / += out, or atomic adds when K is split across CTAs.
// dgrad of a stride-1 convolution is the same kernel on dy with the flipped, transposed filter (packed that way).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;                          // rows per tile
constexpr int KS = 16;                           // channels per stage (two K=8 MMA steps)
constexpr int A_KQ_BYTES = BM * 16;              // one 4-channel K chunk of the A tile (LBO)
constexpr int A_PART_BYTES = 4 * A_KQ_BYTES;     // hi (or lo) part of a stage
constexpr int A_STAGE_BYTES = 2 * A_PART_BYTES;  // 16 KB
constexpr int FL = 4;                            // stages per accumulator flush group (24 MMAs)
constexpr int PF = 3;                            // producer register prefetch distance (stages)
constexpr int NTHREADS = 320;                    // 4 epilogue + 4 producer + MMA + weight-copy warps
constexpr int MAXSTAGE = 8;

struct GTParams {
  const float* in;
  const float* wtc;
  const float* bias;
  float* out;
  int* status;
  int N, iD, iH, iW, oD, oH, oW, kD, kH, kW;
  int sD, sH, sW, pD, pH, pW;       // per-axis stride / (left) padding of the gather
  int gK, gN;                       // channels in / out
  int in_cs, in_co, out_cs, out_co;
  int kchunks;                      // ceil(gK / 16)
  int nstages;                      // taps * kchunks
  int ksplit, accumulate;
  int single;                       // 1: single-pass TF32 (hi x hi only)
  int dbg;                          // debug bits: 1 timeline stamps, 2 skip epilogue stores, 4 skip gathers, 8 skip MMAs
  // transposed-gather mode (tmode = 1): out[o] = sum_k in[(o + pad - k) / ts] W[k] over the taps with an exact
  // quotient (ConvTranspose forward = dgrad of a strided convolution).  Output positions are ordered by parity
  // class c = o mod ts so that a 128-row tile has ONE tap set: (oD, oH, oW) hold the per-class extents O / ts,
  // (fD, fH, fW) the real output extents, tiles_per_class the row tiles of one class; the packed weights hold all
  // taps in flipped order (crn_gemm_tc_pack with dgrad != 0).
  int tmode, ts, tsD, tsH, tsW, fD, fH, fW, tiles_per_class;
  long long rows;                   // output rows (tmode: rows of ONE class)
};

struct __align__(8) GTBarriers {
  uint64_t full_a[MAXSTAGE], full_b[MAXSTAGE], empty[MAXSTAGE];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  int abort_flag;
};

// debug timeline (crn_set_flags bit 8): per CTA 8 clock64 stamps, read back with crn_gemm_tc_debug_read
__device__ long long g_gt_dbg[512 * 8];
#define GT_STAMP(slot)                                                                          \
  do {                                                                                          \
    if (p.dbg && blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z) < 512)          \
      g_gt_dbg[(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) * 8 + (slot)] = clock64(); \
  } while (0)

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_kernel(const GTParams p) {
  constexpr int B_KQ_BYTES = BN * 16;
  constexpr int B_PART_BYTES = 4 * B_KQ_BYTES;
  constexpr int B_STAGE_BYTES = 2 * B_PART_BYTES;
  constexpr int NSTAGE = BN == 128 ? 6 : 8;
  constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* abuf = smem;
  uint8_t* bbuf = smem + NSTAGE * A_STAGE_BYTES;
  GTBarriers* B = reinterpret_cast<GTBarriers*>(bbuf + NSTAGE * B_STAGE_BYTES);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)tc::uniform_u32((uint32_t)tid >> 5);   // warp-uniform for the compiler (role branches)
  const int mt = blockIdx.x, nt = blockIdx.y, z = blockIdx.z;
  if (tid == 0) GT_STAMP(0);
  // transposed-gather mode: parity class of this tile, its taps per axis (k = r + ts*j, j < K_c) and the input
  // offset q of tap j = 0 (i = m + q - j)
  int mtile = mt, cls_z = 0, cls_y = 0, cls_x = 0;
  int rz = 0, ry = 0, rx = 0, Kz = p.kD, Ky = p.kH, Kx = p.kW, qz = 0, qy = 0, qx = 0;
  int nstages = p.nstages;
  if (p.tmode) {
    const int cls = mt / p.tiles_per_class;
    mtile = mt - cls * p.tiles_per_class;
    cls_x = cls % p.tsW; cls_y = (cls / p.tsW) % p.tsH; cls_z = cls / (p.tsW * p.tsH);
    rz = (cls_z + p.pD) % p.tsD; ry = (cls_y + p.pH) % p.tsH; rx = (cls_x + p.pW) % p.tsW;
    Kz = p.kD > rz ? (p.kD - rz + p.tsD - 1) / p.tsD : 0;
    Ky = p.kH > ry ? (p.kH - ry + p.tsH - 1) / p.tsH : 0;
    Kx = p.kW > rx ? (p.kW - rx + p.tsW - 1) / p.tsW : 0;
    qz = (cls_z + p.pD - rz) / p.tsD; qy = (cls_y + p.pH - ry) / p.tsH; qx = (cls_x + p.pW - rx) / p.tsW;
    nstages = Kz * Ky * Kx * p.kchunks;
  }
  // (the 64-bit divisions go through a subroutine and lose the compiler's "warp-uniform" tag: restore it, the MMA
  // warp's loop bounds must be uniform for its descriptors to stay in uniform registers)
  const int s0 = (int)tc::uniform_u32((uint32_t)((long long)nstages * z / p.ksplit));
  const int s1 = (int)tc::uniform_u32((uint32_t)((long long)nstages * (z + 1) / p.ksplit));
  const int nst = s1 - s0;
  // stage -> (filter tap, 16-channel chunk); tmode: tap of the class list -> real filter tap
  auto stage_tap = [&](int s, int& kz, int& ky, int& kx, int& kc) {
    const int tap = s / p.kchunks;
    kc = s - tap * p.kchunks;
    if (p.tmode) {
      const int jx = tap % Kx; const int t2 = tap / Kx;
      const int jy = t2 % Ky, jz = t2 / Ky;
      kz = jz; ky = jy; kx = jx;                      // class-local tap indices j
    } else {
      kx = tap % p.kW; const int t2 = tap / p.kW;
      ky = t2 % p.kH; kz = t2 / p.kH;
    }
  };

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      tc::mbar_init(&B->full_a[i], 128); tc::mbar_init(&B->full_b[i], 1); tc::mbar_init(&B->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&B->acc_full[i], 1); tc::mbar_init(&B->acc_empty[i], 128); }
    B->abort_flag = 0;
    tc::mbar_fence_init();
  }
  if (warp == 8) tc::tmem_alloc(&B->tmem_base, TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tc::uniform_u32(B->tmem_base);
  const uint32_t abuf_u32 = tc::smem_u32(abuf), bbuf_u32 = tc::smem_u32(bbuf);
  if (tid == 0) GT_STAMP(1);
  auto fail = [&]() { B->abort_flag = 1; *p.status = 1; };
  volatile int* ab = &B->abort_flag;

  if (warp < 4) {
    // ============================ EPILOGUE
    float sum[BN];
#pragma unroll
    for (int e = 0; e < BN; ++e) sum[e] = 0.f;
    const int G = (nst + FL - 1) / FL;
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
    if (tid == 0) GT_STAMP(4);
    if (!dead && !(p.dbg & 2)) {
      // All MMAs have completed, so the operand ring is free: each warp transposes its 32 x BN tile through it and
      // writes whole rows (BN * 4 contiguous bytes per instruction instead of 32 scattered 16-byte pieces).
      constexpr int LDS = BN + 4;                                  // padded row (floats): conflict-free float4 access
      float* tile = reinterpret_cast<float*>(abuf) + warp * 32 * LDS;
#pragma unroll
      for (int c = 0; c < BN; c += 4)
        *reinterpret_cast<float4*>(tile + lane * LDS + c) = make_float4(sum[c], sum[c + 1], sum[c + 2], sum[c + 3]);
      __syncwarp();
      constexpr int RPI = 128 / BN;                                // rows per instruction (1 or 2)
      const int col = 4 * (lane % (BN / 4)), rsub = lane / (BN / 4);
      const int ncol = p.gN - nt * BN;                             // valid columns of this tile (multiple of 4)
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias && !p.accumulate && (p.ksplit == 1 || z == 0) && col < ncol) {   // parameters may be 4-byte aligned views
        const float* bp = p.bias + nt * BN + col;
        bv = make_float4(__ldg(bp), __ldg(bp + 1), __ldg(bp + 2), __ldg(bp + 3));
      }
      if (col < ncol) {
#pragma unroll 4
        for (int r0 = 0; r0 < 32; r0 += RPI) {
          const int r = r0 + rsub;
          long long row = (long long)mtile * BM + warp * 32 + r;
          if (row >= p.rows) break;
          float4 o = *reinterpret_cast<const float4*>(tile + r * LDS + col);
          o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
          if (p.tmode) {                                  // class row (n, m) -> output position ts*m + c
            long long t = row;
            const int mx = (int)(t % p.oW); t /= p.oW;
            const int my = (int)(t % p.oH); t /= p.oH;
            const int mz = (int)(t % p.oD); const long long nn = t / p.oD;
            row = ((nn * p.fD + (mz * p.tsD + cls_z)) * p.fH + (my * p.tsH + cls_y)) * p.fW + (mx * p.tsW + cls_x);
          }
          float* dst = p.out + row * p.out_cs + p.out_co + nt * BN + col;
          if (p.ksplit > 1) {
            atomicAdd(reinterpret_cast<float4*>(dst), o);
          } else {
            if (p.accumulate) {
              const float4 old = *reinterpret_cast<const float4*>(dst);
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *reinterpret_cast<float4*>(dst) = o;
          }
        }
      }
    }
  } else if (warp < 8) {
    // ============================ PRODUCERS: thread = output row; gather + hi/lo split into the A ring
    const int m = tid - 128;
    const long long row = (long long)mtile * BM + m;
    const bool row_ok = row < p.rows;
    int n = 0, oz = 0, oy = 0, ox = 0;
    if (row_ok) {
      long long r = row;
      ox = (int)(r % p.oW); r /= p.oW;
      oy = (int)(r % p.oH); r /= p.oH;
      oz = (int)(r % p.oD); n = (int)(r / p.oD);
    }
    const int bz = oz * p.sD - p.pD, by = oy * p.sH - p.pH, bx = ox * p.sW - p.pW;
    auto load_stage = [&](int s, float4 (&v)[4]) {
      int kz, ky, kx, kc;
      stage_tap(s, kz, ky, kx, kc);
      // normal: i = o*stride - pad + k;   transposed gather: i = m + q - j
      const int iz = p.tmode ? oz + qz - kz : bz + kz;
      const int iy = p.tmode ? oy + qy - ky : by + ky;
      const int ix = p.tmode ? ox + qx - kx : bx + kx;
      const bool ok = row_ok && (unsigned)iz < (unsigned)p.iD && (unsigned)iy < (unsigned)p.iH &&
                      (unsigned)ix < (unsigned)p.iW;
      const long long off = ((((long long)n * p.iD + iz) * p.iH + iy) * p.iW + ix) * p.in_cs + p.in_co + kc * KS;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[q] = (ok && kc * KS + q * 4 < p.gK && !(p.dbg & 4)) ? __ldg(reinterpret_cast<const float4*>(p.in + off + q * 4))
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    float4 buf[PF][4];
#pragma unroll
    for (int j = 0; j < PF; ++j)
      if (j < nst) load_stage(s0 + j, buf[j]);
    bool dead = false;
    for (int i0 = 0; i0 < nst && !dead; i0 += PF) {
#pragma unroll
      for (int j = 0; j < PF; ++j) {
        const int i = i0 + j;
        if (i >= nst || dead) continue;
        const int slot = i % NSTAGE;
        const uint32_t use = (uint32_t)(i / NSTAGE);
        if (use > 0 && !tc::mbar_wait(&B->empty[slot], (use - 1) & 1, ab)) { fail(); dead = true; continue; }
        uint8_t* dst = abuf + slot * A_STAGE_BYTES + m * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 hi, lo;
          tc::split_tf32(buf[j][q].x, hi.x, lo.x); tc::split_tf32(buf[j][q].y, hi.y, lo.y);
          tc::split_tf32(buf[j][q].z, hi.z, lo.z); tc::split_tf32(buf[j][q].w, hi.w, lo.w);
          *reinterpret_cast<float4*>(dst + q * A_KQ_BYTES) = hi;
          *reinterpret_cast<float4*>(dst + A_PART_BYTES + q * A_KQ_BYTES) = lo;
        }
        tc::fence_async_smem();
        tc::mbar_arrive(&B->full_a[slot]);
        if (i + PF < nst) load_stage(s0 + i + PF, buf[j]);
      }
    }
  } else if (warp == 8) {
    // ============================ MMA ISSUER (one elected thread)
    {  // converged warp, elected lane issues (tc_common.cuh elect_one)
      const int p_single = p.single;
      constexpr uint32_t idesc = tc::make_idesc_tf32(128, BN, 0, 0);
      int st = 0;
      bool dead = false;
      for (int i = 0; i < nst && !dead; ++i) {
        const int slot = i % NSTAGE;
        const uint32_t ph = (uint32_t)(i / NSTAGE) & 1;
        if (i % FL == 0) {
          const int g = i / FL;
          st = g & 1;
          if (g >= 2 && !tc::mbar_wait_all(&B->acc_empty[st], (uint32_t)((g >> 1) - 1) & 1, ab)) { fail(); dead = true; break; }
          tc::fence_after_sync();
        }
        if (!tc::mbar_wait_all(&B->full_a[slot], ph, ab) || !tc::mbar_wait_all(&B->full_b[slot], ph, ab)) { fail(); dead = true; break; }
        tc::fence_after_sync();
        if (i == 0) GT_STAMP(2);
        const uint32_t a_base = abuf_u32 + slot * A_STAGE_BYTES, b_base = bbuf_u32 + slot * B_STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          if (p.dbg & 8) continue;
          const uint64_t dah = tc::make_desc(a_base + ks * 2 * A_KQ_BYTES, A_KQ_BYTES, 128);
          const uint64_t dal = tc::make_desc(a_base + A_PART_BYTES + ks * 2 * A_KQ_BYTES, A_KQ_BYTES, 128);
          const uint64_t dbh = tc::make_desc(b_base + ks * 2 * B_KQ_BYTES, B_KQ_BYTES, 128);
          const uint64_t dbl = tc::make_desc(b_base + B_PART_BYTES + ks * 2 * B_KQ_BYTES, B_KQ_BYTES, 128);
          const uint32_t d = tmem + st * BN;
          tc::mma_tf32_e(d, dah, dbh, idesc, (i % FL == 0 && ks == 0) ? 0u : 1u);
          if (!p_single) {
            tc::mma_tf32_e(d, dal, dbh, idesc, 1u);
            tc::mma_tf32_e(d, dah, dbl, idesc, 1u);
          }
        }
        tc::commit_e(&B->empty[slot]);
        if (i % FL == FL - 1 || i == nst - 1) tc::commit_e(&B->acc_full[st]);
      }
      GT_STAMP(3);
    }
  } else {
    // ============================ WEIGHT COPIES: one cp.async.bulk per stage
    if (lane == 0) {
      bool dead = false;
      for (int i = 0; i < nst && !dead; ++i) {
        const int slot = i % NSTAGE;
        const uint32_t use = (uint32_t)(i / NSTAGE);
        if (use > 0 && !tc::mbar_wait(&B->empty[slot], (use - 1) & 1, ab)) { fail(); dead = true; break; }
        size_t blk = (size_t)(s0 + i);
        if (p.tmode) {                          // class-local tap j -> filter tap k = r + ts*j -> packed (flipped) block
          int kz, ky, kx, kc;
          stage_tap(s0 + i, kz, ky, kx, kc);
          const int t = ((rz + p.tsD * kz) * p.kH + (ry + p.tsH * ky)) * p.kW + (rx + p.tsW * kx);
          blk = (size_t)(p.kD * p.kH * p.kW - 1 - t) * p.kchunks + kc;
        }
        const float* src = p.wtc + ((size_t)nt * p.nstages + blk) * (B_STAGE_BYTES / 4);
        const uint32_t bar = tc::smem_u32(&B->full_b[slot]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)B_STAGE_BYTES)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                bbuf_u32 + slot * B_STAGE_BYTES),
            "l"(src), "r"((uint32_t)B_STAGE_BYTES), "r"(bar)
            : "memory");
      }
    }
  }
  // ---- teardown
  if (tid == 0) GT_STAMP(5);
  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0) GT_STAMP(6);
  if (warp == 8) tc::tmem_dealloc(tmem, TMEM_COLS);
  if (tid == 256) GT_STAMP(7);
}

int gt_bn(int N) { return N <= 64 ? 64 : 128; }

// one launch packs every listed layer: PyTorch conv weight [Cout][Cin][taps] -> per (channel tile, stage) blocks
// [hi|lo][kq 4][BN rows][4]  (stage = (tap, 16-channel chunk); canonical K-major UMMA layout, zero padded)
//   fwd  : k = ci, n = co, tap t          dgrad: k = co, n = ci, tap taps-1-t
__device__ __forceinline__ int gt_find_item(const int64_t* offsets, int n, int64_t e) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (offsets[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

// one thread = the 4 K values (e) of one (channel tile, stage, kq, row): <= 4 reads, one float4 hi + one float4 lo
// store (consecutive threads = consecutive rows = consecutive 16-byte slots of the packed block)
__global__ void gemm_tc_pack_kernel(const crn_gemm_tc_pack_item* items, const int64_t* offsets, int n, int64_t total) {
  const int64_t total4 = total >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int li = gt_find_item(offsets, n, i << 2);
    const crn_gemm_tc_pack_item it = items[li];
    int64_t r = i - (offsets[li] >> 2);                // (nt, stage, kq, nrow)
    const int K = it.dgrad ? it.Cout : it.Cin, Nn = it.dgrad ? it.Cin : it.Cout;
    const int BN = Nn <= 64 ? 64 : 128;
    const int kchunks = (K + KS - 1) / KS;
    const int nrow = (int)(r % BN); r /= BN;
    const int kq = (int)(r & 3); r >>= 2;
    const int nstages = it.taps * kchunks;
    const int s = (int)(r % nstages); const int nt = (int)(r / nstages);
    const int tap = s / kchunks, kc = s - tap * kchunks;
    const int k0 = kc * KS + kq * 4, nn = nt * BN + nrow;
    const int ts = it.dgrad ? it.taps - 1 - tap : tap;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (nn < Nn) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + e;
        if (k < K) {
          const int co = it.dgrad ? k : nn, ci = it.dgrad ? nn : k;
          v[e] = __ldg(it.src + ((int64_t)co * it.Cin + ci) * it.taps + ts);
        }
      }
    }
    float4 hi, lo;
    tc::split_tf32(v[0], hi.x, lo.x); tc::split_tf32(v[1], hi.y, lo.y);
    tc::split_tf32(v[2], hi.z, lo.z); tc::split_tf32(v[3], hi.w, lo.w);
    float* blk = it.dst + ((int64_t)nt * nstages + s) * (2 * 4 * BN * 4) + ((int64_t)kq * BN + nrow) * 4;
    *reinterpret_cast<float4*>(blk) = hi;
    *reinterpret_cast<float4*>(blk + 4 * BN * 4) = lo;
  }
}

template <int BN>
int launch_gt(const GTParams& p, dim3 grid, cudaStream_t st) {
  constexpr int NSTAGE = BN == 128 ? 6 : 8;
  constexpr int B_STAGE_BYTES = 2 * 4 * BN * 16;
  const size_t smem = (size_t)NSTAGE * (A_STAGE_BYTES + B_STAGE_BYTES) + sizeof(GTBarriers) + 64;
  auto kern = gemm_tc_kernel<BN>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      crn_set_error("conv_gemm_tc: cannot set %zu bytes of dynamic shared memory", smem);
      return CRN_ERR_LAUNCH;
    }
    configured = true;
  }
  kern<<<grid, NTHREADS, smem, st>>>(p);
  CRN_LAUNCH_CHECK("conv_gemm_tc");
  return CRN_OK;
}

}  // namespace

extern "C" int64_t crn_gemm_tc_packed_floats(int32_t K, int32_t N, int32_t taps) {
  const int BN = gt_bn(N);
  const int64_t ntiles = (N + BN - 1) / BN, kchunks = (K + KS - 1) / KS;
  return ntiles * taps * kchunks * (2 * 4 * BN * 4);
}

extern "C" int crn_gemm_tc_pack(const crn_gemm_tc_pack_item* items, const int64_t* offsets, int32_t n, int64_t total,
                                void* stream) {
  CRN_REQUIRE(items && offsets && n > 0 && total > 0, "crn_gemm_tc_pack: bad args");
  int64_t blocks = (total / 4 + 255) / 256;
  if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
  gemm_tc_pack_kernel<<<(unsigned)blocks, 256, 0, crn_stream(stream)>>>(items, offsets, n, total);
  CRN_LAUNCH_CHECK("gemm_tc_pack");
  return CRN_OK;
}

// kind 0: y = conv(x) + bias (any stride) or, for a transposed descriptor, y = conv_transpose(x) + bias;
// kind 1: dx = conv^T(dy) of a plain convolution (any stride).  The two "transposed gather" cases (transposed forward,
// strided dgrad) need weights packed with dgrad != 0 (flipped taps; for the transposed conv the [Cin][Cout][taps]
// parameter is passed as Cout' = Cin, Cin' = Cout).
extern "C" int crn_conv_gemm_tc(const crn_conv_desc* d, int32_t kind, const float* in, const float* wtc,
                                const float* bias, float* out, int32_t accumulate, int32_t* status, void* stream) {
  CRN_REQUIRE(d && in && wtc && out && status, "crn_conv_gemm_tc: null pointer");
  CRN_REQUIRE(!d->y_planar && !d->bias_n_stride, "crn_conv_gemm_tc: planar output / per-scene bias unsupported");
  CRN_REQUIRE(!(d->transposed && kind == 1), "crn_conv_gemm_tc: dgrad of a transposed conv = kind 0 on dy (see engine)");
  CRN_REQUIRE(d->Cin % 4 == 0 && d->Cout % 4 == 0 && d->x_cs % 4 == 0 && d->x_co % 4 == 0 && d->y_cs % 4 == 0 &&
                  d->y_co % 4 == 0,
              "crn_conv_gemm_tc: channels, strides and offsets must be multiples of 4");
  GTParams p{};
  p.in = in; p.wtc = wtc; p.out = out; p.status = status; p.accumulate = accumulate;
  p.dbg = (crn_get_flags() >> 8) & 15;
  p.single = crn_single_pass();
  p.N = d->N; p.kD = d->kD; p.kH = d->kH; p.kW = d->kW;
  const int K3[3] = {d->kD, d->kH, d->kW}, I3[3] = {d->iD, d->iH, d->iW}, O3[3] = {d->oD, d->oH, d->oW};
  int s3[3], p3[3];
  for (int a = 0; a < 3; ++a) {
    const bool trivial = K3[a] == 1 && I3[a] == 1 && O3[a] == 1;
    s3[a] = trivial ? 1 : d->stride;
    p3[a] = K3[a] > 1 ? d->pad : 0;
  }
  const bool strided = s3[0] > 1 || s3[1] > 1 || s3[2] > 1;
  const bool tgather = (kind == 0 && d->transposed) || (kind == 1 && strided);
  p.ts = 1; p.tsD = p.tsH = p.tsW = 1;
  if (kind == 0) {
    p.bias = accumulate ? nullptr : bias;
    p.gK = d->Cin; p.gN = d->Cout;
    p.in_cs = d->x_cs; p.in_co = d->x_co; p.out_cs = d->y_cs; p.out_co = d->y_co;
    p.iD = d->iD; p.iH = d->iH; p.iW = d->iW; p.oD = d->oD; p.oH = d->oH; p.oW = d->oW;
    p.sD = s3[0]; p.sH = s3[1]; p.sW = s3[2]; p.pD = p3[0]; p.pH = p3[1]; p.pW = p3[2];
  } else {
    p.bias = nullptr;
    p.gK = d->Cout; p.gN = d->Cin;
    p.in_cs = d->y_cs; p.in_co = d->y_co; p.out_cs = d->x_cs; p.out_co = d->x_co;
    p.iD = d->oD; p.iH = d->oH; p.iW = d->oW; p.oD = d->iD; p.oH = d->iH; p.oW = d->iW;
    p.sD = p.sH = p.sW = 1;
    p.pD = K3[0] - 1 - p3[0]; p.pH = K3[1] - 1 - p3[1]; p.pW = K3[2] - 1 - p3[2];
  }
  int nclasses = 1;
  if (tgather) {
    // output = the FINE grid (p.o*), ordered by parity class; (p.o*) become the per-class extents
    CRN_REQUIRE(p.oD % s3[0] == 0 && p.oH % s3[1] == 0 && p.oW % s3[2] == 0,
                "crn_conv_gemm_tc: transposed gather needs output extents divisible by the stride");
    p.tmode = 1; p.ts = d->stride; p.tsD = s3[0]; p.tsH = s3[1]; p.tsW = s3[2];
    p.fD = p.oD; p.fH = p.oH; p.fW = p.oW;
    p.oD /= s3[0]; p.oH /= s3[1]; p.oW /= s3[2];
    p.sD = p.sH = p.sW = 1; p.pD = p3[0]; p.pH = p3[1]; p.pW = p3[2];
    nclasses = s3[0] * s3[1] * s3[2];
  }
  p.rows = (long long)p.N * p.oD * p.oH * p.oW;
  if (p.rows <= 0) return CRN_OK;
  p.tiles_per_class = (int)crn_ceil_div(p.rows, BM);
  p.kchunks = (p.gK + KS - 1) / KS;
  p.nstages = d->kD * d->kH * d->kW * p.kchunks;
  const int BN = gt_bn(p.gN);
  // split-K when the tile grid cannot fill the machine; needs a dense, exclusively owned output to zero first
  p.ksplit = 1;
  const long long blocks = (long long)nclasses * p.tiles_per_class * crn_ceil_div(p.gN, BN);
  long long min_stages = p.nstages;
  if (tgather) {           // the smallest class: floor(K / ts) taps per axis
    min_stages = p.kchunks;
    for (int a = 0; a < 3; ++a) min_stages *= (K3[a] / s3[a] > 0 ? K3[a] / s3[a] : 1);
  }
  const long long out_rows = p.rows * nclasses;
  const bool dense_out = p.out_co == 0 && p.out_cs == p.gN;
  if (blocks < kNumSMs && min_stages >= 8 && dense_out && !(crn_get_flags() & 16)) {
    long long ks = kNumSMs / blocks;                      // one wave: a second one costs a full CTA set-up
    if (ks > min_stages / 4) ks = min_stages / 4;
    if (ks > 32) ks = 32;
    if (ks > 1) {
      p.ksplit = (int)ks;
      if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * out_rows * p.out_cs, crn_stream(stream));
    }
  }
  cudaStream_t st = crn_stream(stream);
  dim3 grid((unsigned)(nclasses * p.tiles_per_class), (unsigned)crn_ceil_div(p.gN, BN), (unsigned)p.ksplit);
  return BN == 64 ? launch_gt<64>(p, grid, st) : launch_gt<128>(p, grid, st);
}

#ifdef CRN_DIAG
extern "C" int crn_gemm_tc_debug_read(long long* host_dst, int32_t n);
extern "C" int crn_gemm_tc_debug_read(long long* host_dst, int32_t n) {
  return cudaMemcpyFromSymbol(host_dst, g_gt_dbg, sizeof(long long) * (n > 4096 ? 4096 : n)) == cudaSuccess ? CRN_OK
                                                                                                             : CRN_ERR_LAUNCH;
}
#endif  // CRN_DIAG
