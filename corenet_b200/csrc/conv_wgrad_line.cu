// Weight gradient of the NARROW Conv3d k=5 s=1 p=2 decoder layers (stage_6.c1: 28 -> 16 at 64^3, stage_5.c1:
// 56 -> 32 at 32^3; model/reconstruction_decoder.py:82,91) on tcgen05 tensor cores.  Replaces the cuDNN
// backward-filter call:   dW[tap][ci][co] += sum_vox x[vox + tap - 2][ci] * dy[vox][co].
//
// With 16-56 channels a per-tap GEMM would waste the 128 x N tensor-core tile, so filter taps are STACKED into the
// M and N dimensions of one MMA, using the tf32 MN-major operand layout (descriptor layout type 1: one reduction
// row = one voxel = one 128-byte line of 32 channels, 32-byte chunks XOR-swizzled by the row index), in which a
// 32-channel block is addressed as start + block * LBO -- so "block b" can just as well be "the same line shifted
// by b voxels" (LBO = 128 B) or "the next image row / the lo half" (LBO = one staged line):
//   * the image row (z, y) of x is staged ONCE as [hi | lo] lines of W + 8 voxel rows (zero halo), dy likewise;
//   * CB = 1 (Cin <= 32, Cout <= 16):  M = 128 = {hi, lo} x {row y, row y+1} x 32 ci,   dy voxel row = [hi16 | lo16],
//     N = 160 = 5 kx shifts x 32  ->  ONE MMA per 8 voxels covers 2 ky x 5 kx taps and all four hi/lo products
//     (x_hi+x_lo)(dy_hi+dy_lo), i.e. full fp32-class products;
//   * CB = 2 (Cin <= 64, Cout <= 32):  M = 128 = {hi, lo} x 2 channel blocks,  N = 160 = 5 kx shifts x 32 co for the
//     dy_hi line and again for the dy_lo line (two MMAs into the same accumulator);
//   * accumulators (3 x 160 TMEM columns) live across all image rows a CTA processes for one PASS (CB = 1: pass = kz,
//     accumulators = ky pairs; CB = 2: pass = (kz, ky group), accumulators = ky); they are flushed with atomic
//     adds into dW at pass boundaries and every `flush_every` rows (the tensor core's fp32 accumulate truncates);
//   * x rows slide through a ring (one new row per step, 5 or 3 resident), so every row is staged once per pass;
//   * warp roles: 4 epilogue warps (TMEM -> atomics), 4 producer warps (float4 gathers, hi/lo split, swizzled
//     st.shared), 1 MMA warp (one elected thread).  Persistent grid of one CTA per SM; all waits are bounded.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int NTHREADS = 288;
constexpr int NACC = 3, ACOLS = 160;
constexpr int YSLOTS = 3;
constexpr int MAXX = 9;

struct WLParams {
  const float* x;
  const float* dy;
  float* dw;          // [125][CinP][CoutP]
  int* status;
  int N, D, H, W;
  int Cin, Cout;
  int x_cs, x_co, y_cs, y_co, CinP, CoutP;
  int npass;
  int pass_begin[12];  // prefix sums of line-steps per pass (npass + 1 entries)
  int flush_every;
  int single;          // 1: single-pass TF32 (the lo products are not issued)
  int dbg;             // crn_set_flags bit 8: wait-time accounting of wgrad_line_kernel (crn_wgrad_line_debug_read)
};

// debug: per CTA [mma total, mma wait full_x, full_y, acc_empty, producer total, producer wait empty_x + empty_y,
// epilogue total, epilogue wait acc_full] in clock64 cycles
__device__ long long g_wl_dbg[kNumSMs * 8];
#define WL_TIMED(accum, call)                           \
  ([&]() -> bool {                                      \
    const long long t0__ = p.dbg ? clock64() : 0;       \
    const bool ok__ = (call);                           \
    if (p.dbg) accum += clock64() - t0__;               \
    return ok__;                                        \
  }())

struct __align__(8) WLBarriers {
  uint64_t full_x[MAXX], empty_x[MAXX];
  uint64_t full_y[YSLOTS], empty_y[YSLOTS];
  uint64_t acc_full, acc_empty;
  uint32_t tmem_base;
  int abort_flag;
};

template <int CB>
struct WLCfg {
  static constexpr int XSLOTS = CB == 1 ? 8 : 6;             // ring (CB = 1: + 1 mirror slot of slot 0)
};

// one step = one output image row (pass, n, z, y)
struct Step {
  int pass, kz, klo, nk;   // ky window [klo, klo + nk)
  int n, z, y;
  bool fresh;              // window is (re)loaded completely
};

template <int CB>
__device__ __forceinline__ Step decode_step(const WLParams& p, int t, int t0) {
  Step s;
  int ps = 0;
  while (ps + 1 < p.npass && t >= p.pass_begin[ps + 1]) ++ps;
  s.pass = ps;
  if (CB == 1) { s.kz = ps; s.klo = 0; s.nk = 5; }
  else { s.kz = ps >> 1; s.klo = (ps & 1) ? 3 : 0; s.nk = (ps & 1) ? 2 : 3; }
  const int r = t - p.pass_begin[ps];
  s.y = r % p.H;
  const int pl = r / p.H;                                    // plane index among the valid (n, z) of this pass
  const int zlo = s.kz < 2 ? 2 - s.kz : 0, zcnt = p.D - (s.kz < 2 ? 2 - s.kz : s.kz - 2);
  s.n = pl / zcnt;
  s.z = zlo + pl % zcnt;
  s.fresh = (t == t0) || s.y == 0;
  return s;
}

template <int CB, int W>
__global__ void __launch_bounds__(NTHREADS, 1) wgrad_line_kernel(const WLParams p) {
  using Cfg = WLCfg<CB>;
  constexpr int RUNS = (W + 8) / 8;                 // K=8 runs per image row (W + 4 halo voxels, rounded up)
  constexpr int RS = RUNS * 8;                      // staged voxel rows of an x line (voxel q = row - 4)
  constexpr int XH = RS * 128;                      // one 32-channel block of a line (LBO of the M operand)
  constexpr int XL = 2 * CB * XH;                   // x line slot: CB = 1: [h]; CB = 2: [h][cb]
  constexpr int YROWS = RS + 8;                     // staged voxel rows of a dy line (voxel q = row - 6)
  constexpr int YL = CB * YROWS * 128;              // dy line slot: CB = 1: [hi16|lo16]; CB = 2: [hi line][lo line]
  constexpr int XS = Cfg::XSLOTS, XALL = XS + (CB == 1 ? 1 : 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* xring = smem;
  uint8_t* yring = smem + XALL * XL;
  WLBarriers* B = reinterpret_cast<WLBarriers*>(yring + YSLOTS * YL);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)tc::uniform_u32((uint32_t)tid >> 5);   // warp-uniform for the compiler (role branches)
  const int T = p.pass_begin[p.npass];
  const int t0 = (int)((long long)T * blockIdx.x / gridDim.x), t1 = (int)((long long)T * (blockIdx.x + 1) / gridDim.x);

  if (tid == 0) {
    for (int i = 0; i < MAXX; ++i) { tc::mbar_init(&B->full_x[i], 128); tc::mbar_init(&B->empty_x[i], 1); }
    for (int i = 0; i < YSLOTS; ++i) { tc::mbar_init(&B->full_y[i], 128); tc::mbar_init(&B->empty_y[i], 1); }
    tc::mbar_init(&B->acc_full, 1); tc::mbar_init(&B->acc_empty, 128);
    B->abort_flag = 0;
    tc::mbar_fence_init();
  }
  if (warp == 8) tc::tmem_alloc(&B->tmem_base, 512);
  // halo rows (and padding channels) are zero for the whole kernel: clear both rings once
  for (int i = tid; i < (XALL * XL + YSLOTS * YL) / 16; i += NTHREADS)
    reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tc::uniform_u32(B->tmem_base);
  const uint32_t xring_u32 = tc::smem_u32(xring), yring_u32 = tc::smem_u32(yring);
  auto fail = [&]() { B->abort_flag = 1; *p.status = 1; };
  volatile int* ab = &B->abort_flag;
  auto flush_after = [&](int t, const Step& s) -> bool {
    if (t == t1 - 1) return true;
    if (t + 1 >= p.pass_begin[s.pass + 1]) return true;                 // next step belongs to the next pass
    return (t - t0 + 1) % p.flush_every == 0;
  };

  if (warp < 4) {
    // ============================ EPILOGUE: TMEM -> atomic adds into dW at every flush point
    uint32_t nflush = 0;
    bool dead = false;
    long long tw_e = 0;
    const long long t_begin = p.dbg ? clock64() : 0;
    const int blk = warp;                                     // 32-lane block of the M operand this warp owns
    for (int t = t0; t < t1 && !dead; ++t) {
      const Step s = decode_step<CB>(p, t, t0);
      if (!flush_after(t, s)) continue;
      if (!WL_TIMED(tw_e, tc::mbar_wait(&B->acc_full, nflush & 1, ab))) { fail(); dead = true; break; }
      tc::fence_after_sync();
      int ci, kyoff;
      if (CB == 1) { ci = lane; kyoff = blk >> 1; }           // blocks: (hi, y), (lo, y), (hi, y+1), (lo, y+1)
      else { ci = (blk & 1) * 32 + lane; kyoff = 0; }         // blocks: (hi, cb0), (hi, cb1), (lo, cb0), (lo, cb1)
#pragma unroll 1
      for (int a = 0; a < NACC; ++a) {
        const int ky = (CB == 1 ? 2 * a : s.klo + a) + kyoff;
        const bool acc_ok = CB == 1 ? ky <= 4 : a < s.nk;
#pragma unroll 1
        for (int sh = 0; sh < 5; ++sh) {
          const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + a * ACOLS + sh * 32;
          float v[16], u[16];
          tc::tmem_ld16(ta, v);
          tc::tmem_ld16(ta + 16, u);
          if (!acc_ok || ci >= p.Cin) continue;
          const int kx = 4 - sh;
          const int tap = (s.kz * 5 + ky) * 5 + kx;
          float* dst = p.dw + ((long long)tap * p.CinP + ci) * p.CoutP;
          if (CB == 1) {                                      // columns [hi16 | lo16] of dy: add the halves
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
              if (c < p.Cout)
                atomicAdd(reinterpret_cast<float4*>(dst + c),
                          make_float4(v[c] + u[c], v[c + 1] + u[c + 1], v[c + 2] + u[c + 2], v[c + 3] + u[c + 3]));
            }
          } else {
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
              if (c < p.Cout) atomicAdd(reinterpret_cast<float4*>(dst + c), make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
              if (c + 16 < p.Cout)
                atomicAdd(reinterpret_cast<float4*>(dst + 16 + c), make_float4(u[c], u[c + 1], u[c + 2], u[c + 3]));
            }
          }
        }
      }
      tc::fence_before_sync();
      tc::mbar_arrive(&B->acc_empty);
      ++nflush;
    }
    if (p.dbg && tid == 0) {
      g_wl_dbg[blockIdx.x * 8 + 6] = clock64() - t_begin;
      g_wl_dbg[blockIdx.x * 8 + 7] = tw_e;
    }
  } else if (warp < 8) {
    // ============================ PRODUCERS
    // Work = a stream of line jobs (x lines of the window, then the dy line of the step).  The global loads of job
    // k+1 are issued before job k is split / stored, so a full load latency is hidden behind every job.
    const int pt = tid - 128;
    constexpr int XC4 = 8 * CB;                 // float4 columns of an x voxel (32 channels per block)
    constexpr int YC4 = CB == 1 ? 4 : 8;        // float4 columns of a dy voxel
    constexpr int XIT = W * XC4 / 128, YIT = W * YC4 / 128;     // items per thread
    static_assert(XIT <= 4 && YIT <= 4 && XIT >= 1 && YIT >= 1, "line does not fit the register buffer");
    struct Job { int kind, n, zi, yi, ok; };    // kind 0: x line, 1: dy line
    int jt = t0, jj = 0;                        // job iterator: step, index inside the step
    auto next_job = [&](Job& jb) -> bool {
      if (jt >= t1) return false;
      const Step s = decode_step<CB>(p, jt, t0);
      const int nload = s.fresh ? s.nk : 1;
      if (jj < nload) {
        const int ky = s.fresh ? s.klo + jj : s.klo + s.nk - 1;
        jb.kind = 0; jb.n = s.n; jb.zi = s.z + s.kz - 2; jb.yi = s.y + ky - 2;
        jb.ok = (unsigned)jb.yi < (unsigned)p.H;             // zi is valid by construction of the pass
        ++jj;
      } else {
        jb.kind = 1; jb.n = s.n; jb.zi = s.z; jb.yi = s.y; jb.ok = 1;
        jj = 0; ++jt;
      }
      return true;
    };
    auto issue = [&](const Job& jb, float4 (&buf)[4]) {
      if (jb.kind == 0) {
        const float* src = p.x + ((((long long)jb.n * p.D + jb.zi) * p.H + jb.yi) * p.W) * p.x_cs + p.x_co;
#pragma unroll
        for (int u = 0; u < XIT; ++u) {
          const int it = pt + u * 128, v = it / XC4, c4 = it % XC4;
          buf[u] = (jb.ok && c4 * 4 < p.Cin) ? __ldg(reinterpret_cast<const float4*>(src + (long long)v * p.x_cs + c4 * 4))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        const float* src = p.dy + ((((long long)jb.n * p.D + jb.zi) * p.H + jb.yi) * p.W) * p.y_cs + p.y_co;
#pragma unroll
        for (int u = 0; u < YIT; ++u) {
          const int it = pt + u * 128, v = it / YC4, c4 = it % YC4;
          buf[u] = (c4 * 4 < p.Cout) ? __ldg(reinterpret_cast<const float4*>(src + (long long)v * p.y_cs + c4 * 4))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    uint32_t q = 0, yq = 0;         // x / dy line loads so far
    bool dead = false;
    long long tw_p = 0;
    const long long t_begin = p.dbg ? clock64() : 0;
    auto store = [&](const Job& jb, const float4 (&buf)[4]) {
      if (jb.kind == 0) {
        const int slot = (int)(q % XS);
        const uint32_t use = q / XS;
        if (use > 0 && !WL_TIMED(tw_p, tc::mbar_wait(&B->empty_x[slot], (use - 1) & 1, ab))) { fail(); dead = true; return; }
        uint8_t* base = xring + slot * XL;
#pragma unroll
        for (int u = 0; u < XIT; ++u) {
          const int it = pt + u * 128, v = it / XC4, c4 = it % XC4;
          float4 hi, lo;
          tc::split_tf32(buf[u].x, hi.x, lo.x); tc::split_tf32(buf[u].y, hi.y, lo.y);
          tc::split_tf32(buf[u].z, hi.z, lo.z); tc::split_tf32(buf[u].w, hi.w, lo.w);
          const int row = v + 4, cb = c4 >> 3, cw = c4 & 7;
          const uint32_t off = (uint32_t)row * 128 + (uint32_t)((((cw >> 1) ^ (row & 3)) << 5) + ((cw & 1) << 4));
          // CB = 1: blocks [hi][lo];  CB = 2: blocks [hi cb0][hi cb1][lo cb0][lo cb1]
          uint8_t* dh = base + (CB == 1 ? 0 : cb * XH) + off;
          uint8_t* dl = base + (CB == 1 ? XH : (2 + cb) * XH) + off;
          *reinterpret_cast<float4*>(dh) = hi;
          *reinterpret_cast<float4*>(dl) = lo;
          if (CB == 1 && slot == 0) {                          // mirror of slot 0 behind the last slot (row pairs)
            *reinterpret_cast<float4*>(dh + XS * XL) = hi;
            *reinterpret_cast<float4*>(dl + XS * XL) = lo;
          }
        }
        tc::fence_async_smem();
        tc::mbar_arrive(&B->full_x[slot]);
        ++q;
      } else {
        const int slot = (int)(yq % YSLOTS);
        const uint32_t use = yq / YSLOTS;
        if (use > 0 && !WL_TIMED(tw_p, tc::mbar_wait(&B->empty_y[slot], (use - 1) & 1, ab))) { fail(); dead = true; return; }
        uint8_t* base = yring + slot * YL;
#pragma unroll
        for (int u = 0; u < YIT; ++u) {
          const int it = pt + u * 128, v = it / YC4, c4 = it % YC4;
          float4 hi, lo;
          tc::split_tf32(buf[u].x, hi.x, lo.x); tc::split_tf32(buf[u].y, hi.y, lo.y);
          tc::split_tf32(buf[u].z, hi.z, lo.z); tc::split_tf32(buf[u].w, hi.w, lo.w);
          const int row = v + 6;
          if (CB == 1) {                                       // [hi16 | lo16] in one 128-byte row
            const uint32_t oh = (uint32_t)row * 128 + (uint32_t)((((c4 >> 1) ^ (row & 3)) << 5) + ((c4 & 1) << 4));
            const uint32_t ol = (uint32_t)row * 128 + (uint32_t)((((2 + (c4 >> 1)) ^ (row & 3)) << 5) + ((c4 & 1) << 4));
            *reinterpret_cast<float4*>(base + oh) = hi;
            *reinterpret_cast<float4*>(base + ol) = lo;
          } else {                                             // hi line, lo line
            const uint32_t off = (uint32_t)row * 128 + (uint32_t)((((c4 >> 1) ^ (row & 3)) << 5) + ((c4 & 1) << 4));
            *reinterpret_cast<float4*>(base + off) = hi;
            *reinterpret_cast<float4*>(base + YROWS * 128 + off) = lo;
          }
        }
        tc::fence_async_smem();
        tc::mbar_arrive(&B->full_y[slot]);
        ++yq;
      }
    };
    constexpr int PD = 4;             // jobs in flight (load latency under HBM streaming ~ 2 steps)
    Job jobs[PD];
    float4 bufs[PD][4];
    bool have[PD];
#pragma unroll
    for (int k = 0; k < PD; ++k) {
      have[k] = next_job(jobs[k]);
      if (have[k]) issue(jobs[k], bufs[k]);
    }
    while (have[0] && !dead) {
#pragma unroll
      for (int k = 0; k < PD; ++k) {
        if (!have[k] || dead) { have[0] = false; break; }
        store(jobs[k], bufs[k]);
        have[k] = next_job(jobs[k]);
        if (have[k]) issue(jobs[k], bufs[k]);
      }
    }
    if (p.dbg && pt == 0) {
      g_wl_dbg[blockIdx.x * 8 + 4] = clock64() - t_begin;
      g_wl_dbg[blockIdx.x * 8 + 5] = tw_p;
    }
  } else {
    // ============================ MMA ISSUER (one elected thread)
    // This single thread is the critical path (its instruction latency, not the tensor pipe, bounded the first
    // version): no divisions, no per-step decode -- the step state is carried incrementally, descriptors are built
    // once per step and advanced by adding the run offset to the 14-bit start-address field.
    {  // converged warp, elected lane issues (tc_common.cuh elect_one)
      const int p_single = p.single;
      constexpr uint32_t idesc = tc::make_idesc_tf32(128, ACOLS, 1, 1);
      uint32_t q = 0;               // x line loads consumed so far (mirrors the producers' counter)
      uint32_t nflush = 0;
      bool first = true;            // the next MMA of every accumulator overwrites (start of a flush interval)
      bool dead = false;
      const Step s0 = decode_step<CB>(p, t0, t0);
      int pass = s0.pass, y = s0.y;
      int pass_end = p.pass_begin[pass + 1];
      int mod_ctr = 0;                           // (t - t0 + 1) % flush_every, the epilogue's periodic flush rule
      uint32_t xslot = 0, xphase = 0;            // ring position / phase of the next x line load to consume
      uint32_t yslot = 0, yphase = 0;
      long long tw_x = 0, tw_y = 0, tw_a = 0;
      const long long t_begin = p.dbg ? clock64() : 0;
      for (int t = t0; t < t1 && !dead; ++t) {
        const int nk = CB == 1 ? 5 : ((pass & 1) ? 2 : 3);
        const bool fresh = (t == t0) || y == 0;
        const int nload = fresh ? nk : 1;
        for (int j = 0; j < nload; ++j) {
          if (!WL_TIMED(tw_x, tc::mbar_wait(&B->full_x[xslot], xphase, ab))) { fail(); dead = true; break; }
          if (++xslot == XS) { xslot = 0; xphase ^= 1; }
        }
        if (dead) break;
        q += nload;
        if (!WL_TIMED(tw_y, tc::mbar_wait(&B->full_y[yslot], yphase, ab))) { fail(); dead = true; break; }
        if (first && nflush > 0) {              // accumulators were handed to the epilogue: wait until it has read them
          if (!WL_TIMED(tw_a, tc::mbar_wait(&B->acc_empty, (nflush - 1) & 1, ab))) { fail(); dead = true; break; }
        }
        tc::fence_after_sync();
        const uint32_t ybase = yring_u32 + yslot * YL;
        // window line 0 sits nk slots behind the ring head
        uint32_t w0slot = xslot + XS - nk; if (w0slot >= XS) w0slot -= XS;
        uint64_t da[NACC];
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          uint32_t sl = w0slot + (CB == 1 ? 2 * a : a); if (sl >= XS) sl -= XS;
          // rows outside the image are staged as zero lines (no skipping: every accumulator of the pass is
          // (re)initialised by the first run after a flush, which the epilogue relies on)
          da[a] = tc::make_desc_mn32(xring_u32 + sl * XL, XH, 512);
        }
        const uint64_t dbh = tc::make_desc_mn32(ybase, 128, 512);
        const uint64_t dbl = tc::make_desc_mn32(ybase + YROWS * 128, 128, 512);
#pragma unroll
        for (int r = 0; r < RUNS; ++r) {
          const uint32_t acc = (first && r == 0) ? 0u : 1u;
          const uint64_t ro = (uint64_t)(r * 64);            // 1024 bytes >> 4
#pragma unroll
          for (int a = 0; a < NACC; ++a) {
            if (CB == 1) {
              tc::mma_tf32_e(tmem + a * ACOLS, da[a] + ro, dbh + ro, idesc, acc);
            } else if (a < nk) {
              tc::mma_tf32_e(tmem + a * ACOLS, da[a] + ro, dbh + ro, idesc, acc);
              if (!p_single) tc::mma_tf32_e(tmem + a * ACOLS, da[a] + ro, dbl + ro, idesc, 1u);
            }
          }
        }
        first = false;
        tc::commit_e(&B->empty_y[yslot]);
        if (++yslot == YSLOTS) { yslot = 0; yphase ^= 1; }
        // release the x lines the next step drops from the window
        const bool next_fresh = (y + 1 == p.H);
        if (t + 1 < t1) {
          const int ndrop = next_fresh ? nk : 1;
          uint32_t sl = w0slot;
          for (int j = 0; j < ndrop; ++j) {
            tc::commit_e(&B->empty_x[sl]);
            if (++sl == XS) sl = 0;
          }
        }
        if (++mod_ctr == p.flush_every) mod_ctr = 0;
        if (t == t1 - 1 || t + 1 >= pass_end || mod_ctr == 0) {
          tc::commit_e(&B->acc_full);
          ++nflush;
          first = true;
        }
        if (++y == p.H) y = 0;
        if (t + 1 >= pass_end && t + 1 < t1) { ++pass; pass_end = p.pass_begin[pass + 1]; }
      }
      if (p.dbg && lane == 0) {
        g_wl_dbg[blockIdx.x * 8 + 0] = clock64() - t_begin;
        g_wl_dbg[blockIdx.x * 8 + 1] = tw_x;
        g_wl_dbg[blockIdx.x * 8 + 2] = tw_y;
        g_wl_dbg[blockIdx.x * 8 + 3] = tw_a;
      }
    }
  }
  // ---- teardown
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

namespace {
template <int CB, int W>
int launch_wl(WLParams p, cudaStream_t st) {
  constexpr int RS = ((W + 8) / 8) * 8;
  constexpr int XL = 2 * CB * RS * 128, YL = CB * (RS + 8) * 128;
  constexpr int XALL = WLCfg<CB>::XSLOTS + (CB == 1 ? 1 : 0);
  const size_t smem = (size_t)XALL * XL + (size_t)YSLOTS * YL + sizeof(WLBarriers) + 1024 + 64;
  auto kern = wgrad_line_kernel<CB, W>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      crn_set_error("conv_wgrad_line: cannot set %zu bytes of dynamic shared memory", smem);
      return CRN_ERR_LAUNCH;
    }
    configured = true;
  }
  p.npass = CB == 1 ? 5 : 10;
  int tot = 0;
  for (int ps = 0; ps < p.npass; ++ps) {
    const int kz = CB == 1 ? ps : ps >> 1;
    p.pass_begin[ps] = tot;
    tot += p.N * (p.D - (kz < 2 ? 2 - kz : kz - 2)) * p.H;
  }
  p.pass_begin[p.npass] = tot;
  p.flush_every = 64;
  const int grid = tot < kNumSMs ? tot : kNumSMs;
  kern<<<grid, NTHREADS, smem, st>>>(p);
  CRN_LAUNCH_CHECK("conv_wgrad_line");
  return CRN_OK;
}
}  // namespace

extern "C" int crn_conv_wgrad_line_supported(const crn_conv_desc* d) {
  if (!d || d->transposed || d->kD != 5 || d->kH != 5 || d->kW != 5 || d->stride != 1 || d->pad != 2) return 0;
  if (d->iD != d->oD || d->iH != d->oH || d->iW != d->oW || d->y_planar) return 0;
  if (d->Cin % 4 || d->Cout % 4 || d->x_cs % 4 || d->x_co % 4 || d->y_cs % 4 || d->y_co % 4 || d->CoutP % 4) return 0;
  if (d->iD < 3) return 0;
  if (d->Cin <= 32 && d->Cout <= 16 && (d->iW == 64 || d->iW == 32)) return 1;
  if (d->Cin <= 64 && d->Cout <= 32 && d->iW == 32) return 1;
  return 0;
}

// dWf[tap][ci][co] += sum_vox x * dy for Conv3d k=5 s=1 p=2 with narrow channels (same contract as crn_conv_wgrad).
extern "C" int crn_conv_wgrad_line(const crn_conv_desc* d, const float* x, const float* dy, float* dw_packed,
                                   int32_t* status, void* stream) {
  CRN_REQUIRE(d && x && dy && dw_packed && status, "crn_conv_wgrad_line: null pointer");
  CRN_REQUIRE(crn_conv_wgrad_line_supported(d), "crn_conv_wgrad_line: unsupported layer shape");
  WLParams p{};
  p.single = crn_single_pass();
  p.dbg = (crn_get_flags() >> 8) & 1;
  p.x = x; p.dy = dy; p.dw = dw_packed; p.status = status;
  p.N = d->N; p.D = d->iD; p.H = d->iH; p.W = d->iW;
  p.Cin = d->Cin; p.Cout = d->Cout;
  p.x_cs = d->x_cs; p.x_co = d->x_co; p.y_cs = d->y_cs; p.y_co = d->y_co; p.CinP = d->CinP; p.CoutP = d->CoutP;
  cudaStream_t st = crn_stream(stream);
  if (d->Cin <= 32 && d->Cout <= 16) return d->iW == 64 ? launch_wl<1, 64>(p, st) : launch_wl<1, 32>(p, st);
  return launch_wl<2, 32>(p, st);
}

// =============================================================================================================
// ConvTranspose3d k=7 s=2 p=3 (output_padding 1) weight gradient for Cin <= 32, Cout == 16 (stage_5.t1,
// model/reconstruction_decoder.py:85):   dW[k][ci][co] = sum_i x[i][ci] * dy[2i - 3 + k][co].
// With o = 2m + c (c = parity class) and j = m - i the filter index is k = 2j + c + 3 per axis, so on the COARSE
// grid this is a stride-1 weight gradient with 4 shifts j in {-2..1} per axis between x and the class-channel view
// dyc[m][(cz, cy, cx, co)] = dy[2m + c][co].  One fine voxel PAIR (cx = 0, 1) x 16 co is exactly one 128-byte
// MN-major row, so block (cz, cy) of a coarse line is simply the fine line (2mz + cz, 2my + cy):
//   * step = coarse dy line (mz, my): 4 fine lines staged as 4 N-blocks (N = 128), hi and lo copies (2 MMAs);
//   * M = 128 = {hi, lo} x {x line y_i, y_i + 1} x 32 ci exactly as in the k=5 kernel (ring + mirror slot);
//   * accumulator a <-> jx = a - 2 (dyc start address shifted by a rows), 4 x 128 TMEM columns;
//   * pass = (jz, jy pair); all four hi/lo products are formed, taps with k outside 0..6 are dropped in the epilogue.
namespace {

struct TStep {
  int pass, jz, g;         // jz in -2..1, g: jy pair (g = 0: jy = 1, 0; g = 1: jy = -1, -2)
  int n, mz, my;
  bool fresh;
};

__device__ __forceinline__ TStep decode_tstep(const WLParams& p, int t, int t0) {
  TStep s;
  int ps = 0;
  while (ps + 1 < p.npass && t >= p.pass_begin[ps + 1]) ++ps;
  s.pass = ps;
  s.jz = (ps >> 1) - 2; s.g = ps & 1;
  const int r = t - p.pass_begin[ps];
  s.my = r % p.H;
  const int pl = r / p.H;
  const int zlo = s.jz > 0 ? s.jz : 0, zcnt = p.D - (s.jz < 0 ? -s.jz : s.jz);
  s.n = pl / zcnt;
  s.mz = zlo + pl % zcnt;
  s.fresh = (t == t0) || s.my == 0;
  return s;
}

template <int W>
__global__ void __launch_bounds__(NTHREADS, 1) wgrad_tline_kernel(const WLParams p) {
  constexpr int RUNS = (W + 8) / 8;
  constexpr int RS = RUNS * 8;                      // x line rows (coarse voxel i = row - 4)
  constexpr int XH = RS * 128;
  constexpr int XL = 2 * XH;                        // [hi][lo]
  constexpr int YROWS = RS + 8;                     // dyc rows (coarse voxel m = row - 6)
  constexpr int YB = YROWS * 128;                   // one (cz, cy) block line (LBO of the N operand)
  constexpr int YL = 8 * YB;                        // [h][b = cz*2 + cy]
  constexpr int XS = 6, XALL = 7, YS = 2;
  constexpr int TCOLS = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* xring = smem;
  uint8_t* yring = smem + XALL * XL;
  WLBarriers* B = reinterpret_cast<WLBarriers*>(yring + YS * YL);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)tc::uniform_u32((uint32_t)tid >> 5);   // warp-uniform for the compiler (role branches)
  const int T = p.pass_begin[p.npass];
  const int t0 = (int)((long long)T * blockIdx.x / gridDim.x), t1 = (int)((long long)T * (blockIdx.x + 1) / gridDim.x);

  if (tid == 0) {
    for (int i = 0; i < MAXX; ++i) { tc::mbar_init(&B->full_x[i], 128); tc::mbar_init(&B->empty_x[i], 1); }
    for (int i = 0; i < YSLOTS; ++i) { tc::mbar_init(&B->full_y[i], 256); tc::mbar_init(&B->empty_y[i], 1); }
    tc::mbar_init(&B->acc_full, 1); tc::mbar_init(&B->acc_empty, 128);
    B->abort_flag = 0;
    tc::mbar_fence_init();
  }
  if (warp == 8) tc::tmem_alloc(&B->tmem_base, 512);
  for (int i = tid; i < (XALL * XL + YS * YL) / 16; i += NTHREADS)
    reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tc::uniform_u32(B->tmem_base);
  const uint32_t xring_u32 = tc::smem_u32(xring), yring_u32 = tc::smem_u32(yring);
  auto fail = [&]() { B->abort_flag = 1; *p.status = 1; };
  volatile int* ab = &B->abort_flag;
  auto flush_after = [&](int t, const TStep& s) -> bool {
    if (t == t1 - 1) return true;
    if (t + 1 >= p.pass_begin[s.pass + 1]) return true;
    return (t - t0 + 1) % p.flush_every == 0;
  };

  if (warp < 4) {
    // ============================ EPILOGUE
    uint32_t nflush = 0;
    bool dead = false;
    const int ya = warp >> 1;                                  // blocks: (hi, y_i), (lo, y_i), (hi, y_i+1), (lo, y_i+1)
    const int ci = lane;
    for (int t = t0; t < t1 && !dead; ++t) {
      const TStep s = decode_tstep(p, t, t0);
      if (!flush_after(t, s)) continue;
      if (!tc::mbar_wait(&B->acc_full, nflush & 1, ab)) { fail(); dead = true; break; }
      tc::fence_after_sync();
      const int jy = 1 - 2 * s.g - ya;
#pragma unroll 1
      for (int a = 0; a < 4; ++a) {
        const int jx = a - 2;
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {                          // N block = (cz, cy); its 32 columns = (cx, co)
          const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + a * TCOLS + b * 32;
          float v[16], u[16];
          tc::tmem_ld16(ta, v);                                // cx = 0
          tc::tmem_ld16(ta + 16, u);                           // cx = 1
          const int kz = 2 * s.jz + (b >> 1) + 3, ky = 2 * jy + (b & 1) + 3;
          if ((unsigned)kz > 6u || (unsigned)ky > 6u || ci >= p.Cin) continue;
#pragma unroll
          for (int cx = 0; cx < 2; ++cx) {
            const int kx = 2 * jx + cx + 3;
            if ((unsigned)kx > 6u) continue;
            float* dst = p.dw + ((long long)((kz * 7 + ky) * 7 + kx) * p.CinP + ci) * p.CoutP;
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
              if (c >= p.Cout) break;
              atomicAdd(reinterpret_cast<float4*>(dst + c),
                        cx == 0 ? make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]) : make_float4(u[c], u[c + 1], u[c + 2], u[c + 3]));
            }
          }
        }
      }
      tc::fence_before_sync();
      tc::mbar_arrive(&B->acc_empty);
      ++nflush;
    }
  } else if (warp < 8) {
    // ============================ PRODUCERS (jobs: x line | fine lines cz = 0 | fine lines cz = 1)
    const int pt = tid - 128;
    constexpr int XIT = W * 8 / 128;                  // x items per thread (float4)
    constexpr int FW = 2 * W;                         // fine voxels per fine line
    constexpr int YIT = 2 * FW * 4 / 128;             // items per thread of one dy job (2 fine lines x FW x 4 float4)
    static_assert(XIT >= 1 && XIT <= 4 && YIT >= 1 && YIT <= 4, "line does not fit the register buffer");
    struct Job { int kind, n, z, y, ok; };            // kind 0: x line (z, y coarse); 1, 2: dy fine lines of cz = kind - 1
    int jt = t0, jj = 0;
    auto next_job = [&](Job& jb) -> bool {
      if (jt >= t1) return false;
      const TStep s = decode_tstep(p, jt, t0);
      const int nload = s.fresh ? 2 : 1;
      if (jj < nload) {
        const int j = s.fresh ? jj : 1;
        jb.kind = 0; jb.n = s.n; jb.z = s.mz - s.jz; jb.y = s.my - 1 + 2 * s.g + j;
        jb.ok = (unsigned)jb.y < (unsigned)p.H;
        ++jj;
      } else {
        const int half = jj - nload;
        jb.kind = 1 + half; jb.n = s.n; jb.z = s.mz; jb.y = s.my; jb.ok = 1;
        if (half == 1) { jj = 0; ++jt; } else ++jj;
      }
      return true;
    };
    auto issue = [&](const Job& jb, float4 (&buf)[4]) {
      if (jb.kind == 0) {
        const float* src = p.x + ((((long long)jb.n * p.D + jb.z) * p.H + jb.y) * p.W) * p.x_cs + p.x_co;
#pragma unroll
        for (int u = 0; u < XIT; ++u) {
          const int it = pt + u * 128, v = it >> 3, c4 = it & 7;
          buf[u] = (jb.ok && c4 * 4 < p.Cin) ? __ldg(reinterpret_cast<const float4*>(src + (long long)v * p.x_cs + c4 * 4))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        const int cz = jb.kind - 1;
#pragma unroll
        for (int u = 0; u < YIT; ++u) {
          const int it = pt + u * 128, c4 = it & 3, f = (it >> 2) % FW, cy = it / (4 * FW);
          const float* src = p.dy + ((((long long)jb.n * (2 * p.D) + 2 * jb.z + cz) * (2 * p.H) + 2 * jb.y + cy) * (2 * p.W) + f) *
                                        p.y_cs + p.y_co;
          buf[u] = (c4 * 4 < p.Cout) ? __ldg(reinterpret_cast<const float4*>(src + c4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    uint32_t q = 0, yq = 0;
    bool dead = false;
    auto store = [&](const Job& jb, const float4 (&buf)[4]) {
      if (jb.kind == 0) {
        const int slot = (int)(q % XS);
        const uint32_t use = q / XS;
        if (use > 0 && !tc::mbar_wait(&B->empty_x[slot], (use - 1) & 1, ab)) { fail(); dead = true; return; }
        uint8_t* base = xring + slot * XL;
#pragma unroll
        for (int u = 0; u < XIT; ++u) {
          const int it = pt + u * 128, v = it >> 3, cw = it & 7;
          float4 hi, lo;
          tc::split_tf32(buf[u].x, hi.x, lo.x); tc::split_tf32(buf[u].y, hi.y, lo.y);
          tc::split_tf32(buf[u].z, hi.z, lo.z); tc::split_tf32(buf[u].w, hi.w, lo.w);
          const int row = v + 4;
          const uint32_t off = (uint32_t)row * 128 + (uint32_t)((((cw >> 1) ^ (row & 3)) << 5) + ((cw & 1) << 4));
          *reinterpret_cast<float4*>(base + off) = hi;
          *reinterpret_cast<float4*>(base + XH + off) = lo;
          if (slot == 0) {
            *reinterpret_cast<float4*>(base + XS * XL + off) = hi;
            *reinterpret_cast<float4*>(base + XS * XL + XH + off) = lo;
          }
        }
        tc::fence_async_smem();
        tc::mbar_arrive(&B->full_x[slot]);
        ++q;
      } else {
        const int cz = jb.kind - 1;
        const int slot = (int)(yq % YS);
        const uint32_t use = yq / YS;
        if (cz == 0 && use > 0 && !tc::mbar_wait(&B->empty_y[slot], (use - 1) & 1, ab)) { fail(); dead = true; return; }
        uint8_t* base = yring + slot * YL;
#pragma unroll
        for (int u = 0; u < YIT; ++u) {
          const int it = pt + u * 128, c4 = it & 3, f = (it >> 2) % FW, cy = it / (4 * FW);
          float4 hi, lo;
          tc::split_tf32(buf[u].x, hi.x, lo.x); tc::split_tf32(buf[u].y, hi.y, lo.y);
          tc::split_tf32(buf[u].z, hi.z, lo.z); tc::split_tf32(buf[u].w, hi.w, lo.w);
          const int row = (f >> 1) + 6, cx = f & 1, b = cz * 2 + cy;
          const int chunk = cx * 2 + (c4 >> 1);
          const uint32_t off = (uint32_t)b * YB + (uint32_t)row * 128 + (uint32_t)(((chunk ^ (row & 3)) << 5) + ((c4 & 1) << 4));
          *reinterpret_cast<float4*>(base + off) = hi;
          *reinterpret_cast<float4*>(base + 4 * YB + off) = lo;
        }
        tc::fence_async_smem();
        tc::mbar_arrive(&B->full_y[slot]);       // 256 arrivals: both halves
        if (cz == 1) ++yq;
      }
    };
    constexpr int PD = 4;             // jobs in flight (load latency under HBM streaming ~ 2 steps)
    Job jobs[PD];
    float4 bufs[PD][4];
    bool have[PD];
#pragma unroll
    for (int k = 0; k < PD; ++k) {
      have[k] = next_job(jobs[k]);
      if (have[k]) issue(jobs[k], bufs[k]);
    }
    while (have[0] && !dead) {
#pragma unroll
      for (int k = 0; k < PD; ++k) {
        if (!have[k] || dead) { have[0] = false; break; }
        store(jobs[k], bufs[k]);
        have[k] = next_job(jobs[k]);
        if (have[k]) issue(jobs[k], bufs[k]);
      }
    }
  } else {
    // ============================ MMA ISSUER (incremental step state, see wgrad_line_kernel)
    {  // converged warp, elected lane issues (tc_common.cuh elect_one)
      const int p_single = p.single;
      constexpr uint32_t idesc = tc::make_idesc_tf32(128, 128, 1, 1);
      uint32_t nflush = 0;
      bool first = true;            // next MMA of every accumulator overwrites (start of a flush interval)
      bool dead = false;
      const TStep s0 = decode_tstep(p, t0, t0);
      int pass = s0.pass, y = s0.my, mod_ctr = 0;
      int pass_end = p.pass_begin[pass + 1];
      uint32_t xslot = 0, xphase = 0, yslot = 0, yphase = 0;
      for (int t = t0; t < t1 && !dead; ++t) {
        const bool fresh = (t == t0) || y == 0;
        const int nload = fresh ? 2 : 1;
        for (int j = 0; j < nload; ++j) {
          if (!tc::mbar_wait(&B->full_x[xslot], xphase, ab)) { fail(); dead = true; break; }
          if (++xslot == XS) { xslot = 0; xphase ^= 1; }
        }
        if (dead) break;
        if (!tc::mbar_wait(&B->full_y[yslot], yphase, ab)) { fail(); dead = true; break; }
        if (first && nflush > 0) {
          if (!tc::mbar_wait(&B->acc_empty, (nflush - 1) & 1, ab)) { fail(); dead = true; break; }
        }
        tc::fence_after_sync();
        const uint32_t ybase = yring_u32 + yslot * YL;
        uint32_t w0slot = xslot + XS - 2; if (w0slot >= XS) w0slot -= XS;   // window line 0 (line 1 follows; mirror)
        const uint64_t da0 = tc::make_desc_mn32(xring_u32 + w0slot * XL, XH, 512);
        const uint64_t dbh0 = tc::make_desc_mn32(ybase, YB, 512);
        const uint64_t dbl0 = tc::make_desc_mn32(ybase + 4 * YB, YB, 512);
#pragma unroll
        for (int r = 0; r < RUNS; ++r) {
          const uint32_t acc = (first && r == 0) ? 0u : 1u;
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const uint64_t bo = (uint64_t)((r * 8 + a) * 8);    // (r*8 + a) rows of 128 bytes, >> 4
            tc::mma_tf32_e(tmem + a * TCOLS, da0 + (uint64_t)(r * 64), dbh0 + bo, idesc, acc);
            if (!p_single) tc::mma_tf32_e(tmem + a * TCOLS, da0 + (uint64_t)(r * 64), dbl0 + bo, idesc, 1u);
          }
        }
        first = false;
        tc::commit_e(&B->empty_y[yslot]);
        if (++yslot == YS) { yslot = 0; yphase ^= 1; }
        const bool next_fresh = (y + 1 == p.H);
        if (t + 1 < t1) {
          const int ndrop = next_fresh ? 2 : 1;
          uint32_t sl = w0slot;
          for (int j = 0; j < ndrop; ++j) {
            tc::commit_e(&B->empty_x[sl]);
            if (++sl == XS) sl = 0;
          }
        }
        if (++mod_ctr == p.flush_every) mod_ctr = 0;
        if (t == t1 - 1 || t + 1 >= pass_end || mod_ctr == 0) {
          tc::commit_e(&B->acc_full);
          ++nflush;
          first = true;
        }
        if (++y == p.H) y = 0;
        if (t + 1 >= pass_end && t + 1 < t1) { ++pass; pass_end = p.pass_begin[pass + 1]; }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

template <int W>
int launch_wtl(WLParams p, cudaStream_t st) {
  constexpr int RS = ((W + 8) / 8) * 8;
  constexpr int XL = 2 * RS * 128, YL = 8 * (RS + 8) * 128;
  const size_t smem = (size_t)7 * XL + (size_t)2 * YL + sizeof(WLBarriers) + 1024 + 64;
  auto kern = wgrad_tline_kernel<W>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      crn_set_error("conv_wgrad_tline: cannot set %zu bytes of dynamic shared memory", smem);
      return CRN_ERR_LAUNCH;
    }
    configured = true;
  }
  p.npass = 8;
  int tot = 0;
  for (int ps = 0; ps < 8; ++ps) {
    const int jz = (ps >> 1) - 2;
    p.pass_begin[ps] = tot;
    tot += p.N * (p.D - (jz < 0 ? -jz : jz)) * p.H;
  }
  p.pass_begin[8] = tot;
  p.flush_every = 64;
  const int grid = tot < kNumSMs ? tot : kNumSMs;
  kern<<<grid, NTHREADS, smem, st>>>(p);
  CRN_LAUNCH_CHECK("conv_wgrad_tline");
  return CRN_OK;
}
}  // namespace

// =============================================================================================================
// Same operator for the LAST decoder layer (stage_6.t1: Cin <= 16, Cout <= 4 stored with channel stride 4, i.e. the
// 2-class logits gradient; model/reconstruction_decoder.py:95).  Everything is narrower, so more taps are stacked:
//   * x voxel row = [hi16 | lo16] (one 128-byte row), M = 128 = FOUR consecutive x lines y_i = my-1 .. my+2, i.e.
//     all four jy shifts in one MMA (ring of 8 lines + 3 mirror slots keeps any 4 consecutive lines contiguous);
//   * dyc voxel row = 8 classes x 4 co = 32 floats = one 128-byte row gathered from the 4 fine lines (cz, cy)
//     (32-byte chunk b = cz*2 + cy holds the fine voxel pair cx = 0, 1), hi line and lo line; N = 128 = the 4 jx
//     shifts stacked through LBO = 128 B;
//   * ONE accumulator (128 columns) per pass jz: 2 MMAs per 8 coarse voxels cover 16 shifts x 8 classes.
namespace {

template <int W>
__global__ void __launch_bounds__(NTHREADS, 1) wgrad_t2line_kernel(const WLParams p) {
  constexpr int RUNS = (W + 8) / 8;
  constexpr int RS = RUNS * 8;
  constexpr int XL = RS * 128;                      // one x line ([hi16|lo16] rows) = LBO of the M operand
  constexpr int YROWS = RS + 8;
  constexpr int YH = YROWS * 128;                   // hi (or lo) dyc line
  constexpr int YL = 2 * YH;
  constexpr int XS = 8, XALL = 11, YS = 3;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* xring = smem;
  uint8_t* yring = smem + XALL * XL;
  WLBarriers* B = reinterpret_cast<WLBarriers*>(yring + YS * YL);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)tc::uniform_u32((uint32_t)tid >> 5);   // warp-uniform for the compiler (role branches)
  const int T = p.pass_begin[p.npass];
  const int t0 = (int)((long long)T * blockIdx.x / gridDim.x), t1 = (int)((long long)T * (blockIdx.x + 1) / gridDim.x);

  if (tid == 0) {
    for (int i = 0; i < MAXX; ++i) { tc::mbar_init(&B->full_x[i], 128); tc::mbar_init(&B->empty_x[i], 1); }
    for (int i = 0; i < YSLOTS; ++i) { tc::mbar_init(&B->full_y[i], 128); tc::mbar_init(&B->empty_y[i], 1); }
    tc::mbar_init(&B->acc_full, 1); tc::mbar_init(&B->acc_empty, 128);
    B->abort_flag = 0;
    tc::mbar_fence_init();
  }
  if (warp == 8) tc::tmem_alloc(&B->tmem_base, 128);
  for (int i = tid; i < (XALL * XL + YS * YL) / 16; i += NTHREADS)
    reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tc::uniform_u32(B->tmem_base);
  const uint32_t xring_u32 = tc::smem_u32(xring), yring_u32 = tc::smem_u32(yring);
  auto fail = [&]() { B->abort_flag = 1; *p.status = 1; };
  volatile int* ab = &B->abort_flag;
  // pass = jz (npass = 4): decode_tstep's (jz, g) split is not used here
  auto decode = [&](int t) -> TStep {
    TStep s;
    int ps = 0;
    while (ps + 1 < p.npass && t >= p.pass_begin[ps + 1]) ++ps;
    s.pass = ps; s.jz = ps - 2; s.g = 0;
    const int r = t - p.pass_begin[ps];
    s.my = r % p.H;
    const int pl = r / p.H;
    const int zlo = s.jz > 0 ? s.jz : 0, zcnt = p.D - (s.jz < 0 ? -s.jz : s.jz);
    s.n = pl / zcnt; s.mz = zlo + pl % zcnt;
    s.fresh = (t == t0) || s.my == 0;
    return s;
  };
  auto flush_after = [&](int t, const TStep& s) -> bool {
    if (t == t1 - 1) return true;
    if (t + 1 >= p.pass_begin[s.pass + 1]) return true;
    return (t - t0 + 1) % p.flush_every == 0;
  };

  if (warp < 4) {
    // ============================ EPILOGUE
    uint32_t nflush = 0;
    bool dead = false;
    const int jy = 1 - warp;                                   // M block = x line y_i = my - 1 + warp
    const int ci = lane & 15;                                  // lanes 0-15: hi, 16-31: lo (both add)
    for (int t = t0; t < t1 && !dead; ++t) {
      const TStep s = decode(t);
      if (!flush_after(t, s)) continue;
      if (!tc::mbar_wait(&B->acc_full, nflush & 1, ab)) { fail(); dead = true; break; }
      tc::fence_after_sync();
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        if (ci >= p.Cin) continue;
        const int jx = (c0 >> 5) - 2;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {                        // 4-column group = (class b, cx): the 4 co
          const int col = c0 + gq * 4;
          const int b = (col >> 3) & 3, cx = (col >> 2) & 1;
          const int kz = 2 * s.jz + (b >> 1) + 3, ky = 2 * jy + (b & 1) + 3, kx = 2 * jx + cx + 3;
          if ((unsigned)kz > 6u || (unsigned)ky > 6u || (unsigned)kx > 6u) continue;
          float* dst = p.dw + ((long long)((kz * 7 + ky) * 7 + kx) * p.CinP + ci) * p.CoutP;
          atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[gq * 4], v[gq * 4 + 1], v[gq * 4 + 2], v[gq * 4 + 3]));
        }
      }
      tc::fence_before_sync();
      tc::mbar_arrive(&B->acc_empty);
      ++nflush;
    }
  } else if (warp < 8) {
    // ============================ PRODUCERS (jobs: x line only | x line + dyc line of the step)
    // The MMA work per step is short here (18 MMAs), so a step's loads are ONE job (one load latency per step, issued
    // a full step ahead); only the 3 extra x lines of a fresh window are separate jobs.
    const int pt = tid - 128;
    constexpr int XIT = W * 4 / 128;                  // x: W voxels x 4 float4
    constexpr int FW = 2 * W;
    constexpr int YIT = 4 * FW / 128;                 // dy: 4 fine lines x FW voxels x 1 float4
    constexpr int NB = XIT + YIT;
    static_assert(XIT >= 1 && YIT >= 1 && NB <= 6, "line does not fit the register buffer");
    struct Job { int with_y, n, z, y, ok, yz, yy; };
    int jt = t0, jj = 0;
    auto next_job = [&](Job& jb) -> bool {
      if (jt >= t1) return false;
      const TStep s = decode(jt);
      const int nload = s.fresh ? 4 : 1;
      const int j = s.fresh ? jj : 3;
      jb.n = s.n; jb.z = s.mz - s.jz; jb.y = s.my - 1 + j;
      jb.ok = (unsigned)jb.y < (unsigned)p.H;
      jb.yz = s.mz; jb.yy = s.my;
      if (jj + 1 < nload) { jb.with_y = 0; ++jj; }
      else { jb.with_y = 1; jj = 0; ++jt; }
      return true;
    };
    auto issue = [&](const Job& jb, float4 (&buf)[NB]) {
      const float* src = p.x + ((((long long)jb.n * p.D + jb.z) * p.H + jb.y) * p.W) * p.x_cs + p.x_co;
#pragma unroll
      for (int u = 0; u < XIT; ++u) {
        const int it = pt + u * 128, v = it >> 2, c4 = it & 3;
        buf[u] = (jb.ok && c4 * 4 < p.Cin) ? __ldg(reinterpret_cast<const float4*>(src + (long long)v * p.x_cs + c4 * 4))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (jb.with_y) {
#pragma unroll
        for (int u = 0; u < YIT; ++u) {
          const int it = pt + u * 128, f = it % FW, b = it / FW;
          const float* sy = p.dy + ((((long long)jb.n * (2 * p.D) + 2 * jb.yz + (b >> 1)) * (2 * p.H) + 2 * jb.yy + (b & 1)) *
                                        (2 * p.W) + f) * p.y_cs + p.y_co;
          buf[XIT + u] = __ldg(reinterpret_cast<const float4*>(sy));
        }
      }
    };
    uint32_t q = 0, yq = 0;
    bool dead = false;
    auto store = [&](const Job& jb, const float4 (&buf)[NB]) {
      {
        const int slot = (int)(q % XS);
        const uint32_t use = q / XS;
        if (use > 0 && !tc::mbar_wait(&B->empty_x[slot], (use - 1) & 1, ab)) { fail(); dead = true; return; }
        uint8_t* base = xring + slot * XL;
#pragma unroll
        for (int u = 0; u < XIT; ++u) {
          const int it = pt + u * 128, v = it >> 2, c4 = it & 3;
          float4 hi, lo;
          tc::split_tf32(buf[u].x, hi.x, lo.x); tc::split_tf32(buf[u].y, hi.y, lo.y);
          tc::split_tf32(buf[u].z, hi.z, lo.z); tc::split_tf32(buf[u].w, hi.w, lo.w);
          const int row = v + 4;
          const uint32_t oh = (uint32_t)row * 128 + (uint32_t)((((c4 >> 1) ^ (row & 3)) << 5) + ((c4 & 1) << 4));
          const uint32_t ol = (uint32_t)row * 128 + (uint32_t)((((2 + (c4 >> 1)) ^ (row & 3)) << 5) + ((c4 & 1) << 4));
          *reinterpret_cast<float4*>(base + oh) = hi;
          *reinterpret_cast<float4*>(base + ol) = lo;
          if (slot < 3) {                                      // mirrors of slots 0..2 behind the last slot
            *reinterpret_cast<float4*>(base + XS * XL + oh) = hi;
            *reinterpret_cast<float4*>(base + XS * XL + ol) = lo;
          }
        }
        tc::fence_async_smem();
        tc::mbar_arrive(&B->full_x[slot]);
        ++q;
      }
      if (jb.with_y) {
        const int slot = (int)(yq % YS);
        const uint32_t use = yq / YS;
        if (use > 0 && !tc::mbar_wait(&B->empty_y[slot], (use - 1) & 1, ab)) { fail(); dead = true; return; }
        uint8_t* base = yring + slot * YL;
#pragma unroll
        for (int u = 0; u < YIT; ++u) {
          const int it = pt + u * 128, f = it % FW, b = it / FW;
          float4 hi, lo;
          tc::split_tf32(buf[XIT + u].x, hi.x, lo.x); tc::split_tf32(buf[XIT + u].y, hi.y, lo.y);
          tc::split_tf32(buf[XIT + u].z, hi.z, lo.z); tc::split_tf32(buf[XIT + u].w, hi.w, lo.w);
          const int row = (f >> 1) + 6, cx = f & 1;
          const uint32_t off = (uint32_t)row * 128 + (uint32_t)(((b ^ (row & 3)) << 5) + (cx << 4));
          *reinterpret_cast<float4*>(base + off) = hi;
          *reinterpret_cast<float4*>(base + YH + off) = lo;
        }
        tc::fence_async_smem();
        tc::mbar_arrive(&B->full_y[slot]);
        ++yq;
      }
    };
    // PD jobs in flight: the loads of a job are issued PD - 1 jobs (~ steps) before they are needed, which covers the
    // DRAM latency under load (the lines stream from HBM: 4 passes over 200 MB do not stay in L2)
    constexpr int PD = 4;
    Job jobs[PD];
    float4 bufs[PD][NB];
    bool have[PD];
#pragma unroll
    for (int k = 0; k < PD; ++k) {
      have[k] = next_job(jobs[k]);
      if (have[k]) issue(jobs[k], bufs[k]);
    }
    while (have[0] && !dead) {
#pragma unroll
      for (int k = 0; k < PD; ++k) {
        if (!have[k] || dead) { have[0] = false; break; }
        store(jobs[k], bufs[k]);
        have[k] = next_job(jobs[k]);
        if (have[k]) issue(jobs[k], bufs[k]);
      }
    }
  } else {
    // ============================ MMA ISSUER (incremental step state, see wgrad_line_kernel)
    {  // converged warp, elected lane issues (tc_common.cuh elect_one)
      const int p_single = p.single;
      constexpr uint32_t idesc = tc::make_idesc_tf32(128, 128, 1, 1);
      uint32_t nflush = 0;
      bool first = true;
      bool dead = false;
      const TStep s0 = decode(t0);
      int pass = s0.pass, y = s0.my, mod_ctr = 0;
      int pass_end = p.pass_begin[pass + 1];
      uint32_t xslot = 0, xphase = 0, yslot = 0, yphase = 0;
      for (int t = t0; t < t1 && !dead; ++t) {
        const bool fresh = (t == t0) || y == 0;
        const int nload = fresh ? 4 : 1;
        for (int j = 0; j < nload; ++j) {
          if (!tc::mbar_wait(&B->full_x[xslot], xphase, ab)) { fail(); dead = true; break; }
          if (++xslot == XS) { xslot = 0; xphase ^= 1; }
        }
        if (dead) break;
        if (!tc::mbar_wait(&B->full_y[yslot], yphase, ab)) { fail(); dead = true; break; }
        if (first && nflush > 0) {
          if (!tc::mbar_wait(&B->acc_empty, (nflush - 1) & 1, ab)) { fail(); dead = true; break; }
        }
        tc::fence_after_sync();
        const uint32_t ybase = yring_u32 + yslot * YL;
        uint32_t w0slot = xslot + XS - 4; if (w0slot >= XS) w0slot -= XS;   // window line 0; 1..3 follow (mirrors)
        const uint64_t da0 = tc::make_desc_mn32(xring_u32 + w0slot * XL, XL, 512);
        const uint64_t dbh0 = tc::make_desc_mn32(ybase, 128, 512);
        const uint64_t dbl0 = tc::make_desc_mn32(ybase + YH, 128, 512);
#pragma unroll
        for (int r = 0; r < RUNS; ++r) {
          tc::mma_tf32_e(tmem, da0 + (uint64_t)(r * 64), dbh0 + (uint64_t)(r * 64), idesc, (first && r == 0) ? 0u : 1u);
          if (!p_single) tc::mma_tf32_e(tmem, da0 + (uint64_t)(r * 64), dbl0 + (uint64_t)(r * 64), idesc, 1u);
        }
        first = false;
        tc::commit_e(&B->empty_y[yslot]);
        if (++yslot == YS) { yslot = 0; yphase ^= 1; }
        const bool next_fresh = (y + 1 == p.H);
        if (t + 1 < t1) {
          const int ndrop = next_fresh ? 4 : 1;
          uint32_t sl = w0slot;
          for (int j = 0; j < ndrop; ++j) {
            tc::commit_e(&B->empty_x[sl]);
            if (++sl == XS) sl = 0;
          }
        }
        if (++mod_ctr == p.flush_every) mod_ctr = 0;
        if (t == t1 - 1 || t + 1 >= pass_end || mod_ctr == 0) {
          tc::commit_e(&B->acc_full);
          ++nflush;
          first = true;
        }
        if (++y == p.H) y = 0;
        if (t + 1 >= pass_end && t + 1 < t1) { ++pass; pass_end = p.pass_begin[pass + 1]; }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 128);
}

template <int W>
int launch_wt2l(WLParams p, cudaStream_t st) {
  constexpr int RS = ((W + 8) / 8) * 8;
  constexpr int XL = RS * 128, YL = 2 * (RS + 8) * 128;
  const size_t smem = (size_t)11 * XL + (size_t)3 * YL + sizeof(WLBarriers) + 1024 + 64;
  auto kern = wgrad_t2line_kernel<W>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      crn_set_error("conv_wgrad_t2line: cannot set %zu bytes of dynamic shared memory", smem);
      return CRN_ERR_LAUNCH;
    }
    configured = true;
  }
  p.npass = 4;
  int tot = 0;
  for (int ps = 0; ps < 4; ++ps) {
    const int jz = ps - 2;
    p.pass_begin[ps] = tot;
    tot += p.N * (p.D - (jz < 0 ? -jz : jz)) * p.H;
  }
  p.pass_begin[4] = tot;
  p.flush_every = 64;
  const int grid = tot < kNumSMs ? tot : kNumSMs;
  kern<<<grid, NTHREADS, smem, st>>>(p);
  CRN_LAUNCH_CHECK("conv_wgrad_t2line");
  return CRN_OK;
}
}  // namespace

extern "C" int crn_convt7_wgrad_line_supported(const crn_conv_desc* d) {
  if (!d || !d->transposed || d->kD != 7 || d->kH != 7 || d->kW != 7 || d->stride != 2 || d->pad != 3) return 0;
  if (d->oD != 2 * d->iD || d->oH != 2 * d->iH || d->oW != 2 * d->iW || d->y_planar) return 0;
  if (d->Cin % 4 || d->x_cs % 4 || d->x_co % 4 || d->y_cs % 4 || d->y_co % 4 || d->CoutP % 4) return 0;
  if (d->iD < 3) return 0;
  // narrow variant: <= 4 output channels, possibly a 4-channel slice (y_co) of wider gradient rows (y_cs) whose dW
  // columns land at the same offset of a wider packed row (CoutP): the 15-class logits layer runs as 4 such slices
  if (d->Cin <= 16 && d->Cout <= 4 && (d->iW == 64 || d->iW == 32)) return 2;
  if (d->Cin > 32 || d->Cout != 16) return 0;
  return d->iW == 32 || d->iW == 16;
}

// dWf[tap][ci][co] += sum x * dy for ConvTranspose3d k=7 s=2 p=3 with Cin <= 32, Cout == 16 (same contract as
// crn_conv_wgrad; x is the coarse input [N, D, H, W, x_cs], dy the fine output gradient [N, 2D, 2H, 2W, y_cs]).
extern "C" int crn_convt7_wgrad_line(const crn_conv_desc* d, const float* x, const float* dy, float* dw_packed,
                                     int32_t* status, void* stream) {
  CRN_REQUIRE(d && x && dy && dw_packed && status, "crn_convt7_wgrad_line: null pointer");
  CRN_REQUIRE(crn_convt7_wgrad_line_supported(d), "crn_convt7_wgrad_line: unsupported layer shape");
  WLParams p{};
  p.single = crn_single_pass();
  p.x = x; p.dy = dy; p.dw = dw_packed; p.status = status;
  p.N = d->N; p.D = d->iD; p.H = d->iH; p.W = d->iW;
  p.Cin = d->Cin; p.Cout = d->Cout;
  p.x_cs = d->x_cs; p.x_co = d->x_co; p.y_cs = d->y_cs; p.y_co = d->y_co; p.CinP = d->CinP; p.CoutP = d->CoutP;
  cudaStream_t st = crn_stream(stream);
  if (crn_convt7_wgrad_line_supported(d) == 2) return d->iW == 64 ? launch_wt2l<64>(p, st) : launch_wt2l<32>(p, st);
  return d->iW == 32 ? launch_wtl<32>(p, st) : launch_wtl<16>(p, st);
}

// =============================================================================================================
// Conv3d k=5 s=1 p=2 weight gradient for the WIDE coarse layers (stage_4.c1: 112 -> 64 at 16^3, stage_3.c1:
// 224 -> 128 at 8^3; model/reconstruction_decoder.py:66,74).  The per-tap gather GEMM (conv_wgrad_tc.cu) re-reads
// x and dy for each of the 125 taps and is L2-bound there (1.0 ms); here a CTA owns one (kz, ky) pair, stages an
// image row of x (W + 4 voxels) and the matching dy row ONCE, and gets the 5 kx taps from five start-address
// shifts of the same staged x rows (MN-major rows are whole 128-byte lines, so a shift is just +128 B):
//   M = 128 input channels (4 x 32-channel blocks, hi and lo copies), N = 64 / 128 output channels,
//   5 accumulators (kx) x N TMEM columns, 3xTF32, flushed with float4 atomics.
namespace {

struct WXParams {
  const float* x; const float* dy; float* dw; int* status;
  int N, D, H, W, Cin, Cout, x_cs, x_co, y_cs, y_co, CinP, CoutP;
  int mtiles, ntiles;      // 128-channel Cin tiles, BN-channel Cout tiles
  int lines;               // N * D * H image rows per (kz, ky) pass
  int flush_every;
  int single;          // 1: single-pass TF32 (the lo products are not issued)
};

template <int W, int BN>
__global__ void __launch_bounds__(NTHREADS, 1) wgrad_xline_kernel(const WXParams p) {
  constexpr int XR = W + 8;                         // staged x rows (voxel q = row - 2; rows beyond W + 3 unused)
  constexpr int XB = XR * 128;                      // one 32-channel block of the x line (LBO)
  constexpr int XP = 4 * XB;                        // hi (or lo) part: 128 channels
  constexpr int QB = BN / 32;
  constexpr int YB = W * 128;                       // one 32-channel block of the dy line
  constexpr int YP = QB * YB;
  constexpr int STAGE = 2 * XP + 2 * YP;
  constexpr int NS = 5;                             // stages in the ring
  constexpr int RUNS = W / 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  WLBarriers* B = reinterpret_cast<WLBarriers*>(smem + NS * STAGE);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)tc::uniform_u32((uint32_t)tid >> 5);   // warp-uniform for the compiler (role branches)
  // blockIdx.y = (m tile, n tile); blockIdx.x splits the 25 * lines steps
  const int mt = blockIdx.y / p.ntiles, nt = blockIdx.y % p.ntiles;
  const long long T = 25LL * p.lines;
  const int t0 = (int)(T * blockIdx.x / gridDim.x), t1 = (int)(T * (blockIdx.x + 1) / gridDim.x);

  if (tid == 0) {
    for (int i = 0; i < MAXX; ++i) { tc::mbar_init(&B->full_x[i], 128); tc::mbar_init(&B->empty_x[i], 1); }
    tc::mbar_init(&B->acc_full, 1); tc::mbar_init(&B->acc_empty, 128);
    B->abort_flag = 0;
    tc::mbar_fence_init();
  }
  if (warp == 8) tc::tmem_alloc(&B->tmem_base, 512);
  for (int i = tid; i < NS * STAGE / 16; i += NTHREADS)            // halo rows stay zero
    reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tc::uniform_u32(B->tmem_base);
  const uint32_t smem_u32 = tc::smem_u32(smem);
  auto fail = [&]() { B->abort_flag = 1; *p.status = 1; };
  volatile int* ab = &B->abort_flag;
  auto flush_after = [&](int t) -> bool {
    return t == t1 - 1 || (t + 1) % p.lines == 0 || (t - t0 + 1) % p.flush_every == 0;
  };

  if (warp < 4) {
    // ============================ EPILOGUE: lane = ci, accumulator a = kx, columns = co
    uint32_t nflush = 0;
    bool dead = false;
    const int ci = mt * 128 + warp * 32 + lane;
    for (int t = t0; t < t1 && !dead; ++t) {
      if (!flush_after(t)) continue;
      if (!tc::mbar_wait(&B->acc_full, nflush & 1, ab)) { fail(); dead = true; break; }
      tc::fence_after_sync();
      const int pass = t / p.lines;                  // = kz * 5 + ky
#pragma unroll 1
      for (int a = 0; a < 5; ++a) {
        float* dst = p.dw + ((long long)(pass * 5 + a) * p.CinP + ci) * p.CoutP + nt * BN;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          float v[16];
          tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + a * BN + c0, v);
          if (ci >= p.Cin) continue;
#pragma unroll
          for (int c = 0; c < 16; c += 4)
            if (nt * BN + c0 + c < p.Cout)
              atomicAdd(reinterpret_cast<float4*>(dst + c0 + c), make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
        }
      }
      tc::fence_before_sync();
      tc::mbar_arrive(&B->acc_empty);
      ++nflush;
    }
  } else if (warp < 8) {
    // ============================ PRODUCERS: one job = x line (128 channels) + dy line (BN channels) of a step
    const int pt = tid - 128;
    constexpr int XIT = W * 32 / 128, YIT = W * (BN / 4) / 128, NB = XIT + YIT;
    static_assert(NB <= 8, "line does not fit the register buffer");
    struct Job { int n, z, y, xz, xy, ok; };
    int jt = t0;
    auto next_job = [&](Job& jb) -> bool {
      if (jt >= t1) return false;
      const int pass = jt / p.lines, l = jt - pass * p.lines;
      const int kz = pass / 5, ky = pass - kz * 5;
      jb.y = l % p.H; const int r2 = l / p.H;
      jb.z = r2 % p.D; jb.n = r2 / p.D;
      jb.xz = jb.z + kz - 2; jb.xy = jb.y + ky - 2;
      jb.ok = (unsigned)jb.xz < (unsigned)p.D && (unsigned)jb.xy < (unsigned)p.H;
      ++jt;
      return true;
    };
    auto issue = [&](const Job& jb, float4 (&buf)[NB]) {
      const float* sx = p.x + ((((long long)jb.n * p.D + jb.xz) * p.H + jb.xy) * p.W) * p.x_cs + p.x_co + mt * 128;
#pragma unroll
      for (int u = 0; u < XIT; ++u) {
        const int it = pt + u * 128, v = it >> 5, c4 = it & 31;
        buf[u] = (jb.ok && mt * 128 + c4 * 4 < p.Cin) ? __ldg(reinterpret_cast<const float4*>(sx + (long long)v * p.x_cs + c4 * 4))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float* sy = p.dy + ((((long long)jb.n * p.D + jb.z) * p.H + jb.y) * p.W) * p.y_cs + p.y_co + nt * BN;
#pragma unroll
      for (int u = 0; u < YIT; ++u) {
        const int it = pt + u * 128, v = it / (BN / 4), c4 = it % (BN / 4);
        buf[XIT + u] = (jb.ok && nt * BN + c4 * 4 < p.Cout) ? __ldg(reinterpret_cast<const float4*>(sy + (long long)v * p.y_cs + c4 * 4))
                                                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    uint32_t q = 0;
    bool dead = false;
    auto store = [&](const Job& jb, const float4 (&buf)[NB]) {
      const int slot = (int)(q % NS);
      const uint32_t use = q / NS;
      if (use > 0 && !tc::mbar_wait(&B->empty_x[slot], (use - 1) & 1, ab)) { fail(); dead = true; return; }
      uint8_t* base = smem + slot * STAGE;
#pragma unroll
      for (int u = 0; u < XIT; ++u) {
        const int it = pt + u * 128, v = it >> 5, c4 = it & 31;
        float4 hi, lo;
        tc::split_tf32(buf[u].x, hi.x, lo.x); tc::split_tf32(buf[u].y, hi.y, lo.y);
        tc::split_tf32(buf[u].z, hi.z, lo.z); tc::split_tf32(buf[u].w, hi.w, lo.w);
        const int row = v + 2, cb = c4 >> 3, cw = c4 & 7;
        const uint32_t off = (uint32_t)cb * XB + (uint32_t)row * 128 + (uint32_t)((((cw >> 1) ^ (row & 3)) << 5) + ((cw & 1) << 4));
        *reinterpret_cast<float4*>(base + off) = hi;
        *reinterpret_cast<float4*>(base + XP + off) = lo;
      }
#pragma unroll
      for (int u = 0; u < YIT; ++u) {
        const int it = pt + u * 128, v = it / (BN / 4), c4 = it % (BN / 4);
        float4 hi, lo;
        tc::split_tf32(buf[XIT + u].x, hi.x, lo.x); tc::split_tf32(buf[XIT + u].y, hi.y, lo.y);
        tc::split_tf32(buf[XIT + u].z, hi.z, lo.z); tc::split_tf32(buf[XIT + u].w, hi.w, lo.w);
        const int row = v, cb = c4 >> 3, cw = c4 & 7;
        const uint32_t off = (uint32_t)cb * YB + (uint32_t)row * 128 + (uint32_t)((((cw >> 1) ^ (row & 3)) << 5) + ((cw & 1) << 4));
        *reinterpret_cast<float4*>(base + 2 * XP + off) = hi;
        *reinterpret_cast<float4*>(base + 2 * XP + YP + off) = lo;
      }
      tc::fence_async_smem();
      tc::mbar_arrive(&B->full_x[slot]);
      ++q;
    };
    constexpr int PD = 3;
    Job jobs[PD];
    float4 bufs[PD][NB];
    bool have[PD];
#pragma unroll
    for (int k = 0; k < PD; ++k) {
      have[k] = next_job(jobs[k]);
      if (have[k]) issue(jobs[k], bufs[k]);
    }
    while (have[0] && !dead) {
#pragma unroll
      for (int k = 0; k < PD; ++k) {
        if (!have[k] || dead) { have[0] = false; break; }
        store(jobs[k], bufs[k]);
        have[k] = next_job(jobs[k]);
        if (have[k]) issue(jobs[k], bufs[k]);
      }
    }
  } else {
    // ============================ MMA ISSUER
    {  // converged warp, elected lane issues (tc_common.cuh elect_one)
      const int p_single = p.single;
      constexpr uint32_t idesc = tc::make_idesc_tf32(128, BN, 1, 1);
      uint32_t nflush = 0, slot = 0, phase = 0;
      bool first = true, dead = false;
      int mod_ctr = 0;
      int line = t0 % p.lines;
      for (int t = t0; t < t1 && !dead; ++t) {
        if (!tc::mbar_wait(&B->full_x[slot], phase, ab)) { fail(); dead = true; break; }
        if (first && nflush > 0) {
          if (!tc::mbar_wait(&B->acc_empty, (nflush - 1) & 1, ab)) { fail(); dead = true; break; }
        }
        tc::fence_after_sync();
        const uint32_t base = smem_u32 + slot * STAGE;
        const uint64_t dxh = tc::make_desc_mn32(base, XB, 512), dxl = tc::make_desc_mn32(base + XP, XB, 512);
        const uint64_t dyh = tc::make_desc_mn32(base + 2 * XP, YB, 512), dyl = tc::make_desc_mn32(base + 2 * XP + YP, YB, 512);
#pragma unroll
        for (int r = 0; r < RUNS; ++r) {
#pragma unroll
          for (int a = 0; a < 5; ++a) {                    // kx = a: x rows shifted by a
            const uint64_t xo = (uint64_t)((r * 8 + a) * 8), yo = (uint64_t)(r * 64);
            tc::mma_tf32_e(tmem + a * BN, dxh + xo, dyh + yo, idesc, (first && r == 0) ? 0u : 1u);
            if (!p_single) {
              tc::mma_tf32_e(tmem + a * BN, dxl + xo, dyh + yo, idesc, 1u);
              tc::mma_tf32_e(tmem + a * BN, dxh + xo, dyl + yo, idesc, 1u);
            }
          }
        }
        first = false;
        tc::commit_e(&B->empty_x[slot]);
        if (++slot == NS) { slot = 0; phase ^= 1; }
        if (++mod_ctr == p.flush_every) mod_ctr = 0;
        if (++line == p.lines) line = 0;
        if (t == t1 - 1 || line == 0 || mod_ctr == 0) {
          tc::commit_e(&B->acc_full);
          ++nflush;
          first = true;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

template <int W, int BN>
int launch_wx(WXParams p, cudaStream_t st) {
  constexpr int STAGE = 2 * 4 * (W + 8) * 128 + 2 * (BN / 32) * W * 128;
  const size_t smem = (size_t)5 * STAGE + sizeof(WLBarriers) + 1024 + 64;
  auto kern = wgrad_xline_kernel<W, BN>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      crn_set_error("conv_wgrad_xline: cannot set %zu bytes of dynamic shared memory", smem);
      return CRN_ERR_LAUNCH;
    }
    configured = true;
  }
  p.flush_every = 64;
  const int tiles = p.mtiles * p.ntiles;
  long long gx = kNumSMs / tiles;
  if (gx < 1) gx = 1;
  if (gx > 25LL * p.lines) gx = 25LL * p.lines;
  dim3 grid((unsigned)gx, (unsigned)tiles);
  kern<<<grid, NTHREADS, smem, st>>>(p);
  CRN_LAUNCH_CHECK("conv_wgrad_xline");
  return CRN_OK;
}
}  // namespace

extern "C" int crn_conv_wgrad_xline_supported(const crn_conv_desc* d) {
  if (!d || d->transposed || d->kD != 5 || d->kH != 5 || d->kW != 5 || d->stride != 1 || d->pad != 2) return 0;
  if (d->iD != d->oD || d->iH != d->oH || d->iW != d->oW || d->y_planar) return 0;
  if (d->Cin % 4 || d->Cout % 4 || d->x_cs % 4 || d->x_co % 4 || d->y_cs % 4 || d->y_co % 4 || d->CoutP % 4) return 0;
  if (d->Cin < 64 || d->Cout < 32 || d->Cout > 128) return 0;
  return d->iW == 16 || d->iW == 8;
}

// dWf[tap][ci][co] += sum_vox x * dy for wide Conv3d k=5 layers on 16^3 / 8^3 grids (same contract as crn_conv_wgrad).
extern "C" int crn_conv_wgrad_xline(const crn_conv_desc* d, const float* x, const float* dy, float* dw_packed,
                                    int32_t* status, void* stream) {
  CRN_REQUIRE(d && x && dy && dw_packed && status, "crn_conv_wgrad_xline: null pointer");
  CRN_REQUIRE(crn_conv_wgrad_xline_supported(d), "crn_conv_wgrad_xline: unsupported layer shape");
  WXParams p{};
  p.single = crn_single_pass();
  p.x = x; p.dy = dy; p.dw = dw_packed; p.status = status;
  p.N = d->N; p.D = d->iD; p.H = d->iH; p.W = d->iW;
  p.Cin = d->Cin; p.Cout = d->Cout;
  p.x_cs = d->x_cs; p.x_co = d->x_co; p.y_cs = d->y_cs; p.y_co = d->y_co; p.CinP = d->CinP; p.CoutP = d->CoutP;
  p.lines = d->N * d->iD * d->iH;
  // 5 accumulators (kx) x 64 columns = 320 of the 512 TMEM columns: Cout is tiled by 64
  p.mtiles = (d->Cin + 127) / 128; p.ntiles = (d->Cout + 63) / 64;
  cudaStream_t st = crn_stream(stream);
  return d->iW == 16 ? launch_wx<16, 64>(p, st) : launch_wx<8, 64>(p, st);
}

#ifdef CRN_DIAG
extern "C" int crn_wgrad_line_debug_read(long long* host_dst, int32_t n);
extern "C" int crn_wgrad_line_debug_read(long long* host_dst, int32_t n) {
  return cudaMemcpyFromSymbol(host_dst, g_wl_dbg, sizeof(long long) * (n > kNumSMs * 8 ? kNumSMs * 8 : n)) == cudaSuccess
             ? CRN_OK : CRN_ERR_LAUNCH;
}
#endif  // CRN_DIAG
