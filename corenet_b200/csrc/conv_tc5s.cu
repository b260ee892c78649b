// Conv3d k=5 s=1 p=2 FORWARD for <= 16 output channels (stage_6.c1: 28 -> 16 at 64^3, the single most expensive
// launch; model/reconstruction_decoder.py:91) with the 5 kz taps STACKED INTO N.
//
// conv_tc5_kernel issues one 128 x 32 x 8 MMA per (tap, output plane): ncu shows its tensor pipe 16 % busy -- the
// kernel is bound by re-fetching the 4 KB A tile (one staged input plane, tap-shifted) from shared memory for every
// MMA.  But one staged input plane q contributes to the FIVE output planes z = q - kz (+2), whose accumulators are
// adjacent TMEM columns: stacking the weights of kz = 4..0 as consecutive B rows turns those five MMAs into ONE with
// N = 5 x 32, i.e. one A fetch per 5x more columns (2.5x fewer, wider MMAs per output voxel).
//   * item = (8 x 16) xy tile x ZT = 4 output planes; the ZT + 4 = 8 input planes of a K pass (8 channels) are all
//     resident (ring of 8 slots, refilled plane by plane during the last ky sweep of the previous pass);
//   * loop: pass -> ky (weights of all kz, kx for this ky: one 50 KB cp.async.bulk, 2-stage ring) -> input plane q
//     -> kx -> {A_hi x [W_hi | W_lo], A_lo x [W_hi | 0]} over the stack of valid output planes;
//   * the first MMA that touches a not-yet-started accumulator of the (pass, ky) flush group overwrites it; the
//     not-started planes are always the upper end of the stack, so that MMA is split in two at most once per q;
//   * epilogue: register-resident running sums (4 planes x 16 columns), one flush per (pass, ky) group (80 MMAs),
//     one store per item; warp roles and bounded waits as in conv_tc5.cu.
// WIDE variant (17..32 output channels: stage_6.c1 dgrad 16 -> 28, stage_5.c1 forward 56 -> 32): the 32 columns of a
// plane are 32 channels, the stacked B regions are [W_hi] and [W_lo] (same 10 KB per tap) and the three products
// A_hi x W_hi, A_lo x W_hi, A_hi x W_lo are three stacked MMAs.  dgrad = the same kernel on dy with the flipped,
// transposed filter (packed that way).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TX = 8, TY = 16;
constexpr int XS = TX + 4, YS = TY + 4;
constexpr int CHUNK_BYTES = YS * XS * 16;        // one 4-channel chunk of a plane  (3840)
constexpr int PART_BYTES = 2 * CHUNK_BYTES;      // 8 channels                       (7680)
constexpr int PLANE_BYTES = 2 * PART_BYTES;      // hi + lo                          (15360)
constexpr int ZT = 4;
constexpr int MAXSLOT = 10;                      // plane ring capacity
constexpr int NP = 16;                           // padded output channels
constexpr int BLK = 2 * NP;                      // B rows / accumulator columns per output plane: [hi | lo]
// KT = taps per axis: 5 (Conv3d k=5) or 4 (ConvTranspose3d k=7 s=2 seen from the coarse grid, see conv_tc5.cu)
template <int KT>
struct Geo {
  // planes per K pass, and ring slots: KT = 4 has shared memory left for 3 extra slots, so the producers can stage the
  // first planes of the next pass while the MMAs of this pass still run (KT = 5 fills the 227 KB with NPLANE slots)
  static constexpr int NPLANE = ZT + KT - 1, NSLOT = KT == 4 ? NPLANE + 3 : NPLANE;
  static constexpr int SROWS = KT * BLK;                 // rows of one stacked part (kz = KT-1..0)
  static constexpr int KC_BYTES = 2 * SROWS * 16;        // one 4-channel K chunk of a tap: part 0 rows + part 1 rows (LBO)
  static constexpr int TAP_BYTES = 2 * KC_BYTES;         // 10240 (KT = 5) / 8192 (KT = 4)
  static constexpr int WROW_BYTES = KT * TAP_BYTES;      // the KT kx taps of one ky (all kz): 51200 / 32768
};
constexpr int WSTAGES = 2;
constexpr int NTHREADS = 320;
constexpr int ACOLS = BLK;                       // 32 accumulator columns per plane
constexpr int TMEM_COLS = 256;                   // 2 stages x 4 planes x 32 columns

struct TC5SParams {
  const float* in; const float* wtc; const float* bias; float* out; int* status;
  int N, D, H, W, gK, gN, in_cs, in_co, out_cs, out_co, P;
  int tiles_x, tiles_y, tiles_z, nitems;
  int cout_cls;            // GATH / SCAT: channels per parity class of the fine grid (K resp. N = 8 * cout_cls)
  int planar;              // SCAT: out is [N, Cout, 2D, 2H, 2W] instead of channels-last rows
  int single;              // 1: single-pass TF32 (the A_lo / W_lo products are not issued)
  int dbg;                 // crn_set_flags bit 8: per-CTA wait-time accounting (crn_tc5s_debug_read)
  int in_planar;           // GATH: the fine gradient is planar [N, 2, 2D, 2H, 2W] (cout_cls == 2, the FG_BG logits)
};

struct __align__(8) Barriers {
  uint64_t plane_full[MAXSLOT], plane_empty[MAXSLOT];
  uint64_t w_full[WSTAGES], w_empty[WSTAGES];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  int abort_flag;
};

// debug (crn_set_flags bit 8): per CTA [mma total, mma wait acc_empty, w_full, plane_full, producer total, producer wait
// plane_empty, epilogue total, epilogue wait acc_full] in clock64 cycles
__device__ long long g_tc5s_dbg[kNumSMs * 8];
#define TC5S_TIMED(accum, call)                         \
  ([&]() -> bool {                                      \
    const long long t0__ = p.dbg ? clock64() : 0;       \
    const bool ok__ = (call);                           \
    if (p.dbg) accum += clock64() - t0__;               \
    return ok__;                                        \
  }())

__device__ __forceinline__ void decode_item(const TC5SParams& p, int item, int& n, int& z0, int& y0, int& x0) {
  int t = item;
  x0 = (t % p.tiles_x) * TX; t /= p.tiles_x;
  y0 = (t % p.tiles_y) * TY; t /= p.tiles_y;
  z0 = (t % p.tiles_z) * ZT; n = t / p.tiles_z;
}

// GATH: dgrad of ConvTranspose3d k=7 s=2 p=3 -- K = 8 * cout_cls class channels gathered from the fine gradient
// (class channel (c, co) of coarse voxel i is dY[2i + c][co]), 4^3 taps at offsets -2..1, N = Cin (conv_tc5.cu GATH).
// SCAT: forward of that layer for 8 * Cout <= 16 class columns (the FG_BG logits layer, Cout = 2): 4^3 taps at offsets
// -1..2 on the coarse input, the epilogue scatters column (class, co) of coarse voxel i to fine voxel 2i + class.
template <bool WIDE, int KT, bool GATH, bool SCAT>
__global__ void __launch_bounds__(NTHREADS, 1) conv_tc5s_kernel(const TC5SParams p) {
  constexpr int NOUT = WIDE ? 32 : 16;             // output channels held per plane
  constexpr int HLO = (KT == 5 || GATH) ? 2 : 1;   // most negative tap offset
  constexpr int NPLANE = Geo<KT>::NPLANE, NSLOT = Geo<KT>::NSLOT, SROWS = Geo<KT>::SROWS;
  constexpr int KC_BYTES = Geo<KT>::KC_BYTES, TAP_BYTES = Geo<KT>::TAP_BYTES, WROW_BYTES = Geo<KT>::WROW_BYTES;
  extern __shared__ __align__(1024) uint8_t smem[];
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
    // ============================ EPILOGUE: one flush per (pass, ky) group into register sums, one store per item
    long long G = 0;
    bool dead = false;
    long long tw_acc = 0;
    const long long t_begin = p.dbg ? clock64() : 0;
    for (int item = blockIdx.x; item < p.nitems && !dead; item += gridDim.x) {
      int n, z0, y0, x0;
      decode_item(p, item, n, z0, y0, x0);
      const int m = warp * 32 + lane;
      const int y = y0 + (m >> 3), x = x0 + (m & 7);
      float sum[ZT][NOUT];
#pragma unroll
      for (int zz = 0; zz < ZT; ++zz)
#pragma unroll
        for (int e = 0; e < NOUT; ++e) sum[zz][e] = 0.f;
      for (int pass = 0; pass < p.P && !dead; ++pass) {
        for (int ky = 0; ky < KT; ++ky, ++G) {
          const int st = (int)(G & 1);
          if (!TC5S_TIMED(tw_acc, tc::mbar_wait(&B->acc_full[st], (uint32_t)(G >> 1) & 1, ab))) { fail(); dead = true; break; }
          tc::fence_after_sync();
#pragma unroll
          for (int zz = 0; zz < ZT; ++zz) {
            // plane zz always receives at least its kz = HLO contribution (input plane z0 + zz is inside the volume)
            const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + st * (ZT * ACOLS) + zz * ACOLS;
            float v[16];
            tc::tmem_ld16(ta, v);
#pragma unroll
            for (int e = 0; e < 16; ++e) sum[zz][e] += v[e];
            tc::tmem_ld16(ta + NP, v);
#pragma unroll
            for (int e = 0; e < 16; ++e) sum[zz][WIDE ? 16 + e : e] += v[e];   // ND16: hi*lo half folds onto the same 16
          }
          tc::fence_before_sync();
          tc::mbar_arrive(&B->acc_empty[st]);
        }
      }
      if (dead) break;
      if constexpr (SCAT) {
        const int OH = 2 * p.H, OW = 2 * p.W;
        const long long oS = (long long)(2 * p.D) * OH * OW;
#pragma unroll
        for (int zz = 0; zz < ZT; ++zz) {
          if (p.planar && p.cout_cls == 2) {
            // px = 0 / 1 classes of a coarse voxel are neighbouring fine voxels of a plane row: 8-byte stores
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {                      // qd = pz * 2 + py
              const int oz = 2 * (z0 + zz) + (qd >> 1), oy = 2 * y + (qd & 1);
              const long long sp = ((long long)oz * OH + oy) * OW + 2 * x;
#pragma unroll
              for (int co = 0; co < 2; ++co) {
                const float b = p.bias ? __ldg(p.bias + co) : 0.f;
                *reinterpret_cast<float2*>(p.out + ((long long)n * 2 + co) * oS + sp) =
                    make_float2(sum[zz][qd * 4 + co] + b, sum[zz][qd * 4 + 2 + co] + b);
              }
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              if (e >= p.gN) continue;
              const int cls = e / p.cout_cls, co = e - cls * p.cout_cls;
              const int oz = 2 * (z0 + zz) + (cls >> 2), oy = 2 * y + ((cls >> 1) & 1), ox = 2 * x + (cls & 1);
              const long long sp = ((long long)oz * OH + oy) * OW + ox;
              float* dst = p.planar ? p.out + ((long long)n * p.cout_cls + co) * oS + sp
                                    : p.out + ((long long)n * oS + sp) * p.out_cs + p.out_co + co;
              *dst = sum[zz][e] + (p.bias ? __ldg(p.bias + co) : 0.f);
            }
          }
        }
      } else {
#pragma unroll
      for (int zz = 0; zz < ZT; ++zz) {
        const long long pos = (((long long)n * p.D + (z0 + zz)) * p.H + y) * p.W + x;
        float* dst = p.out + pos * p.out_cs + p.out_co;
#pragma unroll
        for (int qd = 0; qd < NOUT / 4; ++qd) {
          const int c = qd * 4;
          if (c < p.gN) {                          // gN is a multiple of 4
            float4 o = make_float4(sum[zz][c], sum[zz][c + 1], sum[zz][c + 2], sum[zz][c + 3]);
            if (p.bias) {
              o.x += __ldg(p.bias + c); o.y += __ldg(p.bias + c + 1); o.z += __ldg(p.bias + c + 2); o.w += __ldg(p.bias + c + 3);
            }
            *reinterpret_cast<float4*>(dst + c) = o;
          }
        }
      }
      }
    }
    if (p.dbg && tid == 0) {
      g_tc5s_dbg[blockIdx.x * 8 + 6] = clock64() - t_begin;
      g_tc5s_dbg[blockIdx.x * 8 + 7] = tw_acc;
    }
  } else if (warp < 8) {
    // ============================ PRODUCERS: halo gather + hi/lo split into the plane ring.
    // Software-pipelined: the gathers of plane L+1 are issued (into registers) right after plane L has been handed over,
    // i.e. BEFORE the wait for slot L+1 to become free, so the global-memory latency overlaps that wait and the
    // turnaround of a slot (freed -> full again) is only split + shared stores.
    const int pt = tid - 128;
    constexpr int NU = (YS * XS * 2 + 127) / 128;
    int item = blockIdx.x, pass = 0, r = 0;
    int n = 0, z0 = 0, y0 = 0, x0 = 0;
    if (item < p.nitems) decode_item(p, item, n, z0, y0, x0);
    float4 a[NU];
    bool inside = false;                           // plane inside the volume (else: nothing to stage, MMAs skip it)
    auto gather = [&]() {
      const int q = z0 - HLO + r;
      inside = q >= 0 && q < p.D;
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        const int u = pt + i * 128;
        a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inside && u < YS * XS * 2) {
          const int kc = u & 1; const int v = u >> 1;
          const int xs = v % XS, ys = v / XS;
          const int y = y0 - HLO + ys, x = x0 - HLO + xs;
          const int k = pass * 8 + kc * 4;
          if ((unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W && k < p.gK) {
            long long off;
            if constexpr (GATH) {                  // class channel k = (cls, co) lives at fine voxel 2i + cls
              if (p.in_planar) {
                // cout_cls == 2: chunk k..k+3 = classes (cls, cls + 1), i.e. px = 0 / 1, x channels (0, 1): the two
                // fine voxels are neighbours of a plane row -> one float2 per channel plane
                const int cls = k >> 1;
                const long long fS = (long long)(2 * p.D) * (2 * p.H) * (2 * p.W);
                const long long sp = ((long long)(2 * q + (cls >> 2)) * (2 * p.H) + 2 * y + ((cls >> 1) & 1)) * (2 * p.W) + 2 * x;
                const float2 c0 = __ldg(reinterpret_cast<const float2*>(p.in + ((long long)n * 2) * fS + sp));
                const float2 c1 = __ldg(reinterpret_cast<const float2*>(p.in + ((long long)n * 2 + 1) * fS + sp));
                a[i] = make_float4(c0.x, c1.x, c0.y, c1.y);
                continue;
              }
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
    };
    if (item < p.nitems) gather();
    long long tw_slot = 0;
    const long long t_begin = p.dbg ? clock64() : 0;
    for (long long L = 0; item < p.nitems; ++L) {
      const int slot = (int)(L % NSLOT);
      const uint32_t use = (uint32_t)(L / NSLOT);
      if (use > 0 && !TC5S_TIMED(tw_slot, tc::mbar_wait(&B->plane_empty[slot], (use - 1) & 1, ab))) { fail(); break; }
      if (inside) {
        uint8_t* dst = ring + slot * PLANE_BYTES;
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
      // next (item, pass, plane)
      if (++r == NPLANE) {
        r = 0;
        if (++pass == p.P) {
          pass = 0;
          item += gridDim.x;
          if (item < p.nitems) decode_item(p, item, n, z0, y0, x0);
        }
      }
      if (item < p.nitems) gather();
    }
    if (p.dbg && pt == 0) {
      g_tc5s_dbg[blockIdx.x * 8 + 4] = clock64() - t_begin;
      g_tc5s_dbg[blockIdx.x * 8 + 5] = tw_slot;
    }
  } else if (warp == 8) {
    // ============================ MMA ISSUER: the whole warp runs the loop converged, one elected lane issues
    {
      constexpr uint32_t A_DESC_HI = (uint32_t)((XS * 16) >> 4) | (1u << 14);   // SBO field | version 1 (bit 46)
      // ring slot / phase parity of plane r = 0 of the current pass and the weight-row / group counters are carried
      // incrementally in 32-bit registers (no 64-bit modulo on the issue path)
      int slot0 = 0;
      uint32_t par0 = 0;
      uint32_t Wn = 0;                              // running weight-row (ky) index
      uint32_t G = 0;                               // running (pass, ky) group index -> accumulator stage
      bool dead = false;
      long long tw_acc = 0, tw_w = 0, tw_plane = 0;
      const long long t_begin = p.dbg ? clock64() : 0;
      const int p_single = p.single, p_dbg = p.dbg;
      for (int item = blockIdx.x; item < p.nitems && !dead; item += gridDim.x) {
        int n, z0, y0, x0;
        decode_item(p, item, n, z0, y0, x0);
        for (int pass = 0; pass < p.P && !dead; ++pass) {
          for (int ky = 0; ky < KT && !dead; ++ky, ++G, ++Wn) {
            const int st = (int)(G & 1);
            if (G >= 2 && !TC5S_TIMED(tw_acc, tc::mbar_wait(&B->acc_empty[st], (uint32_t)((G >> 1) - 1) & 1, ab))) { fail(); dead = true; break; }
            const int ws = (int)(Wn % WSTAGES);
            if (!TC5S_TIMED(tw_w, tc::mbar_wait(&B->w_full[ws], (uint32_t)(Wn / WSTAGES) & 1, ab))) { fail(); dead = true; break; }
            tc::fence_after_sync();
            const uint32_t wbase = wring_u32 + ws * WROW_BYTES;
            const uint32_t acc0 = tmem + st * (ZT * ACOLS);
            int ns = 0;                             // accumulators [0, ns) of this group have been written
#pragma unroll 1
            for (int q = 0; q < NPLANE; ++q) {
              int slot = slot0 + q;
              uint32_t par = par0;
              if (slot >= NSLOT) { slot -= NSLOT; par ^= 1u; }
              if (ky == 0) {                        // planes arrive during the first sweep of a pass
                if (!TC5S_TIMED(tw_plane, tc::mbar_wait(&B->plane_full[slot], par, ab))) { fail(); dead = true; break; }
                tc::fence_after_sync();
              }
              const int zin = z0 - HLO + q;
              if (zin >= 0 && zin < p.D) {
                // output planes zz = q - kz, kz = 0..KT-1, clipped to the item: the stack [zlo, zhi]
                const int zlo = q > KT - 1 ? q - (KT - 1) : 0, zhi = q < ZT - 1 ? q : ZT - 1;
                // B rows: block b holds kz = KT-1 - b; plane zlo needs kz = q - zlo -> first block KT-1 - (q - zlo)
                const uint32_t boff = (uint32_t)(KT - 1 - (q - zlo)) * (BLK * 16);
                // descriptor words: everything that does not depend on (kx, part) is folded into a_q / b_q here, so one
                // MMA costs two uniform adds (smem addresses >> 4 stay below 2^14: the adds never carry into LBO)
                const uint32_t a_q = (((ring_u32 + (uint32_t)slot * PLANE_BYTES) >> 4) | ((uint32_t)(CHUNK_BYTES >> 4) << 16)) +
                                     (uint32_t)(ky * XS);
                const uint32_t b_q = (((wbase + boff) >> 4) & 0x3FFFu) | ((uint32_t)(KC_BYTES >> 4) << 16);
                constexpr uint32_t B_DESC_HI = (uint32_t)(128 >> 4) | (1u << 14);     // SBO | version 1
                const int cnt = zhi - zlo + 1;
                const uint32_t idesc_q = tc::make_idesc_tf32(128, BLK * cnt, 0, 0);
                const uint32_t acc_q = acc0 + zlo * ACOLS;
                const bool fresh = ns <= zhi;          // some planes of the stack get their first contribution now
                const int zs = ns > zlo ? ns : zlo;
                if (fresh) ns = zhi + 1;
                const bool skip = (p_dbg & 2) != 0;
#pragma unroll
                for (int kx = 0; kx < KT; ++kx) {
                  // ND16: (A_hi, [W_hi|W_lo]), (A_lo, [W_hi|0]);   WIDE: (A_hi, W_hi), (A_lo, W_hi), (A_hi, W_lo)
#pragma unroll
                  for (int part = 0; part < (WIDE ? 3 : 2); ++part) {
                    if ((part && p_single) || skip) break;
                    const bool a_lo = part == 1;
                    const bool b_r1 = WIDE ? part == 2 : part == 1;
                    const uint64_t da = ((uint64_t)A_DESC_HI << 32) | (a_q + (uint32_t)(kx + (a_lo ? (PART_BYTES >> 4) : 0)));
                    const uint32_t blo = b_q + (uint32_t)((kx * TAP_BYTES + (b_r1 ? SROWS * 16 : 0)) >> 4);
                    const uint64_t db = ((uint64_t)B_DESC_HI << 32) | blo;
                    if (kx == 0 && part == 0 && fresh) {
                      // planes [zs, zhi] are overwritten (first MMA of the group on them), [zlo, zs) accumulate
                      if (zs > zlo)
                        tc::mma_tf32_e(acc_q, da, db, tc::make_idesc_tf32(128, BLK * (zs - zlo), 0, 0), 1u);
                      tc::mma_tf32_e(acc0 + zs * ACOLS, da, db + (uint64_t)((uint32_t)(zs - zlo) * (BLK * 16 >> 4)),
                                     tc::make_idesc_tf32(128, BLK * (zhi - zs + 1), 0, 0), 0u);
                    } else {
                      tc::mma_tf32_e(acc_q, da, db, idesc_q, 1u);
                    }
                  }
                }
              }
              if (ky == KT - 1) tc::commit_e(&B->plane_empty[slot]); // last sweep of the pass: plane q is dead
            }
            if (dead) break;
            tc::commit_e(&B->w_empty[ws]);
            tc::commit_e(&B->acc_full[st]);
          }
          slot0 += NPLANE;
          if (slot0 >= NSLOT) { slot0 -= NSLOT; par0 ^= 1u; }
        }
      }
      if (p.dbg && lane == 0) {
        g_tc5s_dbg[blockIdx.x * 8 + 0] = clock64() - t_begin;
        g_tc5s_dbg[blockIdx.x * 8 + 1] = tw_acc;
        g_tc5s_dbg[blockIdx.x * 8 + 2] = tw_w;
        g_tc5s_dbg[blockIdx.x * 8 + 3] = tw_plane;
      }
    }
  } else {
    // ============================ WEIGHT COPIES: one cp.async.bulk per (pass, ky)
    if (lane == 0) {
      long long Wn = 0;
      bool dead = false;
      for (int item = blockIdx.x; item < p.nitems && !dead; item += gridDim.x) {
        for (int pass = 0; pass < p.P && !dead; ++pass) {
          for (int ky = 0; ky < KT; ++ky, ++Wn) {
            const int ws = (int)(Wn % WSTAGES);
            const uint32_t use = (uint32_t)(Wn / WSTAGES);
            if (use > 0 && !tc::mbar_wait(&B->w_empty[ws], (use - 1) & 1, ab)) { fail(); dead = true; break; }
            const float* src = p.wtc + ((size_t)pass * KT + ky) * (WROW_BYTES / 4);
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
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, TMEM_COLS);
}

// pack: PyTorch conv weight [Cout][Cin][125] -> wtc[P][ky][kx][kc][region][blk = 4 - kz][32 rows][4 floats]
//   ND16 (N <= 16): region 0 rows [hi(n 0..15) | lo(n 0..15)], region 1 rows [hi(n 0..15) | 0]
//   WIDE (N <= 32): region 0 rows hi(n 0..31),                 region 1 rows lo(n 0..31)
//   fwd: k = ci, n = co, tap t;   dgrad: k = co, n = ci, tap 124 - t
__global__ void tc5s_pack_kernel(const float* __restrict__ w, int Cout, int Cin, int dgrad, int wide, int P,
                                 float* __restrict__ out) {
  const int K = dgrad ? Cout : Cin, Nn = dgrad ? Cin : Cout;
  const long long total = (long long)P * 25 * 2 * 5 * 32 * 4;      // (pass, ky, kx, kc, kz, n, e)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i & 3); long long r = i >> 2;
    const int n = (int)(r & 31); r >>= 5;
    const int kz = (int)(r % 5); r /= 5;
    const int kc = (int)(r & 1); r >>= 1;
    const int kx = (int)(r % 5); r /= 5;
    const int ky = (int)(r % 5); const int pass = (int)(r / 5);
    if (!wide && n >= 16) continue;
    const int k = pass * 8 + kc * 4 + e;
    float v = 0.f;
    if (k < K && n < Nn) {
      const int co = dgrad ? k : n, ci = dgrad ? n : k;
      const int t = (kz * 5 + ky) * 5 + kx;
      v = w[((long long)co * Cin + ci) * 125 + (dgrad ? 124 - t : t)];
    }
    float hi, lo;
    tc::split_tf32(v, hi, lo);
    const long long kcb = (((long long)pass * 5 + ky) * 5 + kx) * (Geo<5>::TAP_BYTES / 4) +
                          (long long)kc * (Geo<5>::KC_BYTES / 4);
    const long long row0 = kcb + ((long long)(4 - kz) * BLK + n) * 4 + e;              // region 0
    const long long row1 = row0 + (long long)Geo<5>::SROWS * 4;                       // region 1
    if (wide) {
      out[row0] = hi;
      out[row1] = lo;
    } else {
      out[row0] = hi; out[row0 + NP * 4] = lo;
      out[row1] = hi; out[row1 + NP * 4] = 0.f;
    }
  }
}

// class-gather dgrad weights (ConvTranspose3d [Cin][Cout][7][7][7]) in the stacked layout:
//   wtc[P][jy][jx][kc][region][blk = 3 - jz][32 rows][4 floats], K index kk = (class, co), N = ci,
//   tap j <-> offset j - 2 per axis, filter index k = c - 1 + 2*j (zero outside 0..6)   (conv_tc5.cu tct_pack, dgrad)
__global__ void tcts_pack_kernel(const float* __restrict__ w, int Cin, int Cout, int wide, int P,
                                 float* __restrict__ out) {
  const long long total = (long long)P * 16 * 2 * 4 * 32 * 4;      // (pass, jy, jx, kc, jz, n, e)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i & 3); long long r = i >> 2;
    const int n = (int)(r & 31); r >>= 5;
    const int jz = (int)(r & 3); r >>= 2;
    const int kc = (int)(r & 1); r >>= 1;
    const int jx = (int)(r & 3); r >>= 2;
    const int jy = (int)(r & 3); const int pass = (int)(r >> 2);
    if (!wide && n >= 16) continue;
    const int kk = pass * 8 + kc * 4 + e;
    float v = 0.f;
    if (n < Cin && kk < 8 * Cout) {
      const int cls = kk / Cout, co = kk - cls * Cout;
      const int kz = (cls >> 2) - 1 + 2 * jz, ky = ((cls >> 1) & 1) - 1 + 2 * jy, kx = (cls & 1) - 1 + 2 * jx;
      if ((unsigned)kz < 7u && (unsigned)ky < 7u && (unsigned)kx < 7u)
        v = w[(((long long)n * Cout + co) * 7 + kz) * 49 + ky * 7 + kx];
    }
    float hi, lo;
    tc::split_tf32(v, hi, lo);
    const long long kcb = (((long long)pass * 4 + jy) * 4 + jx) * (Geo<4>::TAP_BYTES / 4) +
                          (long long)kc * (Geo<4>::KC_BYTES / 4);
    const long long row0 = kcb + ((long long)(3 - jz) * BLK + n) * 4 + e;              // region 0
    const long long row1 = row0 + (long long)Geo<4>::SROWS * 4;                       // region 1
    if (wide) {
      out[row0] = hi;
      out[row1] = lo;
    } else {
      out[row0] = hi; out[row0 + NP * 4] = lo;
      out[row1] = hi; out[row1 + NP * 4] = 0.f;
    }
  }
}

// forward class weights of ConvTranspose3d [Cin][Cout][7][7][7] with 8 * Cout <= 16 in the stacked ND16 layout:
//   K = ci, N column n = class * Cout + co, tap j <-> input offset j - 1 per axis, filter index k = c + 5 - 2*j
__global__ void tctsf_pack_kernel(const float* __restrict__ w, int Cin, int Cout, int P, float* __restrict__ out) {
  const long long total = (long long)P * 16 * 2 * 4 * 16 * 4;      // (pass, jy, jx, kc, jz, n, e)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i & 3); long long r = i >> 2;
    const int n = (int)(r & 15); r >>= 4;
    const int jz = (int)(r & 3); r >>= 2;
    const int kc = (int)(r & 1); r >>= 1;
    const int jx = (int)(r & 3); r >>= 2;
    const int jy = (int)(r & 3); const int pass = (int)(r >> 2);
    const int ci = pass * 8 + kc * 4 + e;
    float v = 0.f;
    if (ci < Cin && n < 8 * Cout) {
      const int cls = n / Cout, co = n - cls * Cout;
      const int kz = (cls >> 2) + 5 - 2 * jz, ky = ((cls >> 1) & 1) + 5 - 2 * jy, kx = (cls & 1) + 5 - 2 * jx;
      if ((unsigned)kz < 7u && (unsigned)ky < 7u && (unsigned)kx < 7u)
        v = w[(((long long)ci * Cout + co) * 7 + kz) * 49 + ky * 7 + kx];
    }
    float hi, lo;
    tc::split_tf32(v, hi, lo);
    const long long kcb = (((long long)pass * 4 + jy) * 4 + jx) * (Geo<4>::TAP_BYTES / 4) +
                          (long long)kc * (Geo<4>::KC_BYTES / 4);
    const long long row0 = kcb + ((long long)(3 - jz) * BLK + n) * 4 + e;              // region 0
    const long long row1 = row0 + (long long)Geo<4>::SROWS * 4;                       // region 1
    out[row0] = hi; out[row0 + NP * 4] = lo;
    out[row1] = hi; out[row1 + NP * 4] = 0.f;
  }
}

template <bool WIDE, int KT = 5, bool GATH = false, bool SCAT = false>
int launch_tc5s(const TC5SParams& p, cudaStream_t st) {
  const size_t smem = (size_t)Geo<KT>::NSLOT * PLANE_BYTES + (size_t)WSTAGES * Geo<KT>::WROW_BYTES + sizeof(Barriers) + 64;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(conv_tc5s_kernel<WIDE, KT, GATH, SCAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      crn_set_error("conv_tc5s: cannot set %zu bytes of dynamic shared memory", smem);
      return CRN_ERR_LAUNCH;
    }
    configured = true;
  }
  const int grid = p.nitems < kNumSMs ? p.nitems : kNumSMs;
  conv_tc5s_kernel<WIDE, KT, GATH, SCAT><<<grid, NTHREADS, smem, st>>>(p);
  CRN_LAUNCH_CHECK("conv_tc5s");
  return CRN_OK;
}

}  // namespace

// K = reduction channels (Cin for the forward operator, Cout for dgrad)
extern "C" int64_t crn_tc5s_packed_floats(int32_t K) { return (int64_t)((K + 7) / 8) * 5 * (Geo<5>::WROW_BYTES / 4); }

// ---- dgrad of ConvTranspose3d k=7 s=2 p=3 with the 4 jz taps stacked into N (Cin <= 32; Cout % 4 == 0 with a
// channels-last dy, or Cout == 2 with the planar FG_BG logit gradient)
extern "C" int64_t crn_tcts_packed_floats(int32_t Cout) { return (int64_t)(8 * Cout / 8) * 4 * (Geo<4>::WROW_BYTES / 4); }

extern "C" int crn_tcts_pack(const float* w, int32_t Cin, int32_t Cout, float* out, void* stream) {
  CRN_REQUIRE(w && out && Cout > 0 && Cin > 0 && Cin <= 32 && (Cout % 4 == 0 || Cout == 2),
              "crn_tcts_pack: Cin <= 32, Cout % 4 == 0 or Cout == 2");
  const int P = Cout;                                   // K = 8 * Cout class channels in passes of 8
  const long long total = (long long)P * 16 * 2 * 4 * 32 * 4;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  tcts_pack_kernel<<<blocks, 256, 0, crn_stream(stream)>>>(w, Cin, Cout, Cin > 16 ? 1 : 0, P, out);
  CRN_LAUNCH_CHECK("tcts_pack");
  return CRN_OK;
}

// ---- forward of ConvTranspose3d k=7 s=2 p=3 with 8 * Cout <= 16 (the FG_BG logits layer), jz taps stacked into N
extern "C" int64_t crn_tctsf_packed_floats(int32_t Cin) { return (int64_t)((Cin + 7) / 8) * 4 * (Geo<4>::WROW_BYTES / 4); }

extern "C" int crn_tctsf_pack(const float* w, int32_t Cin, int32_t Cout, float* out, void* stream) {
  CRN_REQUIRE(w && out && Cout > 0 && Cin > 0 && 8 * Cout <= 16, "crn_tctsf_pack: Cout <= 2");
  const int P = (Cin + 7) / 8;
  cudaMemsetAsync(out, 0, (size_t)crn_tctsf_packed_floats(Cin) * 4, crn_stream(stream));   // rows 16..31 of WIDE slots unused
  const long long total = (long long)P * 16 * 2 * 4 * 16 * 4;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  tctsf_pack_kernel<<<blocks, 256, 0, crn_stream(stream)>>>(w, Cin, Cout, P, out);
  CRN_LAUNCH_CHECK("tctsf_pack");
  return CRN_OK;
}

extern "C" int crn_convt7_tcs_fwd(const crn_conv_desc* d, const float* x, const float* wtc, const float* bias, float* y,
                                  int32_t* status, void* stream) {
  CRN_REQUIRE(d && x && wtc && y && status, "crn_convt7_tcs_fwd: null pointer");
  CRN_REQUIRE(d->transposed && d->kD == 7 && d->kH == 7 && d->kW == 7 && d->stride == 2 && d->pad == 3,
              "crn_convt7_tcs_fwd: only ConvTranspose3d k=7 s=2 p=3");
  CRN_REQUIRE(d->oD == 2 * d->iD && d->oH == 2 * d->iH && d->oW == 2 * d->iW, "crn_convt7_tcs_fwd: output must be 2x input");
  CRN_REQUIRE(d->iW % TX == 0 && d->iH % TY == 0 && d->iD % ZT == 0 && d->iD >= 4,
              "crn_convt7_tcs_fwd: input grid must tile by 8x16x4");
  CRN_REQUIRE(!d->bias_n_stride && d->Cin % 4 == 0 && d->x_cs % 4 == 0 && d->x_co % 4 == 0 && 8 * d->Cout <= 16,
              "crn_convt7_tcs_fwd: Cin, x strides multiples of 4, Cout <= 2");
  TC5SParams p{};
  p.single = crn_single_pass();
  p.dbg = (crn_get_flags() >> 8) & 3;    // bit 0: wait accounting, bit 1: skip the MMAs (issue-overhead probe)
  p.in = x; p.wtc = wtc; p.bias = bias; p.out = y; p.status = status;
  p.N = d->N; p.D = d->iD; p.H = d->iH; p.W = d->iW;
  p.gK = d->Cin; p.gN = 8 * d->Cout; p.cout_cls = d->Cout; p.planar = d->y_planar;
  p.in_cs = d->x_cs; p.in_co = d->x_co; p.out_cs = d->y_cs; p.out_co = d->y_co;
  p.P = (p.gK + 7) / 8;
  p.tiles_x = p.W / TX; p.tiles_y = p.H / TY; p.tiles_z = p.D / ZT;
  p.nitems = p.N * p.tiles_x * p.tiles_y * p.tiles_z;
  return launch_tc5s<false, 4, false, true>(p, crn_stream(stream));
}

extern "C" int crn_convt7_tcs_dgrad(const crn_conv_desc* d, const float* dy, const float* wtc, float* dx,
                                    int32_t* status, void* stream) {
  CRN_REQUIRE(d && dy && wtc && dx && status, "crn_convt7_tcs_dgrad: null pointer");
  CRN_REQUIRE(d->transposed && d->kD == 7 && d->kH == 7 && d->kW == 7 && d->stride == 2 && d->pad == 3,
              "crn_convt7_tcs_dgrad: only ConvTranspose3d k=7 s=2 p=3");
  CRN_REQUIRE(d->oD == 2 * d->iD && d->oH == 2 * d->iH && d->oW == 2 * d->iW, "crn_convt7_tcs_dgrad: output must be 2x input");
  CRN_REQUIRE(d->iW % TX == 0 && d->iH % TY == 0 && d->iD % ZT == 0 && d->iD >= 4,
              "crn_convt7_tcs_dgrad: input grid must tile by 8x16x4");
  CRN_REQUIRE(d->Cin % 4 == 0 && d->Cin <= 32 && d->x_cs % 4 == 0 && d->x_co % 4 == 0,
              "crn_convt7_tcs_dgrad: Cin % 4, Cin <= 32, dx channel stride/offset multiples of 4");
  if (d->y_planar) {
    CRN_REQUIRE(d->Cout == 2, "crn_convt7_tcs_dgrad: a planar dy must have 2 channels");
  } else {
    CRN_REQUIRE(d->Cout % 4 == 0 && d->y_cs % 4 == 0 && d->y_co % 4 == 0,
                "crn_convt7_tcs_dgrad: channels-last dy needs Cout, stride, offset multiples of 4");
  }
  TC5SParams p{};
  p.single = crn_single_pass();
  p.dbg = (crn_get_flags() >> 8) & 3;    // bit 0: wait accounting, bit 1: skip the MMAs (issue-overhead probe)
  p.in = dy; p.wtc = wtc; p.bias = nullptr; p.out = dx; p.status = status;
  p.N = d->N; p.D = d->iD; p.H = d->iH; p.W = d->iW;
  p.gK = 8 * d->Cout; p.gN = d->Cin; p.cout_cls = d->Cout; p.in_planar = d->y_planar;
  p.in_cs = d->y_cs; p.in_co = d->y_co; p.out_cs = d->x_cs; p.out_co = d->x_co;
  p.P = (p.gK + 7) / 8;
  p.tiles_x = p.W / TX; p.tiles_y = p.H / TY; p.tiles_z = p.D / ZT;
  p.nitems = p.N * p.tiles_x * p.tiles_y * p.tiles_z;
  cudaStream_t st = crn_stream(stream);
  return p.gN <= 16 ? launch_tc5s<false, 4, true>(p, st) : launch_tc5s<true, 4, true>(p, st);
}

extern "C" int crn_tc5s_pack2(const float* w, int32_t Cout, int32_t Cin, int32_t dgrad, float* out, void* stream) {
  CRN_REQUIRE(w && out && Cout > 0 && Cin > 0, "crn_tc5s_pack2: bad args");
  const int K = dgrad ? Cout : Cin, N = dgrad ? Cin : Cout;
  CRN_REQUIRE(N <= 32, "crn_tc5s_pack2: N > 32 unsupported");
  const int P = (K + 7) / 8;
  const long long total = (long long)P * 25 * 2 * 5 * 32 * 4;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  tc5s_pack_kernel<<<blocks, 256, 0, crn_stream(stream)>>>(w, Cout, Cin, dgrad ? 1 : 0, N > 16 ? 1 : 0, P, out);
  CRN_LAUNCH_CHECK("tc5s_pack");
  return CRN_OK;
}

#ifdef CRN_DIAG
extern "C" int crn_tc5s_debug_read(long long* host_dst, int32_t n);
extern "C" int crn_tc5s_debug_read(long long* host_dst, int32_t n) {
  return cudaMemcpyFromSymbol(host_dst, g_tc5s_dbg, sizeof(long long) * (n > kNumSMs * 8 ? kNumSMs * 8 : n)) == cudaSuccess
             ? CRN_OK : CRN_ERR_LAUNCH;
}
#endif  // CRN_DIAG

extern "C" int crn_tc5s_pack(const float* w, int32_t Cout, int32_t Cin, float* out, void* stream) {
  CRN_REQUIRE(Cout <= 16, "crn_tc5s_pack: Cout <= 16 (use crn_tc5s_pack2)");
  return crn_tc5s_pack2(w, Cout, Cin, 0, out, stream);
}

// kind 0: y = conv5(x) + bias (gK = Cin, gN = Cout); kind 1: dx = conv5^T(dy) (gK = Cout, gN = Cin); gN <= 32.
// The grid must tile by 8 (x) x 16 (y) x 4 (z).
extern "C" int crn_conv5_tcs2(const crn_conv_desc* d, int32_t kind, const float* in, const float* wtc, const float* bias,
                              float* out, int32_t* status, void* stream) {
  CRN_REQUIRE(d && in && wtc && out && status, "crn_conv5_tcs: null pointer");
  CRN_REQUIRE(!d->transposed && d->kD == 5 && d->kH == 5 && d->kW == 5 && d->stride == 1 && d->pad == 2,
              "crn_conv5_tcs: only Conv3d k=5 s=1 p=2");
  CRN_REQUIRE(d->iD == d->oD && d->iH == d->oH && d->iW == d->oW, "crn_conv5_tcs: shape mismatch");
  CRN_REQUIRE(d->iW % TX == 0 && d->iH % TY == 0 && d->iD % ZT == 0 && d->iD >= 5, "crn_conv5_tcs: grid must tile by 8x16x4");
  CRN_REQUIRE(!d->y_planar && !d->bias_n_stride && d->Cout % 4 == 0 && d->Cin % 4 == 0,
              "crn_conv5_tcs: channels multiples of 4, channels-last output");
  CRN_REQUIRE(d->x_cs % 4 == 0 && d->x_co % 4 == 0 && d->y_cs % 4 == 0 && d->y_co % 4 == 0,
              "crn_conv5_tcs: channel strides/offsets must be multiples of 4");
  TC5SParams p{};
  p.single = crn_single_pass();
  p.dbg = (crn_get_flags() >> 8) & 3;    // bit 0: wait accounting, bit 1: skip the MMAs (issue-overhead probe)
  p.in = in; p.wtc = wtc; p.bias = kind == 0 ? bias : nullptr; p.out = out; p.status = status;
  p.N = d->N; p.D = d->iD; p.H = d->iH; p.W = d->iW;
  if (kind == 0) {
    p.gK = d->Cin; p.gN = d->Cout; p.in_cs = d->x_cs; p.in_co = d->x_co; p.out_cs = d->y_cs; p.out_co = d->y_co;
  } else {
    p.gK = d->Cout; p.gN = d->Cin; p.in_cs = d->y_cs; p.in_co = d->y_co; p.out_cs = d->x_cs; p.out_co = d->x_co;
  }
  CRN_REQUIRE(p.gN <= 32, "crn_conv5_tcs: at most 32 output channels");
  p.P = (p.gK + 7) / 8;
  p.tiles_x = p.W / TX; p.tiles_y = p.H / TY; p.tiles_z = p.D / ZT;
  p.nitems = p.N * p.tiles_x * p.tiles_y * p.tiles_z;
  cudaStream_t st = crn_stream(stream);
  return p.gN <= 16 ? launch_tc5s<false>(p, st) : launch_tc5s<true>(p, st);
}

extern "C" int crn_conv5_tcs(const crn_conv_desc* d, const float* x, const float* wtc, const float* bias, float* y,
                             int32_t* status, void* stream) {
  CRN_REQUIRE(d && d->Cout <= 16, "crn_conv5_tcs: Cout <= 16 (use crn_conv5_tcs2)");
  return crn_conv5_tcs2(d, 0, x, wtc, bias, y, status, stream);
}
