// Weight gradient of the 3-D decoder convolutions with small channel counts
// (Conv3d k=5 and ConvTranspose3d k=7 s=2; model/reconstruction_decoder.py:57-95).
//
//   conv  (SS=1): dW[kz,ky,kx][ci][co] = sum_pos  X[pos + k - pad][ci]      * dY[pos][co]
//   convT (SS=2): dW[kz,ky,kx][ci][co] = sum_pos  X[pos][ci]                * dY[2*pos + k - pad][co]
//
// The output (taps x Cin x Cout) is tiny and the reduction (all voxels) is huge, so the
// generic split-K GEMM tile is mostly padding.  Here a thread owns a "tap row":
// all KX taps along x for one (kz,ky), 4 channels of the centre operand and SW channels
// of the shifted operand = KX*4*SW accumulators in registers, and walks along x with a
// sliding register window over the shifted operand, so each position costs SS+1 shared
// loads for KX*4*SW FMAs (FFMA-bound, fp32).  Blocks are persistent over (n, z, y-block)
// tiles; operands are staged once per tile in shared memory in their global
// channels-last layout (no transpose); partial sums are flushed with one atomicAdd
// per accumulator per block.
#include "common.cuh"

namespace {

struct WRowParams {
  const float* C;     // centre operand  [N, cD, cH, cW, c_cs] (+ c_co)
  const float* S;     // shifted operand [N, sD, sH, sW, s_cs] (+ s_co)
  float* dw;          // [tap][wK][wN]
  int N, cD, cH, cW, sD, sH, sW;
  int c_cs, c_co, s_cs, s_co;
  int CQ, SQ;         // quads of the centre operand (4 wide), groups of the shifted operand (SW wide)
  int pad, KZ, KY;
  int wK, wN;
  int gM, gN;         // logical Cin, Cout (for the flush bounds)
  int TY, nstream, CQb, NCQG;   // centre rows per step (= position streams), centre quads per block, centre-quad groups
  int SR, SWd, RING;  // shifted rows one step reads, staged width (voxels), rows in the ring (SR + TY*SS)
  int nseg, rows_per_seg;       // y segments per plane (load balance) and their height
  int center_floats;  // floats of ONE centre buffer
  int cq_shift, sv_shift, s_vec4;   // log2(CQb) / log2(float4 units per shifted voxel) or -1; shifted tile copied in 16-B units
  long long nitems;
};

__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool ok) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = ok ? 16 : 0;                      // src-size 0 => 16 B of zeros, nothing read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async8(float* dst, const float* src, bool ok) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int sz = ok ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

template <int SW>
__device__ __forceinline__ void load_group(const float* p, float (&v)[SW]) {
  if constexpr (SW == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  }
}

template <int KX, int SS, int SW, bool SHIFT_IS_CI>
__global__ void __launch_bounds__(KX == 5 ? 160 : 256, KX == 5 ? 3 : 2) wgrad_row_kernel(const WRowParams p) {
  extern __shared__ __align__(16) float smem[];
  float* cS = smem;                              // 2 x [TY][cW][CQb*4]   (double buffer)
  float* sS = smem + 2 * p.center_floats;        // [RING][SWd][SQ*SW]    (ring of shifted rows)
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int roles = p.KY * p.SQ * p.CQb;
  const int rho = tid % roles, sigma = tid / roles;
  const bool active = sigma < p.nstream;
  const int cq = rho % p.CQb, sq = (rho / p.CQb) % p.SQ, ky = rho / (p.CQb * p.SQ);
  const int cqg = blockIdx.y % p.NCQG, kz = blockIdx.y / p.NCQG;
  const int cstride = p.CQb * 4;        // floats per voxel in the centre tile
  const int sstride = p.SQ * SW;        // floats per voxel in the shifted tile
  const int row_floats = p.SWd * sstride;

  float acc[KX][4][SW];
#pragma unroll
  for (int k = 0; k < KX; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < SW; ++j) acc[k][i][j] = 0.f;

  // Work item = (scene, z plane, y segment).  The block walks down y: every step computes TY centre rows
  // (one per position stream) and meanwhile cp.async-prefetches the TY*SS shifted rows + TY centre rows of the
  // next step into the ring, so each shifted row is fetched once per (item, kz) and the copy latency hides
  // behind a row of FMAs.  One __syncthreads per step.
  for (long long item = blockIdx.x; item < p.nitems; item += gridDim.x) {
    const int seg = (int)(item % p.nseg);
    const long long pl = item / p.nseg;
    const int z = (int)(pl % p.cD);
    const int n = (int)(pl / p.cD);
    const int sz = z * SS - p.pad + kz;
    if (sz < 0 || sz >= p.sD) continue;            // block-uniform
    const int ybeg = seg * p.rows_per_seg;
    const int yend = min(p.cH, ybeg + p.rows_per_seg);
    if (ybeg >= yend) continue;
    const int nsteps = (yend - ybeg + p.TY - 1) / p.TY;
    const int sy_origin = ybeg * SS - p.pad;       // ring row g holds shifted row sy_origin + g
    const long long cplane = ((long long)n * p.cD + z) * p.cH;
    const long long splane = ((long long)n * p.sD + sz) * p.sH;

    // centre rows y0 .. y0+TY-1 -> buffer buf (zero-fill past the grid)
    auto stage_centre = [&](int buf, int y0) {
      const int upr = p.cW * p.CQb;                // float4 units per row
      float* dst = cS + (size_t)buf * p.center_floats;
      int yy = tid / upr, c = tid - yy * upr;
      while (yy < p.TY) {
        const int y = y0 + yy;
        const bool ok = y < p.cH;
        int x, qq;
        if (p.cq_shift >= 0) { x = c >> p.cq_shift; qq = c & (p.CQb - 1); } else { x = c / p.CQb; qq = c - x * p.CQb; }
        const long long off = ((cplane + (ok ? y : 0)) * p.cW + x) * p.c_cs + p.c_co + (cqg * p.CQb + qq) * 4;
        cp_async16(dst + ((size_t)yy * upr + c) * 4, p.C + off, ok);
        c += nthr;
        while (c >= upr) { c -= upr; ++yy; }
      }
    };
    // ring rows g0 .. g0+cnt-1 (zero-fill outside the grid = conv padding)
    auto stage_shifted = [&](int g0, int cnt) {
      if (p.s_vec4) {
        const int upv = sstride >> 2;
        const int upr = p.SWd * upv;
        int rr = tid / upr, c = tid - rr * upr;
        while (rr < cnt) {
          const int g = g0 + rr;
          const int sy = sy_origin + g;
          int xx, qq;
          if (p.sv_shift >= 0) { xx = c >> p.sv_shift; qq = c & (upv - 1); } else { xx = c / upv; qq = c - xx * upv; }
          const int sx = xx - p.pad;
          const bool ok = (unsigned)sy < (unsigned)p.sH && (unsigned)sx < (unsigned)p.sW;
          const long long off = ok ? ((splane + sy) * p.sW + sx) * p.s_cs + p.s_co + qq * 4 : 0;
          cp_async16(sS + (size_t)(g % p.RING) * row_floats + (size_t)c * 4, p.S + off, ok);
          c += nthr;
          while (c >= upr) { c -= upr; ++rr; }
        }
      } else {
        const int upv = sstride >> 1;
        const int upr = p.SWd * upv;
        int rr = tid / upr, c = tid - rr * upr;
        while (rr < cnt) {
          const int g = g0 + rr;
          const int sy = sy_origin + g;
          const int xx = c / upv, qq = c - xx * upv;
          const int sx = xx - p.pad;
          const bool ok = (unsigned)sy < (unsigned)p.sH && (unsigned)sx < (unsigned)p.sW;
          const long long off = ok ? ((splane + sy) * p.sW + sx) * p.s_cs + p.s_co + qq * 2 : 0;
          cp_async8(sS + (size_t)(g % p.RING) * row_floats + (size_t)c * 2, p.S + off, ok);
          c += nthr;
          while (c >= upr) { c -= upr; ++rr; }
        }
      }
    };

    __syncthreads();                               // the previous item's last step is fully consumed
    stage_shifted(0, p.SR);
    stage_centre(0, ybeg);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    for (int s = 0; s < nsteps; ++s) {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      __syncthreads();                             // step s landed; step s-1 consumed -> its ring rows are free
      if (s + 1 < nsteps) {
        stage_shifted(p.SR + s * p.TY * SS, p.TY * SS);
        stage_centre((s + 1) & 1, ybeg + (s + 1) * p.TY);
      }
      asm volatile("cp.async.commit_group;\n" ::: "memory");
      const int yy = sigma;                        // TY == nstream: one centre row per stream and step
      const int y = ybeg + s * p.TY + yy;
      if (!active || y >= yend) continue;
      // running pointers (in floats): the unrolled body below only uses compile-time offsets from them
      const float* cptr = cS + (size_t)(s & 1) * p.center_floats + (yy * p.cW) * cstride + cq * 4;
      const float* sptr = sS + (size_t)(((s * p.TY + yy) * SS + ky) % p.RING) * row_floats + sq * SW;
      float w[KX][SW];
#pragma unroll
      for (int k = 0; k < KX; ++k) load_group<SW>(sptr + k * sstride, w[k]);
      sptr += KX * sstride;                       // next shifted element to load
      const int nfull = p.cW / KX;                // full groups of KX positions (no bounds checks inside)
      for (int gi = 0; gi < nfull; ++gi) {
#pragma unroll
        for (int jj = 0; jj < KX; ++jj) {
          const float4 c4 = *reinterpret_cast<const float4*>(cptr + jj * cstride);
          const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
          for (int k = 0; k < KX; ++k) {
            const int slot = (SS * jj + k) % KX;      // compile-time after unrolling
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < SW; ++j) acc[k][i][j] = fmaf(cv[i], w[slot][j], acc[k][i][j]);
          }
          // slide the window by SS elements; the staged row holds SS*(cW-1)+KX elements, so only the load
          // after the very last position could run past it (guarded by the host-side +SS padding columns)
#pragma unroll
          for (int e = 0; e < SS; ++e) load_group<SW>(sptr + (SS * jj + e) * sstride, w[(SS * jj + e) % KX]);
        }
        cptr += KX * cstride;
        sptr += SS * KX * sstride;
      }
      const int rem = p.cW - nfull * KX;          // tail positions (< KX), same code with guards
#pragma unroll
      for (int jj = 0; jj < KX; ++jj) {
        if (jj < rem) {
          const float4 c4 = *reinterpret_cast<const float4*>(cptr + jj * cstride);
          const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
          for (int k = 0; k < KX; ++k) {
            const int slot = (SS * jj + k) % KX;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < SW; ++j) acc[k][i][j] = fmaf(cv[i], w[slot][j], acc[k][i][j]);
          }
          if (jj + 1 < rem) {
#pragma unroll
            for (int e = 0; e < SS; ++e) load_group<SW>(sptr + (SS * jj + e) * sstride, w[(SS * jj + e) % KX]);
          }
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  // ---- flush
  if (active && ky < p.KY) {
#pragma unroll
    for (int k = 0; k < KX; ++k) {
      const int tap = (kz * p.KY + ky) * KX + k;
      float* dwt = p.dw + (long long)tap * p.wK * p.wN;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < SW; ++j) {
          const int cc = (cqg * p.CQb + cq) * 4 + i;     // centre channel
          const int sc = sq * SW + j;                    // shifted channel
          const int ci = SHIFT_IS_CI ? sc : cc;
          const int co = SHIFT_IS_CI ? cc : sc;
          if (ci < p.gM && co < p.gN) atomicAdd(dwt + (long long)ci * p.wN + co, acc[k][i][j]);
        }
    }
  }
}

template <int KX, int SS, int SW, bool SHIFT_IS_CI>
int launch(const WRowParams& p, int threads, size_t smem_bytes, int groups, cudaStream_t st) {
  auto kern = wgrad_row_kernel<KX, SS, SW, SHIFT_IS_CI>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    configured = true;
  }
  // persistent grid: exactly what is co-resident (one extra block would serialise a whole second wave)
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem_bytes) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  const long long slots = ((long long)per_sm * kNumSMs) / groups > 0 ? ((long long)per_sm * kNumSMs) / groups : 1;
  // split planes into y segments while the persistent blocks would otherwise run < 8 rounds of uneven length
  // (each segment re-reads KY-SS halo rows, so keep segments >= 8 steps)
  const long long planes = (long long)p.N * p.cD;
  int nseg = 1;
  while (planes * nseg < 8 * slots && (p.cH / (nseg * 2)) >= 8 * p.TY) nseg *= 2;
  WRowParams q = p;
  q.nseg = nseg;
  q.rows_per_seg = ((p.cH + nseg - 1) / nseg + p.TY - 1) / p.TY * p.TY;
  q.nitems = planes * nseg;
  long long gx = slots < q.nitems ? slots : q.nitems;
  kern<<<dim3((unsigned)gx, (unsigned)groups), threads, smem_bytes, st>>>(q);
  CRN_LAUNCH_CHECK("wgrad_row");
  return CRN_OK;
}

}  // namespace

// Returns CRN_ERR_UNSUPPORTED when the shape is outside this kernel's envelope (the caller falls
// back to the generic split-K kernel).
int crn_wgrad_row_try(const crn_conv_desc* d, const float* x, const float* dy, float* dw, cudaStream_t st) {
  if (crn_get_flags() & 2) return CRN_ERR_UNSUPPORTED;
  // the generic split-K GEMM wins once the Cin x Cout tile is large enough to fill its 64x64 tile
  const bool conv5 = !d->transposed && d->kD == 5 && d->kH == 5 && d->kW == 5 && d->stride == 1 && d->pad == 2;
  const bool convT7 = d->transposed && d->kD == 7 && d->kH == 7 && d->kW == 7 && d->stride == 2 && d->pad == 3;
  if (!conv5 && !convT7) return CRN_ERR_UNSUPPORTED;
  // measured crossover on a B200 (profiles/r01_layers_conv_per_step.txt): conv5 up to 56x32, convT7 up to 64x32
  if (d->Cin * d->Cout >= (convT7 ? 2049 : 2048) && !(crn_get_flags() & 32)) return CRN_ERR_UNSUPPORTED;
  if (d->y_planar) return CRN_ERR_UNSUPPORTED;
  WRowParams p{};
  p.dw = dw; p.N = d->N; p.pad = d->pad; p.KZ = d->kD; p.KY = d->kH;
  p.wK = d->CinP; p.wN = d->CoutP; p.gM = d->Cin; p.gN = d->Cout;
  int SW, Cc, Cs;
  if (conv5) {            // centre = dY (positions of y), shifted = X
    p.C = dy; p.c_cs = d->y_cs; p.c_co = d->y_co; p.cD = d->oD; p.cH = d->oH; p.cW = d->oW; Cc = d->Cout;
    p.S = x; p.s_cs = d->x_cs; p.s_co = d->x_co; p.sD = d->iD; p.sH = d->iH; p.sW = d->iW; Cs = d->Cin;
    SW = 4;
  } else {                // centre = X (positions of x), shifted = dY
    p.C = x; p.c_cs = d->x_cs; p.c_co = d->x_co; p.cD = d->iD; p.cH = d->iH; p.cW = d->iW; Cc = d->Cin;
    p.S = dy; p.s_cs = d->y_cs; p.s_co = d->y_co; p.sD = d->oD; p.sH = d->oH; p.sW = d->oW; Cs = d->Cout;
    SW = 2;
  }
  // reads are done in whole groups: the buffers must hold the rounded-up channel counts
  const int Cc4 = (Cc + 3) / 4 * 4, CsG = (Cs + SW - 1) / SW * SW;
  if (p.c_co + Cc4 > p.c_cs || p.s_co + CsG > p.s_cs) return CRN_ERR_UNSUPPORTED;
  if (p.c_cs % 4 || p.c_co % 4 || p.s_cs % 2 || p.s_co % 2) return CRN_ERR_UNSUPPORTED;
  p.CQ = Cc4 / 4; p.SQ = CsG / SW;
  const int KX = d->kW, SS = conv5 ? 1 : 2;
  // centre quads per block: keep roles = KY*SQ*CQb <= 160
  int CQb = p.CQ;
  while (CQb > 1 && p.KY * p.SQ * CQb > 160) CQb = (CQb + 1) / 2;
  while (p.CQ % CQb) --CQb;
  const int roles = p.KY * p.SQ * CQb;
  if (roles > 320 || p.cW < KX) return CRN_ERR_UNSUPPORTED;
  p.CQb = CQb; p.NCQG = p.CQ / CQb;
  const int max_threads = conv5 ? 160 : 256;     // = the kernel's __launch_bounds__
  if (roles > max_threads) return CRN_ERR_UNSUPPORTED;
  int nstream = (conv5 ? 160 : 224) / roles;
  if (nstream < 1) nstream = 1;
  if (nstream > 8) nstream = 8;
  if (nstream > p.cH) nstream = p.cH;
  size_t smem_bytes = 0;
  for (;;) {
    p.TY = nstream; p.nstream = nstream;
    p.SR = (p.TY - 1) * SS + p.KY;
    p.RING = p.SR + p.TY * SS;
    p.SWd = (p.cW - 1) * SS + KX + SS;       // + SS look-ahead columns read (never used) by the window slide
    p.center_floats = p.TY * p.cW * CQb * 4;
    smem_bytes = sizeof(float) * (2 * (size_t)p.center_floats + (size_t)p.RING * p.SWd * p.SQ * SW);
    if (smem_bytes <= 100 * 1024 || nstream == 1) break;
    nstream = nstream / 2;
  }
  if (smem_bytes > 150 * 1024) return CRN_ERR_UNSUPPORTED;
  auto log2_or_neg = [](int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; };
  p.cq_shift = log2_or_neg(p.CQb);
  p.s_vec4 = ((p.SQ * SW) % 4 == 0 && p.s_cs % 4 == 0 && p.s_co % 4 == 0) ? 1 : 0;
  p.sv_shift = p.s_vec4 ? log2_or_neg(p.SQ * SW / 4) : -1;
  int threads = ((nstream * roles + 31) / 32) * 32;
  if (threads > max_threads) return CRN_ERR_UNSUPPORTED;
  const int groups = p.KZ * p.NCQG;
  p.nseg = 1; p.rows_per_seg = p.cH;           // launch() may split planes into y segments
  if (conv5) return launch<5, 1, 4, true>(p, threads, smem_bytes, groups, st);
  return launch<7, 2, 2, false>(p, threads, smem_bytes, groups, st);
}
