// fill_inside_voxels for sm_100a: bit-packed flood fill from the near faces.
//
// Replaces cc/fill_voxels_gpu.cu:136-171 (kernels merge_neighbour_regions
// :96-120 and compress_paths :122-132) and cc/fill_voxels_cpu.cc:158-183.
// Same result, different algorithm: the reference labels every voxel with a
// global-memory union-find (16 B/voxel scratch); here
//   A. pack:   grid -> empty-mask bits E (1 bit/voxel) + seed bits R = E on the
//              x=0 / y=0 / z=0 faces (near-face rule)      [HBM read]
//   B. flood:  R <- fixpoint of "R spreads through E", so the result is the exact
//              connected component.  Grids whose bit planes fit the shared memory
//              of a thread-block cluster (up to 8 CTAs x 192 KB: 128^3 .. 160^3)
//              run fill_flood_cluster_kernel: each CTA keeps a z-slab of E and R in
//              shared memory; x spreads by carry chains (128 voxels per add), y and z
//              by register-carried line sweeps over whole bit rows (the z lines cross
//              the cluster slab by slab through distributed shared memory).  Larger
//              grids fall back to the same sweeps from global memory (any W).
//   C. unpack: out = R ? 0 : 1 in the element type                   [HBM write]
// Algorithmic bytes: read T + write T per voxel; scratch is 0.25 B/voxel.
#include "common.cuh"
#include <cooperative_groups.h>

namespace {
constexpr int NT = 256;
constexpr int MAXNW = 8;   // row words kept in registers: W <= 256

template <typename T>
__device__ __forceinline__ bool occupied(const void* p, int64_t i) {
  return reinterpret_cast<const T*>(p)[i] > (T)0;
}

// kind: 0 signed int, 1 float, 2 unsigned int
__device__ __forceinline__ bool load_occ(const void* p, int64_t i, int elem_size, int kind) {
  if (kind == 1) return elem_size == 4 ? occupied<float>(p, i) : occupied<double>(p, i);
  if (kind == 2) {
    switch (elem_size) {
      case 1: return occupied<uint8_t>(p, i);
      case 2: return occupied<uint16_t>(p, i);
      case 4: return occupied<uint32_t>(p, i);
      default: return occupied<uint64_t>(p, i);
    }
  }
  switch (elem_size) {
    case 1: return occupied<int8_t>(p, i);
    case 2: return occupied<int16_t>(p, i);
    case 4: return occupied<int32_t>(p, i);
    default: return occupied<int64_t>(p, i);
  }
}

__device__ __forceinline__ void store_val(void* p, int64_t i, int elem_size, int kind, int v) {
  if (kind == 1) {
    if (elem_size == 4) reinterpret_cast<float*>(p)[i] = (float)v;
    else reinterpret_cast<double*>(p)[i] = (double)v;
    return;
  }
  switch (elem_size) {
    case 1: reinterpret_cast<uint8_t*>(p)[i] = (uint8_t)v; break;
    case 2: reinterpret_cast<uint16_t*>(p)[i] = (uint16_t)v; break;
    case 4: reinterpret_cast<uint32_t*>(p)[i] = (uint32_t)v; break;
    default: reinterpret_cast<uint64_t*>(p)[i] = (uint64_t)v; break;
  }
}

// one warp per 32 consecutive x positions
__global__ void __launch_bounds__(NT) fill_pack_kernel(const void* __restrict__ grid, int elem_size, int kind,
                                                       int64_t rows /*N*D*H*/, int D, int H, int W, int nw,
                                                       uint32_t* __restrict__ E, uint32_t* __restrict__ R) {
  const int lane = threadIdx.x & 31;
  const int64_t nwords = rows * nw;
  for (int64_t wd = ((int64_t)blockIdx.x * NT + threadIdx.x) >> 5; wd < nwords;
       wd += ((int64_t)gridDim.x * NT) >> 5) {
    const int64_t row = wd / nw;
    const int wi = (int)(wd - row * nw);
    const int x = wi * 32 + lane;
    bool empty = false;
    if (x < W) empty = !load_occ(grid, row * W + x, elem_size, kind);
    const uint32_t e = __ballot_sync(0xffffffffu, empty);
    if (lane == 0) {
      const int y = (int)(row % H);
      const int z = (int)((row / H) % D);
      E[wd] = e;
      R[wd] = (z == 0 || y == 0) ? e : (wi == 0 ? (e & 1u) : 0u);
    }
  }
}

__global__ void __launch_bounds__(NT) fill_unpack_kernel(const uint32_t* __restrict__ R, int elem_size, int kind,
                                                         int64_t rows, int W, int nw, void* __restrict__ out) {
  const int64_t total = rows * W;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int64_t row = i / W;
    const int x = (int)(i - row * W);
    const uint32_t r = __ldg(R + row * nw + (x >> 5));
    store_val(out, i, elem_size, kind, ((r >> (x & 31)) & 1u) ? 0 : 1);
  }
}

// Fast paths for 4-byte elements and W % 128 == 0 (the pipeline's 128^3 float32 / int32 grids): a lane moves 16 B
// (4 voxels), a warp one 128-voxel chunk = 4 mask words, UNR chunks in flight per warp.
template <bool IS_FLOAT, int UNR>
__global__ void __launch_bounds__(NT) fill_pack4_kernel(const uint4* __restrict__ grid, int64_t chunks /*rows*W/128*/,
                                                        int D, int H, int cpr /*chunks per row*/,
                                                        uint32_t* __restrict__ E, uint32_t* __restrict__ R) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * NT + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * NT) >> 5;
  for (int64_t c0 = warp0 * UNR; c0 < chunks; c0 += nwarps * UNR) {
    uint4 v[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (c0 + u < chunks) v[u] = __ldg(grid + (c0 + u) * 32 + lane);
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (c0 + u >= chunks) break;
      uint32_t nib;
      if (IS_FLOAT) {
        nib = (!(__uint_as_float(v[u].x) > 0.f) ? 1u : 0u) | (!(__uint_as_float(v[u].y) > 0.f) ? 2u : 0u) |
              (!(__uint_as_float(v[u].z) > 0.f) ? 4u : 0u) | (!(__uint_as_float(v[u].w) > 0.f) ? 8u : 0u);
      } else {
        nib = (!((int)v[u].x > 0) ? 1u : 0u) | (!((int)v[u].y > 0) ? 2u : 0u) | (!((int)v[u].z > 0) ? 4u : 0u) |
              (!((int)v[u].w > 0) ? 8u : 0u);
      }
      uint32_t w = nib << (4 * (lane & 7));            // 8 lanes = 32 voxels = one word
      w |= __shfl_xor_sync(0xffffffffu, w, 1);
      w |= __shfl_xor_sync(0xffffffffu, w, 2);
      w |= __shfl_xor_sync(0xffffffffu, w, 4);
      if ((lane & 7) == 0) {
        const int64_t chunk = c0 + u;
        const int64_t row = chunk / cpr;
        const int wi = (int)(chunk - row * cpr) * 4 + (lane >> 3);
        const int y = (int)(row % H);
        const int z = (int)((row / H) % D);
        const int64_t wd = row * (cpr * 4) + wi;
        E[wd] = w;
        R[wd] = (z == 0 || y == 0) ? w : (wi == 0 ? (w & 1u) : 0u);
      }
    }
  }
}

template <bool IS_FLOAT>
__global__ void __launch_bounds__(NT) fill_unpack4_kernel(const uint32_t* __restrict__ R, int64_t quads /*voxels/4*/,
                                                          uint4* __restrict__ out) {
  const uint32_t one = IS_FLOAT ? 0x3f800000u : 1u;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < quads; i += (int64_t)gridDim.x * NT) {
    const uint32_t r = __ldg(R + (i >> 3)) >> (4 * (int)(i & 7));      // W % 128 == 0: words never straddle rows
    uint4 o;
    o.x = (r & 1u) ? 0u : one; o.y = (r & 2u) ? 0u : one; o.z = (r & 4u) ? 0u : one; o.w = (r & 8u) ? 0u : one;
    out[i] = o;
  }
}

// r <- all bits of E-runs that contain a bit of r  (r subset of E), multiword.
template <int NW>
__device__ __forceinline__ void row_fill(uint32_t (&r)[NW], const uint32_t (&e)[NW]) {
  uint32_t up[NW];
  uint64_t carry = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) {          // towards higher x: carry chain of (E + r)
    const uint64_t s = (uint64_t)e[w] + (uint64_t)r[w] + carry;
    carry = s >> 32;
    up[w] = r[w] | (e[w] & ((uint32_t)s ^ e[w]));
  }
  carry = 0;
#pragma unroll
  for (int w = NW - 1; w >= 0; --w) {     // towards lower x: same on bit-reversed words
    const uint32_t er = __brev(e[w]), rr = __brev(r[w]);
    const uint64_t s = (uint64_t)er + (uint64_t)rr + carry;
    carry = s >> 32;
    const uint32_t dn = __brev(rr | (er & ((uint32_t)s ^ er)));
    r[w] = up[w] | dn;
  }
}

// One CTA per mesh grid.  Thread t owns line t of the sweep's parallel axis.
template <int NW>
__global__ void __launch_bounds__(1024) fill_flood_kernel(const uint32_t* __restrict__ E, uint32_t* R, int D,
                                                          int H) {
  const int64_t base = (int64_t)blockIdx.x * D * H * NW;
  const uint32_t* Eg = E + base;
  uint32_t* Rg = R + base;
  const int t = threadIdx.x;
  auto sweep = [&](int line, int len, int64_t line_stride, int64_t step_stride, bool reverse) -> bool {
    // walks `len` rows of line `line`; row address = line*line_stride + i*step_stride (in rows)
    bool changed = false;
    uint32_t prev[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) prev[w] = 0u;
    for (int k = 0; k < len; ++k) {
      const int i = reverse ? len - 1 - k : k;
      const int64_t off = ((int64_t)line * line_stride + (int64_t)i * step_stride) * NW;
      uint32_t e[NW], r[NW], r0[NW];
#pragma unroll
      for (int w = 0; w < NW; ++w) { e[w] = __ldg(Eg + off + w); r0[w] = Rg[off + w]; }
      bool grew = false;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        r[w] = r0[w] | (e[w] & prev[w]);
        grew |= (r[w] != r0[w]);
      }
      if (grew) {
        row_fill<NW>(r, e);
#pragma unroll
        for (int w = 0; w < NW; ++w) Rg[off + w] = r[w];
        changed = true;
      }
#pragma unroll
      for (int w = 0; w < NW; ++w) prev[w] = r[w];
    }
    return changed;
  };

  // initial in-row fill of the seeds (x = 0 seeds spread along their rows)
  for (int row = t; row < D * H; row += blockDim.x) {
    uint32_t e[NW], r[NW];
    const int64_t off = (int64_t)row * NW;
#pragma unroll
    for (int w = 0; w < NW; ++w) { e[w] = __ldg(Eg + off + w); r[w] = Rg[off + w]; }
    row_fill<NW>(r, e);
#pragma unroll
    for (int w = 0; w < NW; ++w) Rg[off + w] = r[w];
  }
  __syncthreads();
  for (;;) {
    bool changed = false;
    // along y (rows z*H + y): one thread per z
    for (int z = t; z < D; z += blockDim.x) changed |= sweep(z, H, H, 1, false);
    __syncthreads();
    // along z (rows z*H + y): one thread per y
    for (int y = t; y < H; y += blockDim.x) changed |= sweep(y, D, 1, H, false);
    __syncthreads();
    for (int z = t; z < D; z += blockDim.x) changed |= sweep(z, H, H, 1, true);
    __syncthreads();
    for (int y = t; y < H; y += blockDim.x) changed |= sweep(y, D, 1, H, true);
    if (!__syncthreads_or(changed ? 1 : 0)) break;
  }
}

// Any row width: the same line sweeps with the row words streamed from global memory (W > 256).
__global__ void __launch_bounds__(1024) fill_flood_wide_kernel(const uint32_t* __restrict__ E, uint32_t* R, int D,
                                                               int H, int nw) {
  const int64_t base = (int64_t)blockIdx.x * D * H * nw;
  const uint32_t* Eg = E + base;
  uint32_t* Rg = R + base;
  const int t = threadIdx.x;
  // in-row fill of row `off` given that R was OR-ed with `prev & E` already; returns true if the row changed
  auto fill_row = [&](int64_t off, const uint32_t* prev_row) -> bool {
    bool grew = false;
    for (int w = 0; w < nw; ++w) {
      const uint32_t e = __ldg(Eg + off + w), r0 = Rg[off + w];
      const uint32_t r = r0 | (prev_row ? (e & prev_row[w]) : 0u);
      if (r != r0) { Rg[off + w] = r; grew = true; }
    }
    if (!grew && prev_row) return false;
    uint64_t carry = 0;
    bool any = grew;
    for (int w = 0; w < nw; ++w) {
      const uint32_t e = __ldg(Eg + off + w), r = Rg[off + w];
      const uint64_t s = (uint64_t)e + (uint64_t)r + carry;
      carry = s >> 32;
      const uint32_t up = r | (e & ((uint32_t)s ^ e));
      if (up != r) { Rg[off + w] = up; any = true; }
    }
    carry = 0;
    for (int w = nw - 1; w >= 0; --w) {
      const uint32_t e = __ldg(Eg + off + w), r = Rg[off + w];
      const uint32_t er = __brev(e), rr = __brev(r);
      const uint64_t s = (uint64_t)er + (uint64_t)rr + carry;
      carry = s >> 32;
      const uint32_t dn = __brev(rr | (er & ((uint32_t)s ^ er)));
      if (dn != r) { Rg[off + w] = dn; any = true; }
    }
    return any;
  };
  for (int row = t; row < D * H; row += blockDim.x) fill_row((int64_t)row * nw, nullptr);
  __syncthreads();
  auto sweep = [&](int line, int len, int64_t line_stride, int64_t step_stride, bool reverse) -> bool {
    bool changed = false;
    const uint32_t* prev = nullptr;
    for (int k = 0; k < len; ++k) {
      const int i = reverse ? len - 1 - k : k;
      const int64_t off = ((int64_t)line * line_stride + (int64_t)i * step_stride) * nw;
      if (prev) changed |= fill_row(off, prev);
      prev = Rg + off;
    }
    return changed;
  };
  for (;;) {
    bool changed = false;
    for (int z = t; z < D; z += blockDim.x) changed |= sweep(z, H, H, 1, false);
    __syncthreads();
    for (int y = t; y < H; y += blockDim.x) changed |= sweep(y, D, 1, H, false);
    __syncthreads();
    for (int z = t; z < D; z += blockDim.x) changed |= sweep(z, H, H, 1, true);
    __syncthreads();
    for (int y = t; y < H; y += blockDim.x) changed |= sweep(y, D, 1, H, true);
    if (!__syncthreads_or(changed ? 1 : 0)) break;
  }
}

// ---- shared-memory / cluster flood ---------------------------------------------------------------------------
// One cluster of CS CTAs per grid; CTA `rank` owns the z-slab [rank*zs, rank*zs + zs) and keeps its E (empty) and R
// (reached) bit planes in shared memory ([zs][H][NW] words each).  All propagation is by register-carried LINE SWEEPS
// over whole 32*NW-bit rows: a sweep walks a line of rows, ORs `E & previous row` into the row and, if the row grew,
// re-closes it along x with the carry chains (row_fill), so one sweep carries the flood around corners in its plane.
//   * in-plane closure (CTA-local): one thread per z-plane sweeps +y then -y until nothing changes -- every plane has
//     its own seeds (the y = 0 row and the x = 0 column), so nearly all of the outside region is reached here;
//   * z sweeps: one thread per (y) row line walks the planes; the line crosses the cluster slab by slab: stage s of
//     the +z sweep runs in CTA s and takes its incoming row from the last plane of CTA s-1 through distributed shared
//     memory (cluster.sync between stages), then the same downwards;
//   * repeated until a z phase reaches nothing new in any CTA.
constexpr int CL_NT = 256;
constexpr int CL_MAXW = 24576;                   // words per slab array: 2 arrays x 96 KB
namespace cg = cooperative_groups;

template <int NW>
__global__ void __launch_bounds__(CL_NT) fill_flood_cluster_kernel(const uint32_t* __restrict__ E, uint32_t* R, int D,
                                                                   int H, int zs) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int64_t gbase = (int64_t)(blockIdx.x / CS) * D * H * NW;
  extern __shared__ uint32_t sm[];
  const int cap = zs * H * NW;                   // words per slab array (same in every CTA)
  uint32_t* Es = sm;
  uint32_t* Rs = Es + cap;
  __shared__ int flags[2];
  const int z0 = rank * zs;
  const int nz = max(0, min(zs, D - z0));
  const int words = nz * H * NW;
  const int rows = nz * H;
  const int tid = threadIdx.x;
  for (int i = tid; i < words; i += CL_NT) {
    Es[i] = __ldg(E + gbase + (int64_t)z0 * H * NW + i);
    Rs[i] = R[gbase + (int64_t)z0 * H * NW + i];
  }
  if (tid < 2) flags[tid] = 0;
  __syncthreads();

  // walks `count` rows starting at row index `row` with stride `step` (in rows); prev = reached bits of the row before
  auto sweep = [&](int row, int count, int step, uint32_t (&prev)[NW]) -> bool {
    bool ch = false;
    for (int k = 0; k < count; ++k, row += step) {
      uint32_t e[NW], r[NW];
      bool grew = false;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        e[w] = Es[row * NW + w];
        const uint32_t r0 = Rs[row * NW + w];
        r[w] = r0 | (e[w] & prev[w]);
        grew |= r[w] != r0;
      }
      if (grew) {
        row_fill<NW>(r, e);
#pragma unroll
        for (int w = 0; w < NW; ++w) Rs[row * NW + w] = r[w];
        ch = true;
      }
#pragma unroll
      for (int w = 0; w < NW; ++w) prev[w] = r[w];
    }
    return ch;
  };

  // seeds spread along their rows
  for (int row = tid; row < rows; row += CL_NT) {
    uint32_t e[NW], r[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) { e[w] = Es[row * NW + w]; r[w] = Rs[row * NW + w]; }
    row_fill<NW>(r, e);
#pragma unroll
    for (int w = 0; w < NW; ++w) Rs[row * NW + w] = r[w];
  }
  __syncthreads();
  for (int it = 0;; ++it) {
    // ---- in-plane closure: thread zl sweeps plane zl along +y, then -y, until no plane changes
    for (;;) {
      bool ch = false;
      for (int zl = tid; zl < nz; zl += CL_NT) {
        uint32_t prev[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) prev[w] = 0u;
        ch |= sweep(zl * H, H, 1, prev);
#pragma unroll
        for (int w = 0; w < NW; ++w) prev[w] = 0u;
        ch |= sweep(zl * H + H - 1, H, -1, prev);
      }
      if (!__syncthreads_or(ch ? 1 : 0)) break;
    }
    // ---- z sweeps across the cluster, slab by slab
    bool changed = false;
    cluster.sync();                              // every slab's in-plane closure is visible to its neighbours
    for (int s = 0; s < CS; ++s) {               // +z
      if (rank == s && nz > 0) {
        const uint32_t* below = rank > 0 ? cluster.map_shared_rank(Rs, rank - 1) + (size_t)(zs - 1) * H * NW : nullptr;
        for (int y = tid; y < H; y += CL_NT) {
          uint32_t prev[NW];
#pragma unroll
          for (int w = 0; w < NW; ++w) prev[w] = below ? below[y * NW + w] : 0u;
          changed |= sweep(y, nz, H, prev);
        }
      }
      cluster.sync();
    }
    for (int s = CS - 1; s >= 0; --s) {          // -z
      if (rank == s && nz > 0) {
        const bool has_above = rank + 1 < CS && z0 + zs < D;
        const uint32_t* above = has_above ? cluster.map_shared_rank(Rs, rank + 1) : nullptr;
        for (int y = tid; y < H; y += CL_NT) {
          uint32_t prev[NW];
#pragma unroll
          for (int w = 0; w < NW; ++w) prev[w] = above ? above[y * NW + w] : 0u;
          changed |= sweep((nz - 1) * H + y, nz, -H, prev);
        }
      }
      cluster.sync();
    }
    // ---- did the z phase reach anything new anywhere in the cluster?  (the in-plane closure had converged)
    const int any_local = __syncthreads_or(changed ? 1 : 0);
    if (tid == 0) { flags[it & 1] = any_local; flags[(it + 1) & 1] = 0; }
    cluster.sync();
    int any = 0;
    for (int c = 0; c < CS; ++c) any |= *cluster.map_shared_rank(&flags[it & 1], c);
    if (!any) break;
  }
  cluster.sync();                                // no CTA may exit while its shared memory can still be read
  for (int i = tid; i < words; i += CL_NT) R[gbase + (int64_t)z0 * H * NW + i] = Rs[i];
}

template <int NW>
int launch_cluster_flood(const uint32_t* E, uint32_t* R, int N, int D, int H, cudaStream_t st) {
  // the smallest cluster whose slabs fit the shared memory (2 arrays of <= CL_MAXW words per CTA); at least 8 CTAs per
  // grid when the grid is deep enough: the in-plane sweeps run one thread per plane, so more slabs = more parallel lines
  int cs = 1;
  while (cs < 8 && ((int64_t)((D + cs - 1) / cs) * H * NW > CL_MAXW || (D + cs - 1) / cs > 16)) cs <<= 1;
  if ((int64_t)((D + cs - 1) / cs) * H * NW > CL_MAXW) return 1;      // does not fit: caller falls back
  const int zs = (D + cs - 1) / cs;
  const size_t smem = (size_t)2 * zs * H * NW * sizeof(uint32_t);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(fill_flood_cluster_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * CL_MAXW * 4);
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(N * cs), 1, 1);
  cfg.blockDim = dim3(CL_NT, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, fill_flood_cluster_kernel<NW>, E, R, D, H, zs);
  return e == cudaSuccess ? 0 : -1;
}

inline unsigned grid_for(int64_t total) {
  int64_t b = crn_ceil_div(total, NT);
  if (b > 16LL * kNumSMs) b = 16LL * kNumSMs;
  return (unsigned)(b < 1 ? 1 : b);
}
}  // namespace

extern "C" int64_t crn_fill_workspace_bytes(int32_t N, int32_t D, int32_t H, int32_t W) {
  const int64_t nw = (W + 31) / 32;
  return 2 * (int64_t)N * D * H * nw * (int64_t)sizeof(uint32_t);
}

extern "C" int crn_fill_inside(const void* grid_in, void* grid_out, int32_t elem_size, int32_t dtype_kind,
                               int32_t N, int32_t D, int32_t H, int32_t W, void* workspace, void* stream) {
  CRN_REQUIRE(grid_in && grid_out && workspace, "crn_fill_inside: null pointer");
  CRN_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0, "crn_fill_inside: empty grid");
  CRN_REQUIRE(elem_size == 1 || elem_size == 2 || elem_size == 4 || elem_size == 8,
              "crn_fill_inside: element size must be 1/2/4/8");
  CRN_REQUIRE(dtype_kind >= 0 && dtype_kind <= 2 && !(dtype_kind == 1 && elem_size < 4),
              "crn_fill_inside: bad dtype kind");
  const int nw = (W + 31) / 32;
  cudaStream_t st = crn_stream(stream);
  const int64_t rows = (int64_t)N * D * H;
  uint32_t* E = reinterpret_cast<uint32_t*>(workspace);
  uint32_t* R = E + rows * nw;
  const bool fast4 = elem_size == 4 && dtype_kind != 2 && W % 128 == 0 &&
                     (reinterpret_cast<uintptr_t>(grid_in) & 15) == 0 && (reinterpret_cast<uintptr_t>(grid_out) & 15) == 0;
  if (fast4) {
    const int64_t chunks = rows * (W / 128);
    const unsigned g = grid_for(crn_ceil_div(chunks, 2) * 32);
    if (dtype_kind == 1)
      fill_pack4_kernel<true, 2><<<g, NT, 0, st>>>(reinterpret_cast<const uint4*>(grid_in), chunks, D, H, W / 128, E, R);
    else
      fill_pack4_kernel<false, 2><<<g, NT, 0, st>>>(reinterpret_cast<const uint4*>(grid_in), chunks, D, H, W / 128, E, R);
  } else {
    fill_pack_kernel<<<grid_for(rows * nw * 32), NT, 0, st>>>(grid_in, elem_size, dtype_kind, rows, D, H, W, nw,
                                                             E, R);
  }
  int threads = D > H ? D : H;
  threads = ((threads + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  int rc = 1;                                    // 1 = not handled by the cluster kernel
  if (!(crn_get_flags() & 4096)) {               // bit 12 of crn_set_flags: force the line-sweep kernels (A/B)
    switch (nw) {
      case 1: rc = launch_cluster_flood<1>(E, R, N, D, H, st); break;
      case 2: rc = launch_cluster_flood<2>(E, R, N, D, H, st); break;
      case 3: rc = launch_cluster_flood<3>(E, R, N, D, H, st); break;
      case 4: rc = launch_cluster_flood<4>(E, R, N, D, H, st); break;
      case 5: rc = launch_cluster_flood<5>(E, R, N, D, H, st); break;
      case 6: rc = launch_cluster_flood<6>(E, R, N, D, H, st); break;
      case 7: rc = launch_cluster_flood<7>(E, R, N, D, H, st); break;
      case 8: rc = launch_cluster_flood<8>(E, R, N, D, H, st); break;
      default: break;
    }
  }
  if (rc < 0) {
    crn_set_error("crn_fill_inside: cluster launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return CRN_ERR_LAUNCH;
  }
  if (rc == 1) {
    switch (nw) {
      case 1: fill_flood_kernel<1><<<N, threads, 0, st>>>(E, R, D, H); break;
      case 2: fill_flood_kernel<2><<<N, threads, 0, st>>>(E, R, D, H); break;
      case 3: fill_flood_kernel<3><<<N, threads, 0, st>>>(E, R, D, H); break;
      case 4: fill_flood_kernel<4><<<N, threads, 0, st>>>(E, R, D, H); break;
      case 5: fill_flood_kernel<5><<<N, threads, 0, st>>>(E, R, D, H); break;
      case 6: fill_flood_kernel<6><<<N, threads, 0, st>>>(E, R, D, H); break;
      case 7: fill_flood_kernel<7><<<N, threads, 0, st>>>(E, R, D, H); break;
      case 8: fill_flood_kernel<8><<<N, threads, 0, st>>>(E, R, D, H); break;
      default: fill_flood_wide_kernel<<<N, threads, 0, st>>>(E, R, D, H, nw); break;
    }
  }
  if (fast4) {
    if (dtype_kind == 1)
      fill_unpack4_kernel<true><<<grid_for(rows * W / 4), NT, 0, st>>>(R, rows * W / 4, reinterpret_cast<uint4*>(grid_out));
    else
      fill_unpack4_kernel<false><<<grid_for(rows * W / 4), NT, 0, st>>>(R, rows * W / 4, reinterpret_cast<uint4*>(grid_out));
  } else {
    fill_unpack_kernel<<<grid_for(rows * W), NT, 0, st>>>(R, elem_size, dtype_kind, rows, W, nw, grid_out);
  }
  crn_count_launches(2);
  CRN_LAUNCH_CHECK("fill_inside");
  return CRN_OK;
}
