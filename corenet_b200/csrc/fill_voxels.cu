// fill_inside_voxels for sm_100a: bit-packed flood fill from the near faces.
//
// Replaces cc/fill_voxels_gpu.cu:136-171 (kernels merge_neighbour_regions
// :96-120 and compress_paths :122-132) and cc/fill_voxels_cpu.cc:158-183.
// Same result, different algorithm: the reference labels every voxel with a
// global-memory union-find (16 B/voxel scratch); here
//   A. pack:   grid -> empty-mask bits E (1 bit/voxel) + seed bits R = E on the
//              x=0 / y=0 / z=0 faces (near-face rule)      [HBM read]
//   B. flood:  R <- fixpoint of "R spreads through E" using bit-parallel row
//              fills (carry-chain trick, 128 voxels per add) and register-
//              carried sweeps along +-y and +-z; loops until nothing changes,
//              so the result is the exact connected component      [L2 resident]
//   C. unpack: out = R ? 0 : 1 in the element type                   [HBM write]
// Algorithmic bytes: read T + write T per voxel; scratch is 0.25 B/voxel.
#include "common.cuh"

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
  if (nw > MAXNW) {
    crn_set_error("crn_fill_inside: W=%d > %d unsupported", W, MAXNW * 32);
    return CRN_ERR_UNSUPPORTED;
  }
  cudaStream_t st = crn_stream(stream);
  const int64_t rows = (int64_t)N * D * H;
  uint32_t* E = reinterpret_cast<uint32_t*>(workspace);
  uint32_t* R = E + rows * nw;
  fill_pack_kernel<<<grid_for(rows * nw * 32), NT, 0, st>>>(grid_in, elem_size, dtype_kind, rows, D, H, W, nw,
                                                           E, R);
  int threads = D > H ? D : H;
  threads = ((threads + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  switch (nw) {
    case 1: fill_flood_kernel<1><<<N, threads, 0, st>>>(E, R, D, H); break;
    case 2: fill_flood_kernel<2><<<N, threads, 0, st>>>(E, R, D, H); break;
    case 3: fill_flood_kernel<3><<<N, threads, 0, st>>>(E, R, D, H); break;
    case 4: fill_flood_kernel<4><<<N, threads, 0, st>>>(E, R, D, H); break;
    case 5: fill_flood_kernel<5><<<N, threads, 0, st>>>(E, R, D, H); break;
    case 6: fill_flood_kernel<6><<<N, threads, 0, st>>>(E, R, D, H); break;
    case 7: fill_flood_kernel<7><<<N, threads, 0, st>>>(E, R, D, H); break;
    default: fill_flood_kernel<8><<<N, threads, 0, st>>>(E, R, D, H); break;
  }
  fill_unpack_kernel<<<grid_for(rows * W), NT, 0, st>>>(R, elem_size, dtype_kind, rows, W, nw, grid_out);
  crn_count_launches(2);
  CRN_LAUNCH_CHECK("fill_inside");
  return CRN_OK;
}
