// Surface voxeliser: CUDA replacement for the EGL/GLSL rasteriser path.
//
// Replaces geometry/voxelization.py:115-164 (one GL draw of T points),
// geometry/shaders/voxelize.geom:37-60 (per-triangle view->voxel transform,
// dominant-axis swizzle) and geometry/shaders/voxelize.frag:29-58 (bounds test,
// floor, SSBO store; sub-grid index math), plus the label/max merge at
// data/batched_example.py:186-196.
//
// Restated GL rules: orthographic projection along the
// dominant axis of the face normal (strict comparisons, ties -> no swizzle);
// R x R pixel lattice with R = round(max(W,H,D*pdm)*mult); vertices snapped to
// 1/256 pixel; a fragment for every pixel whose centre is inside the triangle
// (top-left rule on exact edge hits) or -- conservative -- whose square
// overlaps the triangle; attributes interpolated (or extrapolated) at the
// pixel centre; no depth test, no culling.
#include "common.cuh"

namespace {
constexpr int NT = 256;

struct TriSetup {
  float v[3][3];            // voxel-space vertices
  int axA, axB;             // in-plane axes (0=x,1=y,2=z)
  double u[3], w[3];        // window coords (pixels) of the vertices
  long long U[3], V[3];     // snapped, 1/256 px
  long long A[3], B[3], C[3];  // oriented edge functions E = A*px + B*py + C  (fixed point)
  bool tl[3];               // top-left flag per edge
  double inv_area;
  int i0, i1, j0, j1;       // pixel range
  bool valid;
};

__device__ __forceinline__ void setup_triangle(const float* __restrict__ tri, const float* __restrict__ M,
                                               int W, int H, int D, int R, int pdm, bool conservative,
                                               TriSetup& s) {
  s.valid = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float x = tri[k * 3 + 0], y = tri[k * 3 + 1], z = tri[k * 3 + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      // ((m0*x + m1*y) + m2*z) + m3, every step rounded to fp32 (no FMA contraction)
      float t = __fadd_rn(__fmul_rn(M[r * 4 + 0], x), __fmul_rn(M[r * 4 + 1], y));
      t = __fadd_rn(t, __fmul_rn(M[r * 4 + 2], z));
      s.v[k][r] = __fadd_rn(t, M[r * 4 + 3]);
    }
  }
  // face normal (voxelize.geom:44) -- only the ordering of |components| matters
  float e1[3], e2[3];
  float n1 = 0.f, n2 = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    e1[r] = __fsub_rn(s.v[1][r], s.v[0][r]); e2[r] = __fsub_rn(s.v[2][r], s.v[0][r]);
    n1 = __fadd_rn(n1, __fmul_rn(e1[r], e1[r])); n2 = __fadd_rn(n2, __fmul_rn(e2[r], e2[r]));
  }
  n1 = __fsqrt_rn(n1); n2 = __fsqrt_rn(n2);
#pragma unroll
  for (int r = 0; r < 3; ++r) { e1[r] = __fdiv_rn(e1[r], n1); e2[r] = __fdiv_rn(e2[r], n2); }
  const float nx = fabsf(__fsub_rn(__fmul_rn(e1[1], e2[2]), __fmul_rn(e1[2], e2[1])));
  const float ny = fabsf(__fsub_rn(__fmul_rn(e1[2], e2[0]), __fmul_rn(e1[0], e2[2])));
  const float nz = fabsf(__fsub_rn(__fmul_rn(e1[0], e2[1]), __fmul_rn(e1[1], e2[0])));
  int axA = 0, axB = 1;                                  // screen = (X, Y), depth Z
  if (nx > ny && nx > nz) { axA = 1; axB = 2; }          // .yzxw: screen = (Y, Z), depth X
  else if (ny > nx && ny > nz) { axA = 2; axB = 0; }     // .zxyw: screen = (Z, X), depth Y
  s.axA = axA; s.axB = axB;
  const double ext[3] = {(double)W, (double)H, (double)D * pdm};
  long long mnU = LLONG_MAX, mxU = LLONG_MIN, mnV = LLONG_MAX, mxV = LLONG_MIN;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    s.u[k] = __ddiv_rn(__dmul_rn((double)s.v[k][axA], (double)R), ext[axA]);
    s.w[k] = __ddiv_rn(__dmul_rn((double)s.v[k][axB], (double)R), ext[axB]);
    if (!(fabs(s.u[k]) < 1e9) || !(fabs(s.w[k]) < 1e9)) return;   // NaN / absurd
    s.U[k] = llrint(__dmul_rn(s.u[k], 256.0));
    s.V[k] = llrint(__dmul_rn(s.w[k], 256.0));
    mnU = min(mnU, s.U[k]); mxU = max(mxU, s.U[k]);
    mnV = min(mnV, s.V[k]); mxV = max(mxV, s.V[k]);
  }
  const long long area2 = (s.U[1] - s.U[0]) * (s.V[2] - s.V[0]) - (s.V[1] - s.V[0]) * (s.U[2] - s.U[0]);
  if (area2 == 0) return;
  const long long sg = area2 > 0 ? 1 : -1;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int a = k, b = (k + 1) % 3;
    // E(p) = sg * ((Ub-Ua)(py-Va) - (Vb-Va)(px-Ua))
    s.A[k] = -sg * (s.V[b] - s.V[a]);
    s.B[k] = sg * (s.U[b] - s.U[a]);
    s.C[k] = -(s.A[k] * s.U[a] + s.B[k] * s.V[a]);
    s.tl[k] = s.A[k] > 0 || (s.A[k] == 0 && s.B[k] > 0);
  }
  const double area_d = __dsub_rn(__dmul_rn(s.u[1] - s.u[0], s.w[2] - s.w[0]),
                                  __dmul_rn(s.w[1] - s.w[0], s.u[2] - s.u[0]));
  if (area_d == 0.0) return;
  s.inv_area = area_d;   // (kept as the area; the barycentrics divide by it)
  long long i0, i1, j0, j1;
  auto fdiv = [](long long a, long long b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); };  // floor
  auto cdiv = [&](long long a, long long b) { return -fdiv(-a, b); };                            // ceil
  if (conservative) {
    i0 = cdiv(mnU, 256) - 1; i1 = fdiv(mxU, 256);
    j0 = cdiv(mnV, 256) - 1; j1 = fdiv(mxV, 256);
  } else {
    i0 = cdiv(mnU - 128, 256); i1 = fdiv(mxU - 128, 256);
    j0 = cdiv(mnV - 128, 256); j1 = fdiv(mxV - 128, 256);
  }
  i0 = max(i0, 0LL); j0 = max(j0, 0LL);
  i1 = min(i1, (long long)R - 1); j1 = min(j1, (long long)R - 1);
  if (i0 > i1 || j0 > j1) return;
  s.i0 = (int)i0; s.i1 = (int)i1; s.j0 = (int)j0; s.j1 = (int)j1;
  s.valid = true;
}

__device__ __forceinline__ void shade_fragment(const TriSetup& s, int i, int j, int mesh, int W, int H, int D,
                                               int side, float* __restrict__ grid) {
  // attribute interpolation at the pixel centre
  const double su = i + 0.5, sv = j + 0.5;
  const double l1 = __ddiv_rn(__dsub_rn(__dmul_rn(su - s.u[0], s.w[2] - s.w[0]),
                                        __dmul_rn(sv - s.w[0], s.u[2] - s.u[0])), s.inv_area);
  const double l2 = __ddiv_rn(__dsub_rn(__dmul_rn(s.u[1] - s.u[0], sv - s.w[0]),
                                        __dmul_rn(s.w[1] - s.w[0], su - s.u[0])), s.inv_area);
  const double l0 = (1.0 - l1) - l2;
  float p[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    p[r] = (float)__dadd_rn(__dadd_rn(__dmul_rn(l0, (double)s.v[0][r]), __dmul_rn(l1, (double)s.v[1][r])),
                            __dmul_rn(l2, (double)s.v[2][r]));
  // voxelize.frag:36-38
  if (p[0] < 0 || p[1] < 0 || p[2] < 0 || p[0] >= (float)W || p[1] >= (float)H || p[2] >= (float)D) return;
  if (side <= 0) {
    const int cx = (int)floorf(p[0]), cy = (int)floorf(p[1]), cz = (int)floorf(p[2]);
    grid[(((int64_t)mesh * D + cz) * H + cy) * W + cx] = 1.0f;
  } else {
    int c[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int v = (int)floorf(__fmul_rn(p[r], (float)side)) + side / 2;
      c[r] = 2 * (v / side) + ((v % side) == side - 1 ? 1 : 0);
    }
    grid[(((int64_t)mesh * (2 * D + 1) + c[2]) * (2 * H + 1) + c[1]) * (2 * W + 1) + c[0]] = 1.0f;
  }
}

// one warp per triangle; lanes stride over the pixels of its bounding box
// Triangles are staged in SHARED MEMORY: a warp owns one triangle at a time; its lane 0 runs the fixed-point edge /
// snapping set-up (fp64-heavy) ONCE into the warp's shared TriSetup record, all 32 lanes then load that record (one
// broadcast read) and cover a strided share of the pixel range. (Round 1 ran the set-up redundantly in all 32 lanes and
// kept the record in registers; staging 32 triangles per block was measured too: 10x slower on the m9 meshes, the
// rasterisation needs one warp per triangle in flight.)
__global__ void __launch_bounds__(NT) voxelize_kernel(const float* __restrict__ tris,
                                                      const int32_t* __restrict__ tri_mesh, int T,
                                                      const float* __restrict__ v2x, int W, int H, int D,
                                                      int R, int pdm, int side, int conservative,
                                                      float* __restrict__ grid) {
  __shared__ TriSetup staged[NT / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = blockIdx.x * (NT / 32) + warp; t < T; t += gridDim.x * (NT / 32)) {
    const int mesh = tri_mesh[t];
    if (lane == 0)
      setup_triangle(tris + (int64_t)t * 9, v2x + (int64_t)mesh * 16, W, H, D, R, pdm, conservative != 0, staged[warp]);
    __syncwarp();
    const TriSetup s = staged[warp];  // broadcast read into registers: the pixel loop re-reads the edge terms per pixel
    __syncwarp();
    if (s.valid) {
      const int nx = s.i1 - s.i0 + 1;
      const long long npix = (long long)nx * (s.j1 - s.j0 + 1);
      for (long long q = lane; q < npix; q += 32) {
        const int j = s.j0 + (int)(q / nx), i = s.i0 + (int)(q % nx);
        const long long px = (long long)i * 256 + 128, py = (long long)j * 256 + 128;
        bool in = true;
#pragma unroll
        for (int e3 = 0; e3 < 3; ++e3) {
          const long long e = s.A[e3] * px + s.B[e3] * py + s.C[e3];
          if (conservative) {
            const long long slack = 128 * (llabs(s.A[e3]) + llabs(s.B[e3]));
            in = in && (e + slack >= 0);
          } else {
            in = in && (e > 0 || (e == 0 && s.tl[e3]));
          }
        }
        if (in) shade_fragment(s, i, j, mesh, W, H, D, side, grid);
      }
    }
  }
}

__global__ void __launch_bounds__(NT) merge_grids_kernel(const float* __restrict__ mesh_grids,
                                                         const int32_t* __restrict__ mesh_scene,
                                                         const float* __restrict__ labels, int M,
                                                         int64_t voxels, int32_t* __restrict__ out) {
  const int64_t total = (int64_t)M * voxels;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * NT) {
    const int m = (int)(i / voxels);
    const int64_t v = i - (int64_t)m * voxels;
    const int val = (int)(labels[m] * __ldg(mesh_grids + i));
    if (val > 0) atomicMax(out + (int64_t)mesh_scene[m] * voxels + v, val);
  }
}
}  // namespace

extern "C" int crn_voxelize_mesh(const float* triangles, const int32_t* tri_mesh, int32_t T,
                                 const float* view2voxel, int32_t M, int32_t D, int32_t H, int32_t W,
                                 int32_t image_resolution, int32_t depth_mult, int32_t sub_grid_side,
                                 int32_t conservative, float* grid, void* stream) {
  CRN_REQUIRE(triangles && tri_mesh && view2voxel && grid, "crn_voxelize_mesh: null pointer");
  CRN_REQUIRE(T >= 0 && M > 0 && D > 0 && H > 0 && W > 0 && image_resolution > 0 && depth_mult >= 1,
              "crn_voxelize_mesh: bad sizes");
  CRN_REQUIRE(sub_grid_side <= 0 || sub_grid_side % 2 == 1, "crn_voxelize_mesh: sub-grid side must be odd");
  if (T == 0) return CRN_OK;
  int64_t blocks = crn_ceil_div(T, NT / 32);
  if (blocks > 32LL * kNumSMs) blocks = 32LL * kNumSMs;
  voxelize_kernel<<<(unsigned)blocks, NT, 0, crn_stream(stream)>>>(triangles, tri_mesh, T, view2voxel, W, H, D,
                                                                  image_resolution, depth_mult, sub_grid_side,
                                                                  conservative, grid);
  CRN_LAUNCH_CHECK("voxelize");
  return CRN_OK;
}

extern "C" int crn_merge_mesh_grids(const float* mesh_grids, const int32_t* mesh_scene, const float* labels,
                                    int32_t M, int64_t voxels, int32_t* out, void* stream) {
  CRN_REQUIRE(mesh_grids && mesh_scene && labels && out && M > 0 && voxels > 0, "crn_merge_mesh_grids: bad args");
  int64_t b = crn_ceil_div((int64_t)M * voxels, NT);
  if (b > 16LL * kNumSMs) b = 16LL * kNumSMs;
  merge_grids_kernel<<<(unsigned)b, NT, 0, crn_stream(stream)>>>(mesh_grids, mesh_scene, labels, M, voxels, out);
  CRN_LAUNCH_CHECK("merge_grids");
  return CRN_OK;
}
