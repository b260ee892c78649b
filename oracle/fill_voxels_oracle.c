/* CPU oracle for fill_inside_voxels.  TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Plain-C restatement of the reference's sequential algorithm
 * (/root/reference/src/corenet/cc/fill_voxels_cpu.cc):
 *   - raster-scan connected-component labelling of equal-occupancy regions with
 *     a union-find; a voxel on the x=0 / y=0 / z=0 border sees an EMPTY
 *     neighbour belonging to region 0 = "outside" (:92-102); far faces have no
 *     such neighbour (near-face rule);
 *   - the smaller region id always becomes the root (:39-47), so region 0 stays
 *     the root of everything connected to the outside;
 *   - voxels whose root is not 0 are set to 1, the rest are left untouched
 *     (:150-154).
 * Pinned by tests/test_oracle_fill.py against the reference's own known-answer
 * grids (test/voxelization_test.py:150-248) and against oracle/_ref (the real
 * reference source compiled in the build container).
 *
 *   gcc -O2 -shared -fPIC -o oracle/_build/libfill_oracle.so oracle/fill_voxels_oracle.c
 */
#include <stdint.h>
#include <stdlib.h>

static int64_t find_root(int64_t* parent, int64_t e) {
  int64_t r = e;
  while (parent[r] >= 0) r = parent[r];
  while (parent[e] >= 0) { int64_t nx = parent[e]; parent[e] = r; e = nx; }   /* path compression */
  return r;
}

static void merge(int64_t* parent, int64_t a, int64_t b) {
  a = find_root(parent, a);
  b = find_root(parent, b);
  if (a < b) parent[b] = a; else if (b < a) parent[a] = b;
}

/* occ: D*H*W bytes (1 = occupied i.e. value > 0).  out: D*H*W bytes, 1 where the
 * reference would write a 1 (region root != 0), else 0. Returns 0, or -1 on OOM. */
int fill_oracle_one(const uint8_t* occ, uint8_t* out, int D, int H, int W) {
  const int64_t n = (int64_t)D * H * W;
  int64_t* region = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  int64_t* parent = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
  if (!region || !parent) { free(region); free(parent); return -1; }
  int64_t nregions = 0;
  parent[nregions++] = -1;                                    /* region 0 = outside */
  for (int z = 0; z < D; z++)
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) {
        const int64_t i = ((int64_t)z * H + y) * W + x;
        const int cur = occ[i];
        /* neighbours towards the near faces; outside the grid: empty, region 0 */
        const int vl = x > 0 ? occ[i - 1] : 0, vb = y > 0 ? occ[i - W] : 0, vu = z > 0 ? occ[i - (int64_t)H * W] : 0;
        const int64_t rl = x > 0 ? region[i - 1] : 0, rb = y > 0 ? region[i - W] : 0,
                      ru = z > 0 ? region[i - (int64_t)H * W] : 0;
        if (cur == vl && cur == vu) merge(parent, rl, ru);
        if (cur == vl && cur == vb) merge(parent, rl, rb);
        if (cur == vu && cur == vb) merge(parent, ru, rb);
        int64_t best = INT64_MAX;
        if (cur == vl && rl < best) best = rl;
        if (cur == vb && rb < best) best = rb;
        if (cur == vu && ru < best) best = ru;
        if (best == INT64_MAX) { best = nregions; parent[nregions++] = -1; }
        region[i] = best;
      }
  for (int64_t i = 0; i < n; i++) out[i] = find_root(parent, region[i]) > 0 ? 1 : 0;
  free(region);
  free(parent);
  return 0;
}

int fill_oracle(const uint8_t* occ, uint8_t* out, int N, int D, int H, int W) {
  const int64_t n = (int64_t)D * H * W;
  for (int b = 0; b < N; b++)
    if (fill_oracle_one(occ + b * n, out + b * n, D, H, W)) return -1;
  return 0;
}
