"""CPU oracle for voxelize_mesh.  TEST INFRASTRUCTURE ONLY.

The reference rasterises with OpenGL through EGL
(/root/reference/src/corenet/geometry/voxelization.py:115-164,
shaders/voxelize.geom:37-60, shaders/voxelize.frag:29-58), which cannot run
here (no moderngl / libEGL / GL driver).  This is a numpy restatement of the GL
rules those files rely on:

  * per triangle: v_i = (view2voxel[mesh] @ (tri_i, 1)).xyz, fp32           (geom:31-43)
  * dominant axis of |normal| with STRICT comparisons, ties -> no swizzle    (geom:44,53-55)
  * orthographic window coordinates u = coord * R / extent with
    R = round(max(W, H, D*pdm) * mult), extent = (W, H, D*pdm)            (voxelization.py:123-124,146-148)
  * vertices snapped to 1/256 pixel; a fragment per pixel whose centre is inside
    (top-left rule on exact edge hits), or -- conservative -- whose square
    overlaps the triangle; no depth test / culling                        (rasterizer.py:198-210)
  * attributes interpolated (extrapolated) at the pixel centre; bounds test;
    floor -> voxel, or the sub-grid index math                            (frag:32-57)

PARITY STATUS: pinned only by the reference's three known-answer tests
(test/voxelization_test.py:53-147, replayed in tests/test_oracle_voxelize.py).
Beyond those vectors parity with a hardware GL rasteriser is UNPINNED
(sub-pixel snapping / fill-rule details of the GPU vendor are not specified).
Loops are pure Python: small cases only.
"""
import numpy as np

F = np.float32


def _transform(tri, M):
  """((m0*x + m1*y) + m2*z) + m3 with every step rounded to fp32."""
  out = np.zeros((3, 3), F)
  for k in range(3):
    x, y, z = F(tri[k, 0]), F(tri[k, 1]), F(tri[k, 2])
    for r in range(3):
      t = F(F(M[r, 0] * x) + F(M[r, 1] * y))
      t = F(t + F(M[r, 2] * z))
      out[k, r] = F(t + M[r, 3])
  return out


def _axes(v):
  with np.errstate(all="ignore"):
    e1 = (v[1] - v[0]).astype(F)
    e2 = (v[2] - v[0]).astype(F)
    n1, n2 = F(0), F(0)
    for r in range(3):
      n1 = F(n1 + F(e1[r] * e1[r]))
      n2 = F(n2 + F(e2[r] * e2[r]))
    n1, n2 = np.sqrt(n1, dtype=F), np.sqrt(n2, dtype=F)
    e1 = (e1 / n1).astype(F)
    e2 = (e2 / n2).astype(F)
    nx = abs(F(F(e1[1] * e2[2]) - F(e1[2] * e2[1])))
    ny = abs(F(F(e1[2] * e2[0]) - F(e1[0] * e2[2])))
    nz = abs(F(F(e1[0] * e2[1]) - F(e1[1] * e2[0])))
  if nx > ny and nx > nz:
    return 1, 2
  if ny > nx and ny > nz:
    return 2, 0
  return 0, 1


def _fdiv(a, b):
  return a // b          # Python ints: floor division


def _cdiv(a, b):
  return -((-a) // b)


def voxelize_mesh_oracle(triangles, mesh_num_tri, resolution, view2voxel, sub_grid_sampling=False,
                         image_resolution_multiplier=4, conservative_rasterization=False,
                         projection_depth_multiplier=1):
  triangles = np.asarray(triangles, F).reshape(-1, 3, 3)
  mesh_num_tri = np.asarray(mesh_num_tri, np.int64)
  view2voxel = np.asarray(view2voxel, F)
  M = len(mesh_num_tri)
  if view2voxel.ndim == 2:
    view2voxel = np.broadcast_to(view2voxel, (M, 4, 4))
  D, H, W = resolution
  pdm, mult = projection_depth_multiplier, image_resolution_multiplier
  if sub_grid_sampling and mult % 2 == 0:
    raise ValueError("image_resolution_multiplier must be odd with sub_grid_sampling")
  R = int(round(max(W, H, D * pdm) * mult))
  side = int(mult) if sub_grid_sampling else -1
  shape = (M, 2 * D + 1, 2 * H + 1, 2 * W + 1) if sub_grid_sampling else (M, D, H, W)
  grid = np.zeros(shape, F)
  ext = (float(W), float(H), float(D * pdm))
  tri_mesh = np.repeat(np.arange(M), mesh_num_tri)
  for t_i in range(triangles.shape[0]):
    mesh = tri_mesh[t_i]
    v = _transform(triangles[t_i], view2voxel[mesh])
    axA, axB = _axes(v)
    u = [float(v[k, axA]) * float(R) / ext[axA] for k in range(3)]
    w = [float(v[k, axB]) * float(R) / ext[axB] for k in range(3)]
    if not all(abs(x) < 1e9 for x in u + w):
      continue
    U = [int(np.rint(x * 256.0)) for x in u]
    V = [int(np.rint(x * 256.0)) for x in w]
    area2 = (U[1] - U[0]) * (V[2] - V[0]) - (V[1] - V[0]) * (U[2] - U[0])
    if area2 == 0:
      continue
    sg = 1 if area2 > 0 else -1
    A, B, Cc, TL = [], [], [], []
    for k in range(3):
      a, b = k, (k + 1) % 3
      Ak = -sg * (V[b] - V[a])
      Bk = sg * (U[b] - U[a])
      A.append(Ak); B.append(Bk); Cc.append(-(Ak * U[a] + Bk * V[a]))
      TL.append(Ak > 0 or (Ak == 0 and Bk > 0))
    area_d = (u[1] - u[0]) * (w[2] - w[0]) - (w[1] - w[0]) * (u[2] - u[0])
    if area_d == 0.0:
      continue
    mnU, mxU, mnV, mxV = min(U), max(U), min(V), max(V)
    if conservative_rasterization:
      i0, i1 = _cdiv(mnU, 256) - 1, _fdiv(mxU, 256)
      j0, j1 = _cdiv(mnV, 256) - 1, _fdiv(mxV, 256)
    else:
      i0, i1 = _cdiv(mnU - 128, 256), _fdiv(mxU - 128, 256)
      j0, j1 = _cdiv(mnV - 128, 256), _fdiv(mxV - 128, 256)
    i0, j0, i1, j1 = max(i0, 0), max(j0, 0), min(i1, R - 1), min(j1, R - 1)
    for j in range(j0, j1 + 1):
      for i in range(i0, i1 + 1):
        px, py = i * 256 + 128, j * 256 + 128
        inside = True
        for k in range(3):
          e = A[k] * px + B[k] * py + Cc[k]
          if conservative_rasterization:
            inside = inside and (e + 128 * (abs(A[k]) + abs(B[k])) >= 0)
          else:
            inside = inside and (e > 0 or (e == 0 and TL[k]))
        if not inside:
          continue
        su, sv = i + 0.5, j + 0.5
        l1 = ((su - u[0]) * (w[2] - w[0]) - (sv - w[0]) * (u[2] - u[0])) / area_d
        l2 = ((u[1] - u[0]) * (sv - w[0]) - (w[1] - w[0]) * (su - u[0])) / area_d
        l0 = (1.0 - l1) - l2
        p = [F((l0 * float(v[0, r]) + l1 * float(v[1, r])) + l2 * float(v[2, r])) for r in range(3)]
        if (p[0] < 0 or p[1] < 0 or p[2] < 0 or p[0] >= F(W) or p[1] >= F(H) or p[2] >= F(D)):
          continue
        if side <= 0:
          cx, cy, cz = int(np.floor(p[0])), int(np.floor(p[1])), int(np.floor(p[2]))
          grid[mesh, cz, cy, cx] = 1
        else:
          c = []
          for r in range(3):
            vv = int(np.floor(F(p[r] * F(side)))) + side // 2
            c.append(2 * (vv // side) + (1 if vv % side == side - 1 else 0))
          grid[mesh, c[2], c[1], c[0]] = 1
  return grid


def get_sub_grid_centers(grid):
  """voxelization.py:167-182."""
  return grid[:, 1::2, 1::2, 1::2]
