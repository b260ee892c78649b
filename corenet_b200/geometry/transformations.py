"""4x4 transformation helpers used around the hot path (tiny; plain torch, so they run on whatever device their
inputs live on).

Same names / argument meaning as the reference's
src/corenet/geometry/transformations.py (scale :25, translate :40, rotate :61, transform_points_homogeneous :108,
transform_mesh :139, transform_points :172, look_at_lh :179, look_at_rh :201, perspective_lh :223, perspective_rh :244,
ortho_lh :265, chain :289); pinned by the reference's own known answers (test/transformations_test.py) and against
the reference functions in tests/test_host.py.
"""
import torch as t
from torch.nn import functional as F


def _f32(v, device=None):
  return t.as_tensor(v, dtype=t.float32, device=device)


def scale(v) -> t.Tensor:
  v = _f32(v)
  assert v.dim() == 1
  return t.diag(t.cat([v, v.new_ones([1])]))


def translate(v) -> t.Tensor:
  v = _f32(v)
  n = v.shape[-1]
  m = t.eye(n + 1, dtype=t.float32, device=v.device).expand(*v.shape[:-1], n + 1, n + 1).clone()
  m[..., :n, n] = v
  return m


def look_at_rh(eye, center, up) -> t.Tensor:
  eye, center, up = _f32(eye), _f32(center), _f32(up)
  f = F.normalize(center - eye, dim=-1)
  s = F.normalize(t.linalg.cross(f, up), dim=-1)
  u = t.linalg.cross(s, f)
  m = t.eye(4, dtype=t.float32)
  m[0, :3], m[1, :3], m[2, :3] = s, u, -f
  m[0, 3], m[1, 3], m[2, 3] = -t.dot(s, eye), -t.dot(u, eye), t.dot(f, eye)
  return m


def perspective_rh(fov_y, aspect, z_near, z_far) -> t.Tensor:
  fov_y, aspect, z_near, z_far = _f32(fov_y), _f32(aspect), _f32(z_near), _f32(z_far)
  th = t.tan(fov_y / 2)
  m = t.zeros(4, 4, dtype=t.float32)
  m[0, 0] = 1.0 / (aspect * th)
  m[1, 1] = 1.0 / th
  m[2, 2] = -(z_far + z_near) / (z_far - z_near)
  m[2, 3] = -(2 * z_far * z_near) / (z_far - z_near)
  m[3, 2] = -1
  return m


def ortho_lh(left, right, bottom, top, z_near, z_far) -> t.Tensor:
  l, r, b, tp, n, f = [float(x) for x in (left, right, bottom, top, z_near, z_far)]
  m = t.eye(4, dtype=t.float32)
  m[0, 0], m[0, 3] = 2 / (r - l), -(r + l) / (r - l)
  m[1, 1], m[1, 3] = 2 / (tp - b), -(tp + b) / (tp - b)
  m[2, 2], m[2, 3] = 2 / (f - n), -(f + n) / (f - n)
  return m


def transform_points_homogeneous(points, matrix, w: float) -> t.Tensor:
  points, matrix = _f32(points), _f32(matrix)
  assert points.shape[-1] == 3 and matrix.shape[-2:] == (4, 4)
  points = t.constant_pad_nd(points, [0, 1], value=w)
  return t.einsum("...nm,...vm->...vn", matrix, points)


def transform_points(points, matrix) -> t.Tensor:
  """Affine points through 4x4 matrices, perspective divide included (transformations.py:172-176)."""
  h = transform_points_homogeneous(points, matrix, w=1)
  return h[..., :3] / h[..., 3:4]


def transform_mesh(mesh, matrix, vertices_are_points: bool = True) -> t.Tensor:
  """float32[..., num_tri, 3, 3] triangle vertices through float32[..., 4, 4] (transformations.py:139-169): the
  triangles are flattened to 3 * num_tri points (w = 1, perspective divide) or vectors (w = 0)."""
  mesh = _f32(mesh)
  matrix = _f32(matrix, mesh.device)
  assert mesh.shape[-2:] == (3, 3) and matrix.shape[-2:] == (4, 4) and mesh.shape[:-3] == matrix.shape[:-2]
  pts = mesh.reshape(mesh.shape[:-3] + (mesh.shape[-3] * 3, 3))
  if vertices_are_points:
    out = transform_points(pts, matrix)
  else:
    out = transform_points_homogeneous(pts, matrix, w=0)[..., :3]
  return out.reshape(mesh.shape)


def rotate(angle, axis) -> t.Tensor:
  """Rotation by `angle` radians about `axis` (transformations.py:61-105), Rodrigues' form
  R = cos * I + sin * [a]_x + (1 - cos) * a a^T."""
  angle, axis = _f32(angle), _f32(axis)
  assert axis.shape == (3,) and angle.shape == ()
  a = F.normalize(axis, dim=-1)
  x, y, z = a.unbind(-1)
  o = t.zeros_like(x)
  cross = t.stack([o, -z, y, z, o, -x, -y, x, o]).reshape(3, 3)
  r = t.cos(angle) * t.eye(3) + t.sin(angle) * cross + (1 - t.cos(angle)) * t.outer(a, a)
  m = t.eye(4, dtype=t.float32)
  m[:3, :3] = r
  return m


def look_at_lh(eye, center, up) -> t.Tensor:
  """Left-handed view matrix (transformations.py:179-198): +z looks from eye to center."""
  eye, center, up = _f32(eye), _f32(center), _f32(up)
  f = F.normalize(center - eye, dim=-1)
  s = F.normalize(t.linalg.cross(up, f), dim=-1)
  u = t.linalg.cross(f, s)
  m = t.eye(4, dtype=t.float32)
  m[0, :3], m[1, :3], m[2, :3] = s, u, f
  m[0, 3], m[1, 3], m[2, 3] = -t.dot(s, eye), -t.dot(u, eye), -t.dot(f, eye)
  return m


def perspective_lh(fov_y, aspect, z_near, z_far) -> t.Tensor:
  """Left-handed perspective projection (transformations.py:223-241): perspective_rh with the z column negated."""
  m = perspective_rh(fov_y, aspect, z_near, z_far)
  m[2, 2], m[3, 2] = -m[2, 2], 1
  return m


def chain(transforms) -> t.Tensor:
  """Product of a non-empty list of matrices, left to right (transformations.py:289-295)."""
  assert transforms
  result = transforms[0]
  for m in transforms[1:]:
    result = t.mm(result, m)
  return result
