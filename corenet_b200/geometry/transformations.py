"""4x4 transformation helpers used around the hot path (host side, tiny).

Same names / argument meaning as the reference's
src/corenet/geometry/transformations.py (scale :25, translate :41,
look_at_rh :206, perspective_rh :250, ortho_lh :273, transform_points_homogeneous :108).
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
