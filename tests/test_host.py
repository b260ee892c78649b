"""Host-side logic that needs no GPU: module surface, state_dict compatibility, config, sharding, errors."""
import dataclasses

import numpy as np
import pytest
import torch as t

from corenet_b200 import configuration as C
from corenet_b200.model.core_net import CoreNet


def test_state_dict_matches_reference_fixture(golden):
  t.manual_seed(0)
  m = CoreNet(C.default_config(2))
  sd = m.state_dict()
  assert len(sd) == 458 and len(list(m.parameters())) == 266
  assert sum(p.numel() for p in m.parameters()) == 36141888
  assert list(sd.keys()) == [str(k) for k in golden["param_names"]]
  got = np.array([v.double().abs().sum().item() for v in sd.values()])
  np.testing.assert_allclose(got, golden["param_abssum"], rtol=1e-12)
  assert sd["encoder.stage2.a.op_a.bn.num_batches_tracked"].dtype == t.int64
  assert tuple(sd["decoder.stage_1.t1.weight"].shape) == (67, 256, 4, 4, 4)
  assert tuple(sd["decoder.rt_skip_2.compress_channels.weight"].shape) == (96, 2051, 1, 1)


def test_config_roundtrip_and_geometry_guard():
  cfg = C.default_config(15)
  assert C.CoreNetConfig.from_dict(cfg.to_dict()) == cfg
  with pytest.raises(ValueError):     # only 128^3 / last_upscale_factor=2 exists
    CoreNet(C.CoreNetConfig(decoder=C.DecoderConfig(resolution=(32, 32, 32), num_output_channels=2)))


def test_no_cpu_fallback():
  m = CoreNet(C.default_config(2))
  img = t.zeros(1, 3, 256, 256, dtype=t.uint8)
  with pytest.raises(RuntimeError):
    m(img, t.eye(4)[None], t.full((1, 3), 0.5))
  from corenet_b200.cc import fill_voxels
  with pytest.raises(ValueError):
    fill_voxels.fill_inside_voxels_gpu(t.zeros(1, 4, 4, 4))
  from corenet_b200 import ops
  with pytest.raises(ValueError):
    ops.iou_fgbg(t.zeros(1, 4, 4, 4, dtype=t.int64), t.zeros(1, 2, 4, 4, 4))


def test_engine_layer_table():
  from corenet_b200 import engine
  m = CoreNet(C.default_config(2))
  eng = engine.get_engine(m)
  assert len(eng.layers) == 53 + 1 + 6 + 5 + 4          # encoder convs, linear, convT, conv3d, skip 1x1
  names = {l.name + ".weight" for l in eng.layers}
  conv_like = {n for n, p in m.named_parameters() if p.dim() >= 2}
  assert names == conv_like
  macs = 0
  for l in eng.layers:
    assert l.cinp % 4 == 0 and l.coutp % 4 == 0
  assert eng.dec_plan[-1][0] == 6


def test_shard_indices_follow_reference_sampler(monkeypatch):
  """Vector computed with the reference's DistributedSampler (distributed.py:204-224: randperm seeded 0x1234,
  zero-padded, contiguous blocks); re-checked against the class itself when a reference checkout is present."""
  from corenet_b200.trainer import shard_indices
  assert shard_indices(10, 0, 4) == [6, 1, 0]
  assert shard_indices(10, 3, 4) == [3, 0, 0]                      # padded with index 0
  assert shard_indices(10, 1, 4, pad=False) == [0, 7, 5]
  parts = [shard_indices(10, r, 4, pad=False) for r in range(4)]
  assert sorted(sum(parts, [])) == list(range(10))
  from baseline import ref_import
  if ref_import.import_reference() is not None:
    from corenet import distributed as ref_dist
    # torch >= 2.x: Sampler.__init__ no longer takes the data source the reference passes (distributed.py:209)
    monkeypatch.setattr(t.utils.data.Sampler, "__init__", lambda self, *a, **k: None, raising=False)
    for n, world, pad in ((10, 4, True), (10, 4, False), (37, 8, True), (5, 2, False), (64, 8, True)):
      for r in range(world):
        smp = ref_dist.DistributedSampler(list(range(n)), r, world, pad)
        assert [int(i) for i in smp] == shard_indices(n, r, world, pad), (n, world, pad, r)


def test_reference_state_roundtrip_with_dropin_module():
  """The reference's own state.py (encode_state / decode_state, :74-97) over the overlaid module (CPU: no forward)."""
  from baseline import ref_import
  if ref_import.import_reference() is None:
    pytest.skip("no reference checkout")
  from corenet_b200 import compat
  compat.install()
  from corenet import configuration as ref_cfg
  from corenet import state as ref_state
  from corenet.model import core_net
  assert core_net.CoreNet is CoreNet
  cfg = ref_cfg.CoreNetConfig(decoder=ref_cfg.DecoderConfig(
      resolution=(128, 128, 128), num_output_channels=15, last_upscale_factor=2, latent_channels=64,
      skip_fraction=0.75))
  t.manual_seed(1)
  m = core_net.CoreNet(cfg)
  opt = t.optim.Adam(m.parameters(), lr=4e-4, eps=1e-4)
  st = ref_state.State(global_step=7, model=m, optimizer=opt, extra_metadata={"note": 1})
  st2 = ref_state.decode_state(ref_state.encode_state(st), "cpu")
  assert isinstance(st2.model, CoreNet) and st2.global_step == 7 and st2.extra_metadata == {"note": 1}
  assert st2.model.config.decoder.num_output_channels == 15
  for (k1, v1), (k2, v2) in zip(m.state_dict().items(), st2.model.state_dict().items()):
    assert k1 == k2 and t.equal(v1, v2)
  # the encoder checkpoint interface of create_initial_state (state.py:69): 371 keys
  enc = m.encoder.state_dict()
  assert len(enc) == 371
  st2.model.encoder.load_state_dict(enc)


def test_json5_config_surface():
  text = """// generated
  {
    string_templates: [{key: "data_dir", value: "data"}, {key: 'out', value: "{data_dir}/o",},],
    train: {
      data: {
        datasets: [{dataset_path: "{data_dir}/a.json", /* c */ high_realism: true,},],
        data_loader: {num_data_workers: 6, batch_size: 4, prefetch_factor: 2,},
        voxelization_config: {
          task_type: "SEMANTIC",
          resolution: {depth: 128, height: 128, width: 128,},
          sub_grid_sampling: false, conservative_rasterization: false,
          voxelization_image_resolution_multiplier: 8, voxelization_projection_depth_multiplier: 1,
        },
      },
      initial_learning_rate: 0.0004, adam_epsilon: 0.0001, note: "a, } b // not a comment",
    },
    output_path: "{out}/m7",
  }"""
  cfg = C.expand_templates(C.parse_json5(text))
  assert cfg["output_path"] == "data/o/m7" and cfg["train"]["note"] == "a, } b // not a comment"
  assert cfg["train"]["data"]["datasets"][0]["dataset_path"] == "data/a.json"
  hp = C.hot_path_settings(cfg)
  assert hp["batch_size"] == 4 and hp["loss"] == "xent_times_iou_agnostic" and hp["resolution"] == (128, 128, 128)
  v = hp["voxelization_config"]
  assert v.task_type == C.TaskType.SEMANTIC and v.voxelization_image_resolution_multiplier == 8
  assert not v.conservative_rasterization
  import glob
  for p in glob.glob("/root/reference/configs/*/*.json5"):        # build container only
    hp = C.hot_path_settings(C.load_config(p))
    assert hp["resolution"] == (128, 128, 128) and hp["batch_size"] in (4, 8)


def test_super_resolution_host_logic_32_to_128():
  """y1-style plumbing without a GPU: SuperResolutionInference over a stub 32^3 model (super_resolution.py:45-112)."""
  from corenet_b200.super_resolution import SuperResolutionInference
  from oracle import corenet_oracle as O
  seen = []

  def stub(image, cam, v2x, grid_offsets):
    seen.append((v2x, grid_offsets))
    n_off, b = grid_offsets.shape[:2]
    return grid_offsets.sum(-1)[:, :, None, None, None, None] + t.zeros(n_off, b, 3, 32, 32, 32)
  sr = SuperResolutionInference(stub, (32, 32, 32))
  go = t.tensor([[0.5, 0.25, 0.75], [0.1, 0.2, 0.3]])
  v2x = t.eye(4)[None].expand(2, 4, 4)
  out = sr(t.zeros(2, 3, 8, 8, dtype=t.uint8), v2x, v2x, go, (128, 128, 128))
  native = O.native_offsets(4, go)
  assert t.allclose(seen[0][1], native) and t.allclose(seen[0][0], v2x @ O.scale([0.25] * 3))
  want = O.interleave_pmfs(native.sum(-1)[:, :, None, None, None, None] + t.zeros(64, 2, 3, 32, 32, 32), 4)
  assert t.equal(out, want) and tuple(out.shape) == (2, 3, 128, 128, 128)
  with pytest.raises(ValueError):
    sr.get_resolution_multiplier((100, 128, 128))


def test_transformations_match_oracle():
  from corenet_b200.geometry import transformations as tt
  from oracle import corenet_oracle as O
  import math
  assert t.equal(tt.scale([1, 2, 3]), O.scale([1, 2, 3]))
  assert t.equal(tt.translate([[1, 2, 3], [4, 5, 6]]), O.translate([[1, 2, 3], [4, 5, 6]]))
  assert t.equal(tt.perspective_rh(math.pi / 3, 1, 1e-4, 10), O.perspective_rh(math.pi / 3, 1, 1e-4, 10))
  a = tt.look_at_rh([.5, .5, -.87], [.5, .5, .5], [0, -1, 0])
  assert t.allclose(a, O.look_at_rh([.5, .5, -.87], [.5, .5, .5], [0, -1, 0]), atol=0, rtol=0)
  assert t.equal(tt.ortho_lh(0, 4, 3, 0, 0, 5), O.ortho_lh(0, 4, 3, 0, 0, 5))


def test_compat_overlay_without_a_reference_checkout():
  """compat.install() in a process that has no reference on its path: `corenet.*` resolves to this package (the
  overlaid hot-path modules plus the fallbacks for configuration / super_resolution / batched_example)."""
  import subprocess
  import sys
  from tests.conftest import ROOT
  code = (
      "import sys; sys.path.insert(0, %r)\n"
      "import corenet_b200.compat as compat; compat.install()\n"
      "import corenet.model.core_net as cn, corenet.configuration as cfg, corenet.cc.fill_voxels as fv\n"
      "import corenet.super_resolution as sr, corenet.geometry.voxelization as vz\n"
      "from corenet.model import losses\n"
      "from corenet.data import batched_example\n"
      "m = cn.CoreNet(cfg.default_config(2))\n"
      "assert type(m).__module__ == 'corenet_b200.model.core_net' and len(m.state_dict()) == 458\n"
      "assert sr.super_resolution_from_state.__module__ == 'corenet_b200.super_resolution'\n"
      "assert vz.voxelize_mesh.__module__ == 'corenet_b200.geometry.voxelization'\n"
      "assert losses.iou_fgbg and batched_example.voxelize and fv.get_module\n"
      "print('ok')\n") % ROOT
  r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/")
  assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-800:]


def test_precision_mode_switch_is_host_side_state():
  """engine.set_precision validates its argument and flips bit 13 of crn_set_flags (host-only call: no GPU needed);
  the default is the fp32-class 3xTF32 split."""
  from corenet_b200 import engine
  assert engine.PRECISION == "3xtf32"
  with pytest.raises(ValueError):
    engine.set_precision("bf16")
  try:
    engine.set_precision("tf32")
    assert engine.PRECISION == "tf32"
  finally:
    engine.set_precision("3xtf32")
  assert engine.PRECISION == "3xtf32"


def test_transformations_reference_known_answers():
  """src/corenet/test/transformations_test.py:26-115 on this package's helpers."""
  import math
  from corenet_b200.geometry import transformations as tt
  assert t.equal(tt.scale((1, 2, 3)), t.tensor(((1, 0, 0, 0), (0, 2, 0, 0), (0, 0, 3, 0), (0, 0, 0, 1)), dtype=t.float32))
  assert t.equal(tt.translate((1, 2, 3)),
                 t.tensor(((1, 0, 0, 1), (0, 1, 0, 2), (0, 0, 1, 3), (0, 0, 0, 1)), dtype=t.float32))
  two = tt.translate([[[1, 2, 3], [4, 5, 6]]])
  assert two.shape == (1, 2, 4, 4) and two[0, 1, :3, 3].tolist() == [4, 5, 6] and two[0, 0, :3, 3].tolist() == [1, 2, 3]
  assert t.allclose(tt.rotate(math.pi / 2, (0, 0, 1)),
                    t.tensor(((0, -1, 0, 0), (1, 0, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1)), dtype=t.float32),
                    rtol=1e-5, atol=1e-5)
  m1 = ((1, 0, 0, 0), (0, 2, 0, 0), (0, 0, 3, 0), (0, 0, 0, 1))
  m2 = ((1, 0, 0, 1), (0, 1, 0, 2), (0, 0, 1, 3), (0, 0, 0, 1))
  p1 = ((12, 34, 56), (34, 32, 30), (11, 11, 18), (5, 6, 7))
  p2 = ((1, 2, 3), (4, 5, 6), (6, 5, 4), (3, 2, 1))
  want = t.tensor((((12, 68, 168), (34, 64, 90), (11, 22, 54), (5, 12, 21)),
                   ((2, 4, 6), (5, 7, 9), (7, 7, 7), (4, 4, 4))), dtype=t.float32)
  h = tt.transform_points_homogeneous((p1, p2), (m1, m2), w=1)
  assert t.equal(h[..., :3] / h[..., 3:4], want) and t.equal(tt.transform_points((p1, p2), (m1, m2)), want)
  mesh = (((12, 34, 56), (34, 32, 30), (11, 11, 18)), ((1, 2, 3), (4, 5, 6), (6, 5, 4)))
  want = t.tensor((((12, 68, 168), (34, 64, 90), (11, 22, 54)), ((1, 4, 9), (4, 10, 18), (6, 10, 12))), dtype=t.float32)
  assert t.equal(tt.transform_mesh(mesh, m1), want)


def _reference_or_skip():
  from baseline import ref_import
  if ref_import.import_reference() is None:
    pytest.skip("no reference checkout / staged copy")


def test_transformations_and_batching_match_the_reference_functions():
  """rotate / look_at / perspective / transform_* / chain against the reference module on random inputs, and
  data.batched_example.batch (one batched product over all triangles) against the reference's per-mesh loop
  (batched_example.py:67-97)."""
  import types
  import warnings
  _reference_or_skip()
  from corenet.data import batched_example as rb
  from corenet.geometry import transformations as rt
  from corenet_b200.data import batched_example as mb
  from corenet_b200.geometry import transformations as tt
  g = t.Generator().manual_seed(0)
  with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for _ in range(4):
      ang, ax = float(t.rand(1, generator=g) * 6 - 3), t.randn(3, generator=g)
      assert t.allclose(tt.rotate(ang, ax), rt.rotate(ang, ax), atol=1e-6)
      e, c, u = t.randn(3, generator=g), t.randn(3, generator=g), t.randn(3, generator=g)
      assert t.allclose(tt.look_at_lh(e, c, u), rt.look_at_lh(e, c, u), atol=1e-6)
      assert t.allclose(tt.look_at_rh(e, c, u), rt.look_at_rh(e, c, u), atol=1e-6)
      assert t.equal(tt.perspective_lh(1.0, 1.3, 0.01, 9.0), rt.perspective_lh(1.0, 1.3, 0.01, 9.0))
      mesh, mat = t.randn(2, 7, 3, 3, generator=g), t.randn(2, 4, 4, generator=g)
      assert t.allclose(tt.transform_mesh(mesh, mat), rt.transform_mesh(mesh, mat), rtol=1e-5, atol=1e-6)
      assert t.allclose(tt.transform_mesh(mesh, mat, False), rt.transform_mesh(mesh, mat, False), rtol=1e-5, atol=1e-6)
      ms = [t.randn(4, 4, generator=g) for _ in range(3)]
      assert t.equal(tt.chain(ms), rt.chain(ms))

    def elem(nm, seed):
      gg = t.Generator().manual_seed(seed)
      ntri = t.randint(3, 9, (nm,), generator=gg).to(t.int32)
      return types.SimpleNamespace(
          view_transform=t.randn(4, 4, generator=gg), camera_transform=t.randn(4, 4, generator=gg), mesh_num_tri=ntri,
          o2w_transforms=t.randn(nm, 4, 4, generator=gg), mesh_vertices=t.randn(int(ntri.sum()), 3, 3, generator=gg),
          mesh_labels=t.arange(nm, dtype=t.int32), input_image=t.zeros(3, 8, 8, dtype=t.uint8), scene_id=f"s{seed}")
    exs = [elem(2, 1), elem(3, 2), elem(1, 3)]
    a, b = rb.batch(exs), mb.batch(exs)
  assert t.allclose(a.vertices, b.vertices, rtol=1e-5, atol=1e-5)
  for f in ("view_transform", "camera_transform", "input_image", "grid_sampling_offset"):
    assert t.equal(getattr(a, f), getattr(b, f)), f
  assert a.scene_id == b.scene_id and all(t.equal(x, y) for x, y in zip(a.mesh_num_tri, b.mesh_num_tri))
  assert b.to("cpu").vertices.shape == a.vertices.shape and b.grid is None


def test_voxelize_batch_passes_the_config_through(monkeypatch):
  """data.batched_example.voxelize_batch = pipeline.voxelize_batch (pipeline.py:126-150): task type -> voxel content,
  VoxelizationConfig fields -> rasteriser settings (the CUDA step itself is stubbed: host logic only)."""
  from corenet_b200 import configuration as C
  from corenet_b200.data import batched_example as be
  seen = {}

  def fake(vertices, mesh_num_tri, offsets, resolution, **kw):
    seen.update(kw, resolution=resolution, n=len(mesh_num_tri))
    return "v2x", "grid"
  monkeypatch.setattr(be, "voxelize", fake)
  ex = be.BatchedExample(vertices=t.zeros(5, 3, 3), view_transform=t.eye(4)[None], camera_transform=t.eye(4)[None],
                         mesh_num_tri=[t.tensor([2, 3], dtype=t.int32)], mesh_labels=[t.tensor([7, 9], dtype=t.int32)],
                         input_image=t.zeros(1, 3, 8, 8, dtype=t.uint8), scene_id=["s"],
                         grid_sampling_offset=t.full((1, 3), 0.5))
  cfg = C.VoxelizationConfig(task_type=C.TaskType.SEMANTIC, resolution=C.Resolution(depth=128, height=64, width=32),
                             sub_grid_sampling=False, conservative_rasterization=True,
                             voxelization_image_resolution_multiplier=7, voxelization_projection_depth_multiplier=2)
  out = be.voxelize_batch(ex, cfg)
  assert out.grid == "grid" and out.v2x_transform == "v2x" and out.vertices is ex.vertices
  assert seen["resolution"] == (128, 64, 32) and seen["image_resolution_multiplier"] == 7
  assert seen["conservative_rasterization"] is True and seen["projection_depth_multiplier"] == 2
  assert seen["sub_grid_sampling"] is False and seen["voxel_content_fn"](0, 1) == 9
  cfg = dataclasses.replace(cfg, task_type=C.TaskType.FG_BG)
  be.voxelize_batch(ex, cfg)
  assert seen["voxel_content_fn"](0, 1) == 1


def test_super_resolution_interleave_equals_the_reference_class():
  """SuperResolutionInference (offsets, v2x rescale, pmf interleave) against the reference's own class
  (super_resolution.py:45-112) around the same stub model with spatially varying output, multipliers 2 and 4."""
  import warnings
  _reference_or_skip()
  with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    from corenet import super_resolution as rs
    from corenet_b200.super_resolution import SuperResolutionInference
    calls = []

    def stub(image, cam, v2x, grid_offsets):
      calls.append((cam.clone(), v2x.clone(), grid_offsets.clone()))
      n_off, b = grid_offsets.shape[:2]
      g = t.Generator().manual_seed(n_off)
      return t.rand(n_off, b, 3, 8, 8, 8, generator=g) + grid_offsets.sum(-1)[:, :, None, None, None, None]
    go = t.tensor([[0.5, 0.25, 0.75], [0.1, 0.2, 0.3]])
    g = t.Generator().manual_seed(1)
    cam, v2x = t.randn(2, 4, 4, generator=g), t.randn(2, 4, 4, generator=g)
    img = t.zeros(2, 3, 8, 8, dtype=t.uint8)
    for mult in (2, 4):
      out_res = (8 * mult,) * 3
      a = rs.SuperResolutionInference(stub, (8, 8, 8))(img, cam, v2x, go, out_res)
      b = SuperResolutionInference(stub, (8, 8, 8))(img, cam, v2x, go, out_res)
      assert t.equal(a, b) and tuple(b.shape) == (2, 3) + out_res
      (c1, x1, o1), (c2, x2, o2) = calls[-2], calls[-1]
      assert t.equal(c1, c2) and t.equal(x1, x2) and t.equal(o1, o2)
