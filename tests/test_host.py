"""Host-side logic that needs no GPU: module surface, state_dict compatibility, config, sharding, errors."""
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


def test_shard_indices():
  from corenet_b200.trainer import shard_indices
  assert shard_indices(10, 0, 4) == [0, 4, 8]
  assert shard_indices(10, 3, 4) == [3, 7, 1]
  parts = [shard_indices(10, r, 4, pad=False) for r in range(4)]
  assert sorted(sum(parts, [])) == list(range(10))


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
