"""BatchRenorm with the reference's parameters/buffers, computed by the CUDA kernels.

Boundary row a6: same constructor, state_dict keys (`weight`, `bias`,
`running_mean`, `running_var`, `num_batches_tracked`) and train/eval semantics
as src/corenet/model/batch_renorm.py:18-62 of the reference.  Inside CoreNet
the engine drives the kernels directly; this standalone forward exists so the
module can be used (and parity-tested) on its own.
"""
import torch as t

from corenet_b200 import ops


class BatchRenorm(t.nn.Module):
  def __init__(self, num_channels: int, eps: float = 1e-5, momentum: float = 0.01):
    super().__init__()
    self.eps = eps
    self.momentum = momentum
    self.weight = t.nn.Parameter(t.ones(num_channels, dtype=t.float32))
    self.bias = t.nn.Parameter(t.zeros(num_channels, dtype=t.float32))
    self.register_buffer("running_mean", t.zeros(num_channels, dtype=t.float32))
    self.register_buffer("running_var", t.ones(num_channels, dtype=t.float32))
    self.register_buffer("num_batches_tracked", t.tensor(0, dtype=t.int64))

  def forward(self, x: t.Tensor) -> t.Tensor:
    assert x.dim() >= 2
    return ops.batch_renorm(x, self.weight, self.bias, self.running_mean, self.running_var,
                            self.num_batches_tracked, self.training, self.eps, self.momentum)
