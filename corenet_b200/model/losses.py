"""Training losses with the reference's names and argument order.

`iou_fgbg` (FG_BG task) and `xent_times_iou_agnostic` (SEMANTIC task) are what
`TrainPipeline` selects (src/corenet/pipeline.py:154-158 of the reference); they
follow src/corenet/model/losses.py:64-114 and :144-160, fused into one pass
over the logits forward and one backward (csrc/loss.cu).
"""
from corenet_b200.ops import iou_fgbg, xent_times_iou_agnostic  # noqa: F401
