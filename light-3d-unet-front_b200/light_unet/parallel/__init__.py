"""Multi-GPU partitioning of the hot path (new capability; the reference is single-device).

  * training: data parallel, one process per GPU.  The Focal Tversky loss is a ratio of *batch-global* sums
    (losses.py:44-49), so the three sums {sum p*t, sum p, sum t} are all-reduced before the ratio and the parameter
    gradients are reduced with SUM -- a dp-N step on N x B patches is then the reference's single-process step on
    the concatenated batch (not the mean of per-rank losses).
  * inference: volumes (cases) or the windows of one volume are dealt to ranks; no data-path collective.
"""
from .data_parallel import DataParallelStep, flatten_grads, shard_cases, shard_windows

__all__ = ["DataParallelStep", "flatten_grads", "shard_cases", "shard_windows"]
