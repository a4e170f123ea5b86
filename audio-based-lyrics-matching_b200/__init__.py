"""wealy_b200 -- B200-native (sm_100a) retrieval-and-scoring hot path of WEALY.

Drop-in mirror of the reference's Python interface for this path only:
  * wealy_b200.tensor_ops  <->  /root/reference/lib/tensor_ops.py  (pairwise_distance_matrix, ...)
  * wealy_b200.losses      <->  /root/reference/lib/losses.py      (NTXentLoss, CLEWSLoss)
  * wealy_b200.layers      <->  /root/reference/lib/layers.py MeanPool + the avg-pool collate (pooled-embedding producer)
  * wealy_b200.evaluation  --   the evaluator the reference has not released (AP / MAP / MR1 / top-k)
Everything computes in hand-written CUDA behind the C ABI of include/wealy_b200.h
(lib/libwealy_b200.so); there is no CPU fallback: a missing library or a non-CUDA tensor raises.
"""
from . import _native  # noqa: F401  (fails loudly if the CUDA library is missing)
from . import tensor_ops, evaluation  # noqa: F401
from . import losses, layers  # noqa: F401

__all__ = ["tensor_ops", "losses", "evaluation", "layers"]
