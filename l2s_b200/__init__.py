"""l2s_b200 -- B200-native (sm_100a) drop-in for the hot path of zcaicaros/L2S: the batched
disjunctive-graph evaluator and the N5 local-search environment step.

    from l2s_b200 import JsspN5, BatchGraph, Evaluator

or put `l2s_b200/dropin` first on `sys.path` so that the reference's own import lines
(`from env.environment import JsspN5, BatchGraph`, `from env.message_passing_evl import Evaluator`)
resolve to this package.
"""
from ._capi import L2SError, lib, lib_path  # noqa: F401
from .environment import BatchGraph, JsspN5  # noqa: F401
from .message_passing_evl import CPM_batch_G, Evaluator  # noqa: F401

__all__ = ["JsspN5", "BatchGraph", "Evaluator", "CPM_batch_G", "L2SError", "lib", "lib_path"]
