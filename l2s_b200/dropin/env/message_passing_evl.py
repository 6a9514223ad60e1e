"""Same module path as the reference's env/message_passing_evl.py."""
from l2s_b200.message_passing_evl import CPM_batch_G, Evaluator  # noqa: F401
