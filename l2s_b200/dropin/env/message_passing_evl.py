"""Same module path as the reference's env/message_passing_evl.py."""
from l2s_b200.message_passing_evl import Evaluator  # noqa: F401
