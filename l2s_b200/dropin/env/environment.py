"""Same module path as the reference's env/environment.py."""
from l2s_b200.environment import BatchGraph, JsspN5  # noqa: F401
