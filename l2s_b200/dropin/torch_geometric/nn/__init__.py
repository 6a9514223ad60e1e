from .conv import GATConv, GINConv, MessagePassing  # noqa: F401
from .glob import global_add_pool, global_max_pool, global_mean_pool  # noqa: F401
