"""Type aliases the reference imports (env/message_passing_evl.py:9)."""
from typing import Optional, Tuple, Union

from torch import Tensor

Adj = Tensor
OptTensor = Optional[Tensor]
PairTensor = Tuple[Tensor, Tensor]
OptPairTensor = Tuple[Tensor, Optional[Tensor]]
Size = Optional[Tuple[int, int]]
NoneType = Optional[Tensor]
__all__ = ["Adj", "OptTensor", "PairTensor", "OptPairTensor", "Size", "NoneType", "Union"]
