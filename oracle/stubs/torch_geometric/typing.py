from typing import Optional
from torch import Tensor
Adj = Tensor
OptTensor = Optional[Tensor]
