from typing import Optional, Tuple, Union

from torch import Tensor

try:  # the shimmed torch_sparse
    from torch_sparse import SparseTensor
except Exception:  # pragma: no cover
    SparseTensor = object

Adj = Union[Tensor, SparseTensor]
OptTensor = Optional[Tensor]
PairTensor = Tuple[Tensor, Tensor]
OptPairTensor = Tuple[Tensor, Optional[Tensor]]
PairOptTensor = Tuple[Optional[Tensor], Optional[Tensor]]
Size = Optional[Tuple[int, int]]
NoneType = Optional[Tensor]
