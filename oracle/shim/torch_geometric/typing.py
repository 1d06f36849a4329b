from typing import Optional, Tuple, Union
from torch import Tensor

class SparseTensor:  # placeholder: the reference never builds one
    pass

class _TorchSparse:
    @staticmethod
    def set_diag(x):
        raise NotImplementedError

torch_sparse = _TorchSparse()
Adj = Union[Tensor, SparseTensor]
OptTensor = Optional[Tensor]
PairTensor = Tuple[Tensor, Tensor]
PairOptTensor = Tuple[Optional[Tensor], Optional[Tensor]]
