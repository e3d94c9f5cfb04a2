"""Pure-torch stand-in for the `torch_scatter` symbols the VQ-GNN reference imports
(v{1,2}/convs.py:17, v{1,2}/utils/vq_softmax.py:5).  TEST INFRASTRUCTURE ONLY (oracle/).
`segment_csr(src, ptr, reduce='sum')` sums contiguous row segments described by a CSR
pointer; `scatter(_add)` is index_add along `dim`."""
from typing import Optional

import torch
from torch import Tensor


def scatter(src: Tensor, index: Tensor, dim: int = 0, out: Optional[Tensor] = None,
            dim_size: Optional[int] = None, reduce: str = "sum") -> Tensor:
    assert reduce in ("sum", "add") and dim == 0
    n = int(index.max()) + 1 if dim_size is None else dim_size
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add(0, index, src)


def scatter_add(src, index, dim: int = 0, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "sum")


def segment_csr(src: Tensor, indptr: Tensor, out: Optional[Tensor] = None,
                reduce: str = "sum") -> Tensor:
    assert reduce in ("sum", "add")
    ptr = indptr.reshape(-1)
    n = ptr.numel() - 1
    seg = torch.repeat_interleave(torch.arange(n, device=src.device), ptr[1:] - ptr[:-1])
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add(0, seg, src)


def gather_csr(src: Tensor, indptr: Tensor, out: Optional[Tensor] = None) -> Tensor:
    ptr = indptr.reshape(-1)
    return torch.repeat_interleave(src, ptr[1:] - ptr[:-1], dim=0)
