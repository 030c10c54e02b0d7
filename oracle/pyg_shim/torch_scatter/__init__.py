"""TEST INFRASTRUCTURE ONLY - stand-in for torch-scatter 2.1.1 (requirements.txt:18).

The reference calls exactly one symbol, ``scatter_add`` (TwoWL/utils.py:5,10), from
inside ``@torch.jit.script`` functions, so this restatement must itself be scriptable.
Semantics restated from the torch-scatter 2.1.1 documentation: sum ``src`` into
``out`` at positions ``index`` along ``dim``; ``dim_size`` fixes the output length.
"""
from typing import Optional

import torch
from torch import Tensor


@torch.jit.script
def scatter_add(src: Tensor, index: Tensor, dim: int = -1,
                out: Optional[Tensor] = None,
                dim_size: Optional[int] = None) -> Tensor:
    if dim < 0:
        dim = src.dim() + dim
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)
