"""flowmap/model/procrustes.py:7-51 (align_rigid) on explicit point sets."""
from __future__ import annotations

import torch
from torch import Tensor

from . import ops


def align_rigid(p: Tensor, q: Tensor, weights: Tensor) -> Tensor:
    """Rigid transformation (*batch, 4, 4) minimising the weighted squared distance between
    transformed p and q; p, q (*batch, point, 3), weights (*batch, point).  Differentiable
    (closed-form SVD adjoint)."""
    *batch, n, _ = p.shape
    rt = ops.align_rigid_rt(p.reshape(-1, n, 3), q.reshape(-1, n, 3), weights.reshape(-1, n))
    bottom = torch.tensor([0., 0., 0., 1.], device=p.device).expand(rt.shape[0], 1, 4)
    return torch.cat((rt, bottom), dim=1).reshape(*batch, 4, 4)
