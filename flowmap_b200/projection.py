"""Function-level mirror of flowmap/model/projection.py for callers outside the fused hot
path (visualiser, COLMAP export, custom losses).  Same names and argument meaning; every
body is a kernel call through flowmap_b200.ops (CUDA float32 tensors only).

Broadcasting: the reference accepts arbitrary `*#batch` shapes.  Supported here are the
shapes its own callers use: intrinsics / extrinsics / transformations carry the leading
(batch...) dimensions and singleton dimensions where the points' grid/point dimensions are
(`rearrange(intrinsics, "b f i j -> b f () () i j")`, model.py:75-79); anything else raises.
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import ops


def sample_image_grid(shape, device=torch.device("cpu")):
    """projection.py:93-113: ((*shape, dim) float xy coordinates of the pixel centres,
    (*shape, dim) integer ij indices)."""
    indices = [torch.arange(n, device=device) for n in shape]
    ij = torch.stack(torch.meshgrid(*indices, indexing="ij"), dim=-1)
    coords = [(i.float() + 0.5) / n for i, n in zip(indices, shape)]
    xy = torch.stack(torch.meshgrid(*reversed(coords), indexing="xy"), dim=-1)
    return xy, ij


def _split(mat: Tensor, target_batch: tuple, mat_dims: int):
    """Leading item dimensions of a (*#batch, r, c) matrix whose remaining batch dimensions
    are singletons; returns (number of leading dims, items)."""
    lead = mat.shape[:-mat_dims]
    if len(lead) > len(target_batch):
        raise ValueError("flowmap_b200.projection: matrix has more batch dimensions than the points")
    lead = (1,) * (len(target_batch) - len(lead)) + tuple(lead)
    k = len(lead)
    while k > 0 and lead[k - 1] == 1:
        k -= 1
    for a, b in zip(lead[:k], target_batch[:k]):
        if a != b:
            raise ValueError("flowmap_b200.projection: unsupported broadcast (matrix batch dims must "
                             "match the leading point dims and be 1 elsewhere)")
    return k


def unproject(coordinates: Tensor, z: Tensor, intrinsics: Tensor) -> Tensor:
    """projection.py:76-90: z * K^-1 [x y 1]^T.  coordinates (*#batch, 2), z (*#batch),
    intrinsics (*#batch, 3, 3) -> (*batch, 3)."""
    batch = tuple(z.shape)
    k = _split(intrinsics, batch, 2)
    items = 1
    for d in batch[:k]:
        items *= d
    n = z.numel() // max(items, 1)
    k4 = ops.intrinsics_to_k4(intrinsics.reshape(-1, 3, 3)[:items] if intrinsics.numel() == items * 9
                              else intrinsics.expand(*batch[:k], *([1] * (len(batch) - k)), 3, 3)
                              .reshape(items, -1, 3, 3)[:, 0])
    grid_dims = batch[k:]
    if tuple(coordinates.shape[:-1]) == grid_dims:  # one coordinate set shared by every item
        xy = coordinates.reshape(1, n, 2)
    else:
        xy = coordinates.expand(*batch, 2).reshape(items, n, 2)
    out = ops.unproject_points(xy.contiguous(), z.reshape(items, n).contiguous(), k4.contiguous())
    return out.reshape(*batch, 3)


def _rigid_inverse(ext: Tensor) -> Tensor:
    """(..., 4, 4) camera-to-world -> (..., 3, 4) world-to-camera [R^T | -R^T t]."""
    r, t = ext[..., :3, :3], ext[..., :3, 3:]
    rt_ = r.transpose(-1, -2)
    return torch.cat((rt_, -rt_ @ t), dim=-1)


def reproject_points(xyz: Tensor, relative_transformations: Tensor, intrinsics: Tensor) -> Tensor:
    """projection.py:116-134 (forward only)."""
    batch = tuple(xyz.shape[:-1])
    k = max(_split(relative_transformations, batch, 2), _split(intrinsics, batch, 2))
    items = 1
    for d in batch[:k]:
        items *= d
    n = xyz.numel() // 3 // max(items, 1)
    tail = [1] * (len(batch) - k)
    rt = relative_transformations.expand(*batch[:k], *tail, 4, 4).reshape(items, -1, 4, 4)[:, 0, :3, :]
    k4 = ops.intrinsics_to_k4(intrinsics.expand(*batch[:k], *tail, 3, 3).reshape(items, -1, 3, 3)[:, 0])
    xy = ops.reproject(xyz.reshape(items, n, 3).contiguous(), rt.contiguous(), k4.contiguous())
    return xy.reshape(*batch, 2)


def project(points: Tensor, extrinsics: Tensor, intrinsics: Tensor, epsilon: float = 1e-5):
    """projection.py:61-73: world points -> (xy, in_front_of_camera); forward only."""
    if abs(epsilon - 1e-5) > 1e-12:
        raise ValueError("flowmap_b200.projection.project: epsilon is fixed at the reference default 1e-5")
    batch = tuple(points.shape[:-1])
    k = max(_split(extrinsics, batch, 2), _split(intrinsics, batch, 2))
    items = 1
    for d in batch[:k]:
        items *= d
    n = points.numel() // 3 // max(items, 1)
    tail = [1] * (len(batch) - k)
    w2c = _rigid_inverse(extrinsics.expand(*batch[:k], *tail, 4, 4).reshape(items, -1, 4, 4)[:, 0])
    k4 = ops.intrinsics_to_k4(intrinsics.expand(*batch[:k], *tail, 3, 3).reshape(items, -1, 3, 3)[:, 0])
    xy, front = ops.reproject(points.reshape(items, n, 3).contiguous(), w2c.contiguous(),
                              k4.contiguous(), with_in_front=True)
    return xy.reshape(*batch, 2), front.reshape(batch)


earlier = lambda x: x[:, :-1]  # noqa: E731
later = lambda x: x[:, 1:]  # noqa: E731


def compute_forward_flow(surfaces: Tensor, extrinsics: Tensor, intrinsics: Tensor) -> Tensor:
    """projection.py:143-162: positions of frame-i points in frame i+1 (forward only)."""
    rel = torch.cat((_rigid_inverse(later(extrinsics)),
                     torch.tensor([0., 0., 0., 1.], device=extrinsics.device).expand(
                         *later(extrinsics).shape[:-2], 1, 4)), dim=-2) @ earlier(extrinsics)
    extra = surfaces.ndim - 3
    shape = (*rel.shape[:2], *([1] * extra))
    return reproject_points(earlier(surfaces), rel.reshape(*shape, 4, 4),
                            later(intrinsics).reshape(*shape, 3, 3))


def compute_backward_flow(surfaces: Tensor, extrinsics: Tensor, intrinsics: Tensor) -> Tensor:
    """projection.py:165-184: positions of frame-(i+1) points in frame i (forward only)."""
    rel = torch.cat((_rigid_inverse(earlier(extrinsics)),
                     torch.tensor([0., 0., 0., 1.], device=extrinsics.device).expand(
                         *earlier(extrinsics).shape[:-2], 1, 4)), dim=-2) @ later(extrinsics)
    extra = surfaces.ndim - 3
    shape = (*rel.shape[:2], *([1] * extra))
    return reproject_points(later(surfaces), rel.reshape(*shape, 4, 4),
                            earlier(intrinsics).reshape(*shape, 3, 3))


def get_extrinsics(inverse_relative_transformations: Tensor) -> Tensor:
    """projection.py:187-210: (*batch, pair, 4, 4) -> (*batch, pair+1, 4, 4), differentiable."""
    *batch, p, _, _ = inverse_relative_transformations.shape
    rt = inverse_relative_transformations.reshape(-1, p, 4, 4)[:, :, :3, :].contiguous()
    return ops.pose_chain(rt).reshape(*batch, p + 1, 4, 4)


def align_surfaces(depths: Tensor, intrinsics: Tensor, backward_flows: Tensor,
                   backward_weights: Tensor, indices: Tensor | None = None) -> Tensor:
    """projection.py:213-252 from depths + intrinsics (the point cloud is formed inside the
    kernel; the reference's variant takes the materialised surfaces): extrinsics (b, f, 4, 4)."""
    rt = ops.procrustes_poses(depths, backward_weights, ops.intrinsics_to_k4(intrinsics),
                              backward_flows, indices)
    return ops.pose_chain(rt)
