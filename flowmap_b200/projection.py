"""Function-level mirror of flowmap/model/projection.py for callers outside the fused hot
path (visualiser, COLMAP export, custom losses).  Same names and argument meaning; the bodies
are kernel calls through flowmap_b200.ops (CUDA float32 tensors only).

Differentiability: `unproject`, `get_extrinsics` and `align_surfaces` are differentiable through
the kernels' own backward passes.  `project`, `reproject_points` and `compute_{forward,backward}_
flow` have forward-only kernels (the differentiable uses live inside the fused loss kernels): when
an input requires grad under autograd they evaluate the same formulas with ATen ops instead, so a
custom loss built on them after `install()` still trains depth and poses.

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


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, Tensor) and t.requires_grad for t in tensors)


def _project_camera_space_torch(points: Tensor, intrinsics: Tensor, epsilon: float = 1e-5) -> Tensor:
    """projection.py:49-58 with ATen ops (autograd fall-back of the forward-only kernel)."""
    u = points / (points[..., -1:] + epsilon)
    u = u.nan_to_num(posinf=1e8, neginf=-1e8)
    return (intrinsics @ u[..., None])[..., :2, 0]


def _transform_torch(xyz: Tensor, transformation: Tensor) -> Tensor:
    """(T [x; 1])[:3] with (*#batch, 4, 4) transformations (projection.py:27-46)."""
    return (transformation[..., :3, :3] @ xyz[..., None])[..., 0] + transformation[..., :3, 3]


def reproject_points(xyz: Tensor, relative_transformations: Tensor, intrinsics: Tensor) -> Tensor:
    """projection.py:116-134."""
    if _needs_grad(xyz, relative_transformations, intrinsics):
        return _project_camera_space_torch(_transform_torch(xyz, relative_transformations), intrinsics)
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
    """projection.py:61-73: world points -> (xy, in_front_of_camera)."""
    if _needs_grad(points, extrinsics, intrinsics):
        cam = _transform_torch(points, torch.linalg.inv(extrinsics))
        return _project_camera_space_torch(cam, intrinsics, epsilon), cam[..., -1] >= 0
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
    """projection.py:143-162: positions of frame-i points in frame i+1."""
    rel = torch.cat((_rigid_inverse(later(extrinsics)),
                     torch.tensor([0., 0., 0., 1.], device=extrinsics.device).expand(
                         *later(extrinsics).shape[:-2], 1, 4)), dim=-2) @ earlier(extrinsics)
    extra = surfaces.ndim - 3
    shape = (*rel.shape[:2], *([1] * extra))
    return reproject_points(earlier(surfaces), rel.reshape(*shape, 4, 4),
                            later(intrinsics).reshape(*shape, 3, 3))


def compute_backward_flow(surfaces: Tensor, extrinsics: Tensor, intrinsics: Tensor) -> Tensor:
    """projection.py:165-184: positions of frame-(i+1) points in frame i."""
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


def align_surfaces(*args) -> Tensor:
    """projection.py:213-252 -> extrinsics (b, f, 4, 4).  Two call forms:

      align_surfaces(surfaces, backward_flows, backward_weights, indices)   # the reference's signature
      align_surfaces(depths, intrinsics, backward_flows, backward_weights[, indices])

    The second forms the point cloud inside the moment kernel (no (b, f, h, w, 3) tensor).  The first
    accepts ANY xyz image, so it follows the reference's own steps: gather the later points, sample
    the earlier surface at xy + flow (F.grid_sample, bilinear / border / align_corners=False) and
    solve with the align_rigid kernel (closed-form SVD adjoint); differentiable in both forms."""
    if args[0].dim() == 5 and args[0].shape[-1] == 3 and len(args) == 4:
        from .procrustes import align_rigid
        surfaces, backward_flows, backward_weights, indices = args
        b, f, h, w, _ = surfaces.shape
        xy, _ = sample_image_grid((h, w), device=surfaces.device)
        xyz_later = later(surfaces).reshape(b, f - 1, h * w, 3)[:, :, indices]
        xy_earlier = (xy + backward_flows).reshape(b, f - 1, h * w, 2)[:, :, indices]
        sampled = torch.nn.functional.grid_sample(
            earlier(surfaces).reshape(b * (f - 1), h, w, 3).permute(0, 3, 1, 2),
            (xy_earlier * 2 - 1).reshape(b * (f - 1), -1, 1, 2), mode="bilinear", padding_mode="border",
            align_corners=False)
        xyz_earlier = sampled[..., 0].permute(0, 2, 1).reshape(b, f - 1, -1, 3)
        weights = backward_weights.reshape(b, f - 1, h * w)[..., indices]
        return get_extrinsics(align_rigid(xyz_later.contiguous(), xyz_earlier.contiguous(), weights.contiguous()))
    depths, intrinsics, backward_flows, backward_weights = args[:4]
    indices = args[4] if len(args) > 4 else None
    rt = ops.procrustes_poses(depths, backward_weights, ops.intrinsics_to_k4(intrinsics),
                              backward_flows, indices)
    return ops.pose_chain(rt)
