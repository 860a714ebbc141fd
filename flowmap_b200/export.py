"""Export side of the hot path: ``ModelExports`` -> COLMAP sparse model + point cloud
(flowmap/export/colmap.py:56-111, 171-213; flowmap/misc/cropping.py:55-70).

The per-pixel world-space point cloud is computed by ``fm_world_points`` (csrc/fm_io.cu); the
COLMAP binary files are a few hundred bytes of host-side ``struct`` packing, byte-identical to
what the reference writes through third_party/colmap/read_write_model.py:188-202, 334-352.
"""
from __future__ import annotations

import shutil
import struct
from pathlib import Path
from typing import Sequence

import numpy as np
import torch
from torch import Tensor

from ._lib import check, lib
from .ops import _canon, _ptr, _stream, intrinsics_to_k4
from .types import ModelExports

PINHOLE_MODEL_ID = 1  # read_write_model.py CAMERA_MODELS: PINHOLE, 4 params


def center_crop_intrinsics(intrinsics: Tensor | None, old_shape, new_shape) -> Tensor | None:
    """misc/cropping.py:55-70."""
    if intrinsics is None:
        return None
    h_old, w_old = old_shape
    h_new, w_new = new_shape
    intrinsics = intrinsics.clone()
    intrinsics[..., 0, 0] *= w_old / w_new
    intrinsics[..., 1, 1] *= h_old / h_new
    return intrinsics


def world_points(depths: Tensor, intrinsics: Tensor, extrinsics: Tensor) -> Tensor:
    """export/colmap.py:84-101: depths (f, h, w), intrinsics (f, 3, 3), camera-to-world
    extrinsics (f, 4, 4) -> (f*h*w, 3)."""
    depths = _canon(depths, "depths")
    extrinsics = _canon(extrinsics, "extrinsics")
    k4 = _canon(intrinsics_to_k4(intrinsics), "intrinsics")
    f, h, w = depths.shape
    xyz = torch.empty((f * h * w, 3), dtype=torch.float32, device=depths.device)
    with torch.cuda.device(depths.device):
        check(lib().fm_world_points(_ptr(depths), _ptr(k4), _ptr(extrinsics), _ptr(xyz), f, h, w, _stream()),
              "fm_world_points")
    return xyz


def _rotation_to_qvec(rotation: np.ndarray) -> np.ndarray:
    from scipy.spatial.transform import Rotation
    qx, qy, qz, qw = Rotation.from_matrix(rotation).as_quat()
    return np.array((qw, qx, qy, qz))


def write_colmap_model(path: Path, extrinsics: Tensor, intrinsics: Tensor, image_names: Sequence[str],
                       image_shape) -> None:
    """export/colmap.py:171-213: cameras.bin + images.bin (no points3D, as the reference)."""
    h, w = image_shape
    # on the host, so that the files do not depend on which device held the tensors
    extrinsics, intrinsics = extrinsics.detach().cpu(), intrinsics.detach().cpu()
    cams = struct.pack("<Q", len(intrinsics))
    for index, k in enumerate(intrinsics):
        k = k.detach().clone()
        k[0] *= w
        k[1] *= h
        cams += struct.pack("<iiQQ", index + 1, PINHOLE_MODEL_ID, w, h)
        cams += struct.pack("<dddd", float(k[0, 0]), float(k[1, 1]), float(k[0, 2]), float(k[1, 2]))
    imgs = struct.pack("<Q", len(extrinsics))
    for index, (c2w, name) in enumerate(zip(extrinsics, image_names)):
        w2c = c2w.inverse().numpy()
        imgs += struct.pack("<i", index + 1)
        imgs += struct.pack("<dddd", *_rotation_to_qvec(w2c[:3, :3]).tolist())
        imgs += struct.pack("<ddd", *w2c[:3, 3].tolist())
        imgs += struct.pack("<i", index + 1)
        imgs += str(name).encode("utf-8") + b"\x00"
        imgs += struct.pack("<Q", 0)
    path = Path(path)
    path.mkdir(exist_ok=True, parents=True)
    (path / "cameras.bin").write_bytes(cams)
    (path / "images.bin").write_bytes(imgs)


def read_colmap_model(path: Path, device="cpu", reorder: bool = True):
    """export/colmap.py:114-168: -> (extrinsics (f,4,4) camera-to-world, intrinsics (f,3,3)
    normalised, image names), sorted by name when ``reorder``."""
    from scipy.spatial.transform import Rotation
    path = Path(path)
    cam_bytes, img_bytes = (path / "cameras.bin").read_bytes(), (path / "images.bin").read_bytes()
    cameras, off = {}, 8
    for _ in range(struct.unpack_from("<Q", cam_bytes, 0)[0]):
        cam_id, model_id, width, height = struct.unpack_from("<iiQQ", cam_bytes, off)
        off += 24
        num_params = {0: 3, 1: 4}.get(model_id)
        if num_params is None:
            raise ValueError(f"flowmap_b200: unsupported COLMAP camera model {model_id}")
        params = struct.unpack_from("<" + "d" * num_params, cam_bytes, off)
        off += 8 * num_params
        cameras[cam_id] = (model_id, width, height, params)
    records, off = [], 8
    for _ in range(struct.unpack_from("<Q", img_bytes, 0)[0]):
        _, qw, qx, qy, qz, tx, ty, tz, cam_id = struct.unpack_from("<idddddddi", img_bytes, off)
        off += 64
        end = img_bytes.index(b"\x00", off)
        name = img_bytes[off:end].decode("utf-8")
        off = end + 1
        num_points2d = struct.unpack_from("<Q", img_bytes, off)[0]
        off += 8 + 24 * num_points2d
        model_id, width, height, params = cameras[cam_id]
        fx, fy, cx, cy = (params[0], params[0], params[1], params[2]) if model_id == 0 else params
        k = torch.eye(3, dtype=torch.float32)
        k[0, 0], k[1, 1], k[0, 2], k[1, 2] = fx, fy, cx, cy
        k[0] /= width
        k[1] /= height
        w2c = torch.eye(4, dtype=torch.float32)
        w2c[:3, :3] = torch.tensor(Rotation.from_quat([qx, qy, qz, qw]).as_matrix(), dtype=torch.float32)
        w2c[:3, 3] = torch.tensor((tx, ty, tz), dtype=torch.float32)
        records.append((name, w2c.inverse(), k))
    if reorder:
        records.sort(key=lambda r: r[0])
    return (torch.stack([r[1] for r in records]).to(device), torch.stack([r[2] for r in records]).to(device),
            [r[0] for r in records])


_PLY_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4"),
                       ("red", "u1"), ("green", "u1"), ("blue", "u1")])


def write_ply(path: Path, xyz: np.ndarray, rgb: np.ndarray) -> None:
    """export/colmap.py:31-53: vertex element with x y z, zero normals, uint8 colours (rgb in
    [0, 1] scaled by 255 and truncated), binary little endian (plyfile's default)."""
    elements = np.zeros(xyz.shape[0], dtype=_PLY_DTYPE)
    elements["x"], elements["y"], elements["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    colours = (rgb * 255).astype(np.uint8)
    elements["red"], elements["green"], elements["blue"] = colours[:, 0], colours[:, 1], colours[:, 2]
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % xyz.shape[0]
    header += "".join(f"property {'float' if t.kind == 'f' else 'uchar'} {n}\n"
                      for n, (t, _) in _PLY_DTYPE.fields.items())
    header += "end_header\n"
    with open(path, "wb") as fid:
        fid.write(header.encode("ascii"))
        fid.write(elements.tobytes())


def read_ply(path: Path):
    """export/colmap.py:18-28: -> xyz (n, 3), rgb (n, 3) in [0, 1]."""
    data = Path(path).read_bytes()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    header = data[:end].decode("ascii").split("\n")
    count = next(int(line.split()[2]) for line in header if line.startswith("element vertex"))
    elements = np.frombuffer(data, dtype=_PLY_DTYPE, count=count, offset=end)
    xyz = np.stack([elements["x"], elements["y"], elements["z"]], axis=1)
    rgb = np.stack([elements["red"], elements["green"], elements["blue"]], axis=1) / 255.0
    return xyz, rgb


def export_to_colmap(exports: ModelExports, frame_paths, uncropped_exports_shape, uncropped_videos: Tensor,
                     path: Path) -> None:
    """export/colmap.py:56-111: sparse/0/{cameras,images}.bin (intrinsics un-cropped, pixel units
    of the full-resolution frames), sparse/0/points3D.ply (every pixel of every frame in world
    space with its colour), images/ (copies of the frames)."""
    path = Path(path)
    _, _, h_cropped, w_cropped = exports.depths.shape
    intrinsics = center_crop_intrinsics(exports.intrinsics, (h_cropped, w_cropped), uncropped_exports_shape)
    sparse_path = path / "sparse/0"
    _, _, _, h_full, w_full = uncropped_videos.shape
    write_colmap_model(sparse_path, exports.extrinsics[0], intrinsics[0], [Path(p).name for p in frame_paths],
                       (h_full, w_full))
    points = world_points(exports.depths[0], exports.intrinsics[0], exports.extrinsics[0])
    colors = exports.colors[0].permute(0, 2, 3, 1).reshape(-1, 3)
    write_ply(sparse_path / "points3D.ply", points.detach().cpu().numpy(), colors.detach().cpu().numpy())
    (path / "images").mkdir(exist_ok=True, parents=True)
    for frame_path in frame_paths:
        shutil.copy(frame_path, path / "images" / Path(frame_path).name)
