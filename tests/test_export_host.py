"""Host side of the export / checkpoint boundary, CPU only: COLMAP files
byte-identical to the reference's writer, reader round trip, PLY round trip, checkpoint names."""
import numpy as np
import torch

from conftest import load_golden, max_abs

T = torch.as_tensor


def test_colmap_writer_matches_reference_bytes(tmp_path):
    from flowmap_b200.export import center_crop_intrinsics, read_colmap_model, write_colmap_model
    g = load_golden("io_export")
    ext, k = T(g["extrinsics"]), T(g["intrinsics"])
    h, w = g["depths"].shape[1:]
    cropped = center_crop_intrinsics(k[None], (h, w), (h + 4, w + 6))[0]
    assert max_abs(cropped, g["cropped"]) <= 1e-7
    names = [str(n) for n in g["names"]]
    write_colmap_model(tmp_path, ext, cropped, names, (48, 64))
    assert (tmp_path / "cameras.bin").read_bytes() == g["cameras_bin"].tobytes()
    assert (tmp_path / "images.bin").read_bytes() == g["images_bin"].tobytes()
    back_ext, back_k, back_names = read_colmap_model(tmp_path)
    assert back_names == [str(n) for n in g["read_names"]]
    assert max_abs(back_ext, g["read_extrinsics"]) <= 1e-6
    assert max_abs(back_k, g["read_intrinsics"]) <= 1e-6
    assert max_abs(back_ext, ext) <= 1e-5 and max_abs(back_k, cropped) <= 1e-6


def test_colmap_reader_reorders_by_name(tmp_path):
    from flowmap_b200.export import read_colmap_model, write_colmap_model
    g = load_golden("io_export")
    ext, k = T(g["extrinsics"]), T(g["intrinsics"])
    names = ["c.png", "a.png", "e.png", "b.png", "d.png"]
    write_colmap_model(tmp_path, ext, k, names, (10, 20))
    back_ext, _, back_names = read_colmap_model(tmp_path)
    assert back_names == sorted(names)
    order = [names.index(n) for n in back_names]
    assert max_abs(back_ext, ext[order]) <= 1e-5
    _, _, raw_names = read_colmap_model(tmp_path, reorder=False)
    assert raw_names == names


def test_ply_round_trip(tmp_path):
    from flowmap_b200.export import read_ply, write_ply
    rng = np.random.default_rng(0)
    xyz = rng.normal(size=(257, 3)).astype(np.float32)
    rgb = rng.random((257, 3)).astype(np.float32)
    rgb[0] = (0.0, 1.0, 0.5)
    write_ply(tmp_path / "p.ply", xyz, rgb)
    head = (tmp_path / "p.ply").read_bytes()[:400].decode("ascii", "replace")
    assert head.startswith("ply\nformat binary_little_endian 1.0\nelement vertex 257\nproperty float x\n")
    assert "property uchar blue\nend_header\n" in head
    xyz2, rgb2 = read_ply(tmp_path / "p.ply")
    assert np.array_equal(xyz2, xyz)
    assert np.array_equal((rgb2 * 255).round().astype(np.uint8), (rgb * 255).astype(np.uint8))  # truncation


def test_checkpoint_parameter_names_match_reference():
    """model_wrapper_overfit.py:40-49 saves `model.<name>`; a checkpoint of either side must load
    into the other (names, order and shapes of Model.state_dict())."""
    from flowmap_b200.overfit import OverfitCfg, build_model_and_losses
    g = load_golden("io_export")
    for tag, cfg in (("regressed", OverfitCfg(intrinsics="regressed")),
                     ("softmin", OverfitCfg(intrinsics="softmin", softmin_points=64))):
        model, _ = build_model_and_losses(cfg, 5, (12, 16))
        sd = model.state_dict()
        assert list(sd.keys()) == [str(n) for n in g[f"state_{tag}_names"]]
        assert [",".join(map(str, v.shape)) for v in sd.values()] == [str(s) for s in g[f"state_{tag}_shapes"]]
