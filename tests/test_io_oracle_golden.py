"""Pins oracle/flowmap_io_oracle.py (flow preprocessing + export) against
outputs of the unmodified reference (tests/golden/io_*.npz from make_golden_io.py).  CPU only."""
import numpy as np
import torch

from conftest import load_golden, max_abs
from oracle import flowmap_io_oracle as IO

T = torch.as_tensor


def _diff_predictor(videos):
    """Same stand-in predictor as make_golden_io.py::DiffPredictor."""
    d = videos[:, 1:, :2] - videos[:, :-1, :2] + 0.25 * videos[:, 1:, 2:3]
    return 0.08 * d.permute(0, 1, 3, 4, 2)


def test_consistency_mask_and_rescale():
    g = load_golden("io_flow")
    videos, flow = T(g["videos"]), T(g["flow"])
    mask = IO.consistency_mask(videos, flow)
    assert max_abs(mask, g["mask"]) <= 2e-6
    # zero padding is exercised: some targets are outside the frame
    assert (mask < 0.05).any() and (mask > 0.9).any()
    for name, shape in (("down", (15, 18)), ("up", (33, 40)), ("same", (20, 28)), ("odd", (7, 61))):
        assert max_abs(IO.rescale_flow(flow, shape), g[f"flow_{name}"]) <= 1e-6
        assert max_abs(IO.rescale_mask(T(g["mask"]), shape), g[f"mask_{name}"]) <= 1e-6


def test_bidirectional_flows():
    g = load_golden("io_flow")
    fwd, bwd, fm, bm = IO.bidirectional_flows(_diff_predictor, T(g["videos"]), (16, 24))
    assert max_abs(fwd, g["bi_forward"]) <= 1e-6 and max_abs(bwd, g["bi_backward"]) <= 1e-6
    assert max_abs(fm, g["bi_forward_mask"]) <= 2e-6 and max_abs(bm, g["bi_backward_mask"]) <= 2e-6


def test_export():
    g = load_golden("io_export")
    ext, k, depths = T(g["extrinsics"]), T(g["intrinsics"]), T(g["depths"])
    h, w = depths.shape[1:]
    cropped = IO.center_crop_intrinsics(k[None], (h, w), (h + 4, w + 6))[0]
    assert max_abs(cropped, g["cropped"]) <= 1e-7
    cams, imgs = IO.colmap_model_bytes(ext, cropped, [str(n) for n in g["names"]], (48, 64))
    assert cams == g["cameras_bin"].tobytes()
    assert imgs == g["images_bin"].tobytes()
    pts = IO.world_points(depths, k, ext)
    assert np.allclose(pts.numpy(), g["points"], rtol=1e-5, atol=2e-6)
