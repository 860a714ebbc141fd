"""GPU parity of the stages either side of the hot path (csrc/fm_io.cu through the C ABI):
flow preprocessing and export point cloud vs golden vectors of the reference and the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden, max_abs

pytestmark = pytest.mark.gpu
T = torch.as_tensor


def _diff_predictor(videos):
    d = videos[:, 1:, :2] - videos[:, :-1, :2] + 0.25 * videos[:, 1:, 2:3]
    return (0.08 * d.permute(0, 1, 3, 4, 2)).contiguous()


def test_consistency_mask_and_rescale_golden():
    from flowmap_b200 import flow as FL
    g = load_golden("io_flow")
    videos, flow = T(g["videos"]).cuda(), T(g["flow"]).cuda()
    mask = FL.compute_consistency_mask(videos, flow)
    assert max_abs(mask.cpu(), g["mask"]) <= 5e-6
    for name, shape in (("down", (15, 18)), ("up", (33, 40)), ("same", (20, 28)), ("odd", (7, 61))):
        assert max_abs(FL.rescale_flow(flow, shape).cpu(), g[f"flow_{name}"]) <= 2e-6
        assert max_abs(FL.rescale_mask(T(g["mask"]).cuda(), shape).cpu(), g[f"mask_{name}"]) <= 2e-6
    # reverse=True == the reference's flip / predict / flip recipe
    rev = FL.compute_consistency_mask(videos, flow, reverse=True)
    ref = FL.compute_consistency_mask(videos.flip(dims=(1,)).contiguous(), flow.flip(dims=(1,)).contiguous())
    assert torch.equal(rev, ref.flip(dims=(1,)))


def test_bidirectional_flow_golden():
    from flowmap_b200 import flow as FL
    from flowmap_b200.types import Batch
    g = load_golden("io_flow")
    videos = T(g["videos"]).cuda()
    b, f = videos.shape[:2]
    batch = Batch(videos, torch.arange(f)[None].expand(b, f), ["s"] * b, ["d"] * b)
    flows = FL.compute_bidirectional_flow(_diff_predictor, batch, (16, 24))
    assert max_abs(flows.forward.cpu(), g["bi_forward"]) <= 2e-6
    assert max_abs(flows.backward.cpu(), g["bi_backward"]) <= 2e-6
    assert max_abs(flows.forward_mask.cpu(), g["bi_forward_mask"]) <= 5e-6
    assert max_abs(flows.backward_mask.cpu(), g["bi_backward_mask"]) <= 5e-6


def test_preprocessing_vs_oracle_at_video_size():
    """150-frame-scale check on a 6 x 360 x 640 slice: oracle (CPU) vs kernels, out-of-frame flows,
    NaN flow (no tap -> delta = |source|), non-integer rescale factors."""
    from flowmap_b200 import flow as FL
    from oracle import flowmap_io_oracle as IO
    gen = torch.Generator().manual_seed(11)
    b, f, h, w = 1, 6, 360, 640
    lo = torch.rand(b * f, 3, 9, 16, generator=gen)
    videos = torch.nn.functional.interpolate(lo, (h, w), mode="bilinear", align_corners=False).reshape(b, f, 3, h, w)
    flow = 0.02 * torch.randn(b, f - 1, h, w, 2, generator=gen)
    flow[0, 0, :4, :4] = 2.0       # far outside
    flow[0, 1, 5, 7] = -0.5
    ref = IO.consistency_mask(videos, flow)
    out = FL.compute_consistency_mask(videos.cuda(), flow.cuda()).cpu()
    assert max_abs(out, ref) <= 5e-6
    for shape in ((180, 320), (352, 624), (400, 700)):
        # values up to 2.0 next to 0.02: one ulp of the interpolation weight is ~1e-6 of output
        assert max_abs(FL.rescale_flow(flow.cuda(), shape).cpu(), IO.rescale_flow(flow, shape)) <= 1e-5
        assert max_abs(FL.rescale_mask(ref.cuda(), shape).cpu(), IO.rescale_mask(ref, shape)) <= 2e-6
    assert max_abs(FL.resize_videos(videos.cuda(), (90, 161)).cpu(),
                   IO.resize_bilinear(videos.reshape(b * f, 3, h, w), (90, 161)).reshape(b, f, 3, 90, 161)) <= 2e-6


def test_world_points_and_export(tmp_path):
    from flowmap_b200 import export as EX
    from flowmap_b200.types import ModelExports
    g = load_golden("io_export")
    ext, k, depths = T(g["extrinsics"]).cuda(), T(g["intrinsics"]).cuda(), T(g["depths"]).cuda()
    pts = EX.world_points(depths, k, ext)
    assert np.allclose(pts.cpu().numpy(), g["points"], rtol=1e-5, atol=3e-6)
    f, h, w = depths.shape
    colors = torch.rand(1, f, 3, h, w, device="cuda")
    frames = []
    for i in range(f):
        p = tmp_path / f"frame_{i:03d}.png"
        p.write_bytes(b"not really a png")
        frames.append(p)
    exports = ModelExports(ext[None], k[None], colors, depths[None])
    EX.export_to_colmap(exports, frames, (h + 4, w + 6), torch.zeros(1, f, 3, 48, 64), tmp_path / "colmap")
    assert (tmp_path / "colmap/sparse/0/cameras.bin").read_bytes() == g["cameras_bin"].tobytes()
    assert (tmp_path / "colmap/sparse/0/images.bin").read_bytes() == g["images_bin"].tobytes()
    xyz, rgb = EX.read_ply(tmp_path / "colmap/sparse/0/points3D.ply")
    assert np.allclose(xyz, g["points"], rtol=1e-5, atol=3e-6)
    want = (colors[0].permute(0, 2, 3, 1).reshape(-1, 3).cpu().numpy() * 255).astype(np.uint8)
    assert np.array_equal((rgb * 255).round().astype(np.uint8), want)
    assert sorted(p.name for p in (tmp_path / "colmap/images").iterdir()) == [p.name for p in frames]


def test_io_rejects_cpu_tensors():
    from flowmap_b200 import flow as FL
    with pytest.raises(ValueError):
        FL.rescale_mask(torch.zeros(1, 1, 4, 4), (2, 2))
