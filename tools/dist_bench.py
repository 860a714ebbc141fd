"""Times the three pixel-parallel ops (C ABI, CUDA events) on iid and on smooth flows.
Usage: python tools/dist_bench.py [label]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os  # noqa: E402
import flowmap_b200._lib as _l  # noqa: E402
if os.environ.get("FM_SO"):  # experiment: time another build of the library
    _l.SO_PATH = Path(os.environ["FM_SO"]).resolve()
from flowmap_b200 import ops  # noqa: E402
from flowmap_b200._lib import lib  # noqa: E402

F, H, W = 150, 360, 640
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
depths = (0.1 + 0.05 * torch.rand(1, F, H, W, device=dev, generator=g))
weights = torch.sigmoid(torch.randn(1, F - 1, H, W, device=dev, generator=g))
fm = torch.rand(1, F - 1, H, W, device=dev, generator=g)
bm = torch.rand(1, F - 1, H, W, device=dev, generator=g)
s = (H * W) ** 0.5
k4 = torch.tensor([0.85 * s / W, 0.85 * s / H, 0.5, 0.5], device=dev).expand(1, F, 4).contiguous()
msum = ops.mask_sum(fm, bm)
ws = ops.workspace(1, F, H, W, dev)
rt = torch.empty(1, F - 1, 3, 4, device=dev)
g_depth, g_w, g_k4, g_rt = (torch.empty_like(depths), torch.empty_like(weights),
                            torch.empty_like(k4), torch.empty_like(rt))
loss = torch.empty((), device=dev)
P = lambda x: x.data_ptr()  # noqa: E731
st = torch.cuda.current_stream().cuda_stream
L = lib()


def flows(kind):
    if kind == "iid":
        return (0.01 * torch.randn(1, F - 1, H, W, 2, device=dev, generator=g),
                0.01 * torch.randn(1, F - 1, H, W, 2, device=dev, generator=g))
    out = []
    for _ in range(2):
        lo = 0.01 * torch.randn(F - 1, 2, H // 16 + 1, W // 16 + 1, device=dev, generator=g)
        up = torch.nn.functional.interpolate(lo, size=(H, W), mode="bilinear", align_corners=True)
        out.append(up.permute(0, 2, 3, 1)[None].contiguous())
    return tuple(out)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


label = sys.argv[1] if len(sys.argv) > 1 else ""
for kind in ("iid", "smooth"):
    fwd, bwd = flows(kind)
    f1 = lambda: L.fm_procrustes_fwd(P(depths), P(k4), P(bwd), P(weights), None, 0, P(rt), P(ws), 1, F, H, W, st)  # noqa: E731
    f2 = lambda: L.fm_flow_loss_fwd_bwd(P(depths), P(k4), P(rt), P(fwd), P(bwd), P(fm), P(bm), P(msum), 0, 0.01,  # noqa: E731
                                        1000.0, 1, P(loss), P(g_depth), P(g_rt), P(g_k4), P(ws), 1, F, H, W, st)
    f3 = lambda: L.fm_procrustes_bwd(P(depths), P(k4), P(bwd), P(weights), None, 0, None, 1, None, P(g_depth),  # noqa: E731
                                     P(g_w), P(g_k4), P(ws), 1, F, H, W, st)
    t1 = timed(f1); f2(); t2 = timed(f2); t3 = timed(f3)
    print(f"{label:12s} {kind:7s} procrustes_fwd {t1:.4f}  flow {t2:.4f}  procrustes_bwd {t3:.4f} ms  "
          f"(loss {float(loss):.4f}, |g_depth| {float(g_depth.norm()):.5e})", flush=True)
