"""Times the full fused step (softmin + flow + tracking + Adam) and the flow-only step at the
BASELINE shape, optionally against another build of the library (FM_SO=path/to/lib.so).
Usage: python tools/step_bench.py [label]"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import flowmap_b200._lib as _l  # noqa: E402
if os.environ.get("FM_SO"):
    _l.SO_PATH = Path(os.environ["FM_SO"]).resolve()
import bench  # noqa: E402
from flowmap_b200.overfit import FusedOverfitter, OverfitCfg  # noqa: E402
from flowmap_b200.types import Batch, Flows, Tracks  # noqa: E402

F, H, W = bench.F_, bench.H_, bench.W_
dev = torch.device("cuda:0")
inp = bench.synthetic_inputs(F, H, W, seed=0)
batch = Batch(torch.zeros(1, 1, 1, 1, 1, device=dev).expand(1, F, 3, H, W), torch.arange(F, device=dev)[None], ["s"], ["d"])
flows = Flows(*(inp[k].to(dev) for k in ("fwd", "bwd", "fmask", "bmask")))
tracks = [Tracks(xy, vis, s) for xy, vis, s in bench.synthetic_track_arrays(F, seed=0)]


def timed(o, steps=40):
    with torch.no_grad():
        o.model.backbone.depth.copy_(inp["depth"])
        o.model.backbone.weights.copy_(inp["wparam"])
    o.global_step = bench.START_STEP
    for _ in range(5):
        o.training_step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        out = o.training_step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, float(out[0])


label = sys.argv[1] if len(sys.argv) > 1 else ""
full, loss = timed(FusedOverfitter(OverfitCfg(intrinsics="softmin", use_tracking=True), batch, flows, tracks, device=dev))
flow_only, _ = timed(FusedOverfitter(OverfitCfg(), batch, flows, device=dev))
print(f"{label:10s} full step {full:.4f} ms   flow-only step {flow_only:.4f} ms   (loss {loss:.4f})", flush=True)
