# quick GPU check: graph replay == eager trajectories
import sys, torch
sys.path.insert(0, "/root/repo")
import bench
from flowmap_b200.overfit import FusedOverfitter, OverfitCfg
from flowmap_b200.types import Batch, Flows, Tracks
dev = torch.device("cuda:0")
def run(graph, full, f=12, h=72, w=96, steps=8):
    torch.manual_seed(0)
    inp = bench.synthetic_inputs(f, h, w, seed=0)
    batch = Batch(torch.zeros(1, f, 3, h, w, device=dev), torch.arange(f, device=dev)[None], ["s"], ["d"])
    flows = Flows(*(inp[k].to(dev) for k in ("fwd", "bwd", "fmask", "bmask")))
    tracks = [Tracks(xy, vis, s) for xy, vis, s in bench.synthetic_track_arrays(f, n_points=64, interval=3, radius=2)] if full else None
    cfg = OverfitCfg(intrinsics="softmin", use_tracking=True, tracking_enable_after=0, softmin_points=500) if full else OverfitCfg()
    o = FusedOverfitter(cfg, batch, flows, tracks, device=dev)
    o._clock.base_seed = 1234
    with torch.no_grad():
        o.model.backbone.depth.copy_(1.0 + inp["depth"]); o.model.backbone.weights.copy_(inp["wparam"])
    o.use_cuda_graph = graph
    losses = [float(o.training_step()[0]) for _ in range(steps)]
    return losses, o.model.backbone.depth.detach().clone(), o.model.backbone.weights.detach().clone(), len(o._graphs)
for full in (False, True):
    a = run(False, full); b = run(True, full)
    print("full" if full else "flow", "graphs:", b[3], "loss diff", max(abs(x-y) for x, y in zip(a[0], b[0])),
          "depth", float((a[1]-b[1]).abs().max()), "w", float((a[2]-b[2]).abs().max()), a[0][-1], b[0][-1])
