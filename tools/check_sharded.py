"""Multi-GPU check (run under torchrun, one rank per GPU): a pair-sharded fused optimisation
must reproduce the unsharded one.  Every rank builds the same full problem, takes its shard,
runs STEPS sharded steps (one all-reduce each); rank 0 also runs the unsharded optimisation
and every rank compares its shard of the parameters against it."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from flowmap_b200 import parallel  # noqa: E402
from flowmap_b200.overfit import FusedOverfitter, OverfitCfg, ShardedFusedOverfitter  # noqa: E402
from flowmap_b200.types import Batch, Flows  # noqa: E402

STEPS = 3
F, H, W = 17, 96, 128
rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator().manual_seed(0)
depth = 1.0 + torch.rand(F, H, W, generator=g)
wparam = 0.01 * torch.randn(F - 1, H, W, generator=g)
flows = Flows(0.01 * torch.randn(1, F - 1, H, W, 2, generator=g), 0.01 * torch.randn(1, F - 1, H, W, 2, generator=g),
              torch.rand(1, F - 1, H, W, generator=g), torch.rand(1, F - 1, H, W, generator=g))


from flowmap_b200.types import Tracks  # noqa: E402

# segment layout of flowmap/tracking/__init__.py:49-70 (a segment every 4 frames, radius 6)
tracks = []
for mid in range(0, F, 4):
    lo_f, hi_f = max(0, mid - 6), min(F, mid + 7)
    tracks.append(Tracks(torch.rand(1, hi_f - lo_f, 128, 2, generator=g),
                         torch.rand(1, hi_f - lo_f, 128, generator=g) < 0.7, lo_f))
idx = torch.randperm(H * W, generator=g)[:512].to(dev)


def make(cls, cfg, d, wp, fl, *extra, **kw):
    f = d.shape[0]
    batch = Batch(torch.zeros(1, 1, 1, 1, 1, device=dev).expand(1, f, 3, H, W), torch.arange(f, device=dev)[None], ["s"], ["d"])
    o = cls(cfg, batch, fl.to(dev), *extra, device=dev, **kw)
    with torch.no_grad():
        o.model.backbone.depth.copy_(d)
        o.model.backbone.weights.copy_(wp)
    o.injected_indices = idx
    return o


plan = parallel.make_plan(F - 1)
d_l, w_l, fl_l = parallel.shard_inputs(plan, depth, wparam, flows)
a, b = plan.pair_range
rel = lambda x, y: float((x - y).norm() / y.norm())  # noqa: E731
CASES = {
    "flow only, regressed focal": OverfitCfg(lr=1e-3),
    "flow + tracking, regressed focal": OverfitCfg(lr=1e-3, use_tracking=True, tracking_enable_after=1),
    "flow + tracking, softmin sweep then hand-over": OverfitCfg(lr=1e-3, use_tracking=True, tracking_enable_after=1,
                                                              intrinsics="softmin", softmin_points=512,
                                                              regression_after=3, regression_window=2),
}
CASES["flow only, regressed focal, replayed as a CUDA graph (exchange captured)"] = OverfitCfg(lr=1e-3)
for name, cfg in CASES.items():
    graph = "CUDA graph" in name
    steps = 6 if graph else (5 if cfg.intrinsics == "softmin" else STEPS)
    trk = tracks if cfg.use_tracking else None
    sh = make(ShardedFusedOverfitter, cfg, d_l, w_l, fl_l, plan, tracks=trk)
    sh.use_cuda_graph = graph
    losses = [float(sh.training_step()[0]) for _ in range(steps)]
    full = make(FusedOverfitter, cfg, depth, wparam, flows, trk)
    ref_losses = [float(full.training_step()[0]) for _ in range(steps)]
    e_d = rel(sh.model.backbone.depth.detach(), full.model.backbone.depth.detach()[a:b + 1])
    upd_ref = full.model.backbone.weights.detach()[a:b] - wparam[a:b].to(dev)
    upd = sh.model.backbone.weights.detach() - w_l.to(dev)
    e_w = rel(upd, upd_ref)
    e_f = abs(float(sh._focal) - float(full._focal))
    e_l = max(abs(x - y) / abs(y) for x, y in zip(losses, ref_losses))
    print(f"[{name}] rank {rank}/{world} pairs {plan.pair_range}: loss rel err {e_l:.2e}, depth {e_d:.2e}, "
          f"weight-update {e_w:.2e}, focal {e_f:.2e}, collective {sh.reducer.bytes_per_step()} B/step", flush=True)
    assert e_l < 1e-4 and e_d < 1e-5 and e_w < 2e-2 and e_f < 1e-5, name
    assert not graph or len(sh._graphs) == 1, "the step was not captured"
    sh._graphs.clear()  # a graph that captured NCCL kernels must go before the communicator does
    del sh, full
    torch.cuda.synchronize()
    dist.barrier()
dist.destroy_process_group()
